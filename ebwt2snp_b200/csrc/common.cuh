// Shared device helpers: mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP), vector loads,
// decoupled look-back descriptors.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace e2s {

// ---- layout of a shard's device arrays --------------------------------------------------
// Every per-position array is allocated with PAD_L elements before local position 0 and
// enough elements after n_local for the right halo, rounded up so that every tile the
// kernels touch is fully inside the allocation and 16-byte aligned.
constexpr int PAD_L = 176;       // >= 2 (left LCP halo) and >= 160 (K2's fused prefilter reads the BWT bytes of clusters that end in a
                                 // tile but start up to 149 positions before it); multiple of 16 so byte arrays stay 16B aligned
constexpr int HALO_R = 160;      // >= E2S_MAX_C_LEN + 1, multiple of 16
constexpr int MAX_C_LEN = 150;

// ---- mbarrier / bulk-copy PTX wrappers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map needed); bytes % 16 == 0,
// both addresses 16-byte aligned; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- decoupled look-back descriptors ------------------------------------------------------
// One 64-bit word per tile: status in the top 2 bits, 62-bit payload.  Written with a single
// relaxed 64-bit store, read with a volatile 64-bit load, so status and payload are always
// observed together.
constexpr uint64_t ST_INVALID = 0;    // not published yet
constexpr uint64_t ST_AGGREGATE = 1;  // payload = this tile alone
constexpr uint64_t ST_INCLUSIVE = 2;  // payload = everything up to and including this tile
constexpr uint64_t ST_SHIFT = 62;
constexpr uint64_t ST_PAYLOAD = (uint64_t(1) << ST_SHIFT) - 1;

__device__ __forceinline__ void desc_store(uint64_t* d, uint64_t status, uint64_t payload) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(d), "l"((status << ST_SHIFT) | payload) : "memory");
}
__device__ __forceinline__ uint64_t desc_load(const uint64_t* d) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(d) : "memory");
    return v;
}

__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }

}  // namespace e2s
