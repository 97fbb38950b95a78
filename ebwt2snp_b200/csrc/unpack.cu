// Staging kernels: de-interleave the on-disk array-of-structs EGSA records into the
// structure-of-arrays layout the hot kernels stream (replaces the field-by-field istream::read of
// egsa_stream::read_el, ref:include.hpp:120-199).  Reported under "staging", not under the hot path.

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"

namespace e2s {

constexpr int UP_THREADS = 256;
constexpr int UP_RECS = 2048;  // records per tile (multiple of 16 => tile byte offset is 16B aligned)

__device__ __forceinline__ uint32_t load_le(const uint8_t* s, int nbytes) {
    uint32_t v = 0;
    const int nb = nbytes < 4 ? nbytes : 4;  // wider fields are truncated to 32 bits (ref:include.hpp:131)
    for (int b = 0; b < nb; ++b) v |= uint32_t(s[b]) << (8 * b);
    return v;
}

__global__ void __launch_bounds__(UP_THREADS) k_unpack_gesa(const uint8_t* __restrict__ rec, uint64_t count, int x, int y,
                                                            int z, uint32_t* __restrict__ lcp, uint32_t* __restrict__ text,
                                                            uint32_t* __restrict__ suff, uint8_t* __restrict__ bwt) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    const int rs = x + y + z + 1;
    const uint64_t n_tiles = (count + UP_RECS - 1) / UP_RECS;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t r0 = t * UP_RECS;
        const uint32_t nrec = uint32_t(count - r0 < UP_RECS ? count - r0 : UP_RECS);
        const uint32_t nbytes = nrec * rs;
        const uint4* src = reinterpret_cast<const uint4*>(rec + r0 * rs);  // 16B aligned: r0 % 16 == 0
        uint4* dst = reinterpret_cast<uint4*>(s_raw);
        for (uint32_t i = threadIdx.x; i < (nbytes + 15) / 16; i += UP_THREADS) dst[i] = __ldg(src + i);
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < nrec; r += UP_THREADS) {
            const uint8_t* s = s_raw + r * rs;
            if (text) text[r0 + r] = load_le(s, y);
            if (suff) suff[r0 + r] = load_le(s + y, z);
            if (lcp) lcp[r0 + r] = load_le(s + y + z, x);
            if (bwt) bwt[r0 + r] = s[y + z + x];
        }
        __syncthreads();
    }
}

cudaError_t launch_unpack_gesa(const uint8_t* d_rec, uint64_t count, int x, int y, int z, uint32_t* lcp, uint32_t* text,
                               uint32_t* suff, uint8_t* bwt, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    const int rs = x + y + z + 1;
    const size_t smem = size_t(UP_RECS) * rs + 16;
    cudaError_t e = cudaFuncSetAttribute(k_unpack_gesa, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    uint64_t n_tiles = (count + UP_RECS - 1) / UP_RECS;
    unsigned grid = unsigned(n_tiles < 148 * 8 ? n_tiles : 148 * 8);
    k_unpack_gesa<<<grid, UP_THREADS, smem, stream>>>(d_rec, count, x, y, z, lcp, text, suff, bwt);
    return cudaGetLastError();
}

// The record "read" after EOF.  read_el's temporaries share one 8-byte stack slot and a
// failed read leaves it as the last valid record left it: byte 0 = bwt[n-1] (read last); byte b >= 1 = byte b of
// the last-read field wider than b: lcp, then suff, then text for the EGSA record order (text suff lcp), lcp, then
// text, then suff for the BCR order (suff text lcp).  The phantom field of width w is the low w bytes of the slot
// (measured against the reference for ten width combinations in both formats: tests/golden, phantom_tail cases).
__global__ void k_fill_phantom(uint32_t* lcp, uint32_t* text, uint32_t* suff, uint8_t* bwt, uint64_t n_local,
                               uint64_t count, int x, int y, int z, int bcr) {
    const uint8_t b0 = bwt[n_local - 1];
    const uint32_t l = lcp[n_local - 1], t = text[n_local - 1], sf = suff[n_local - 1];
    uint32_t slot = b0;  // bytes 4..7 never reach a 32-bit field
    for (int b = 1; b < 4; ++b) {
        uint32_t v = 0;
        if (b < x) v = l;
        else if (!bcr && b < z) v = sf;
        else if (b < y) v = t;
        else if (bcr && b < z) v = sf;
        slot |= v & (0xffu << (8 * b));
    }
    const uint32_t px = x >= 4 ? slot : slot & ((1u << (8 * x)) - 1u);
    const uint32_t py = y >= 4 ? slot : slot & ((1u << (8 * y)) - 1u);
    const uint32_t pz = z >= 4 ? slot : slot & ((1u << (8 * z)) - 1u);
    for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += uint64_t(gridDim.x) * blockDim.x) {
        lcp[n_local + i] = px;
        text[n_local + i] = py;
        suff[n_local + i] = pz;
        bwt[n_local + i] = b0;
    }
}

cudaError_t launch_fill_phantom(uint32_t* lcp, uint32_t* text, uint32_t* suff, uint8_t* bwt, uint64_t n_local,
                                uint64_t count, int x, int y, int z, int bcr, cudaStream_t stream) {
    k_fill_phantom<<<1, 256, 0, stream>>>(lcp, text, suff, bwt, n_local, count, x, y, z, bcr);
    return cudaGetLastError();
}

}  // namespace e2s
