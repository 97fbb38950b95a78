// Phase 1 (ebwt2clust): LCP boundary stencil + decoupled look-back scan + stream compaction.
//
// Replaces the sequential state machine of cluster_lm / append_entry (ref:ebwt2clust.cpp:54-139)
// by its local-stencil form:
//   ge(i)    = lcp[i] >= k
//   END(i)   = ge(i) & ( (lcp[i-1] > lcp[i] & lcp[i] <= lcp[i+1]) | !ge(i+1) )        1 <= i <= n-2
//   START(i) = ge(i) & ( !ge(i-1) | END(i-1) )
//   fix-ups  : END(0) = 0;  END(1) = 1 if ge(0) & !ge(1)   (ref:ebwt2clust.cpp:83-84: a cluster opened
//              at 0 is still "open" at 1);  END(n-1) and position n are left to the host tail rule
//              (e2s_cluster_merge), because they depend on the post-EOF phantom record.
//   the j-th START pairs with the j-th END; record = (START, (END-START+1) mod 2^16), kept iff >= min_len.
//
// One persistent CTA per resident slot walks tiles round-robin.  Per tile:
//   * the LCP tile (256 rows x V u32) is brought to shared memory by TMA (cp.async.bulk.tensor.2d,
//     hardware 64B/128B swizzle so that the blocked per-thread 128-bit reads are bank-conflict free),
//     STAGES tiles in flight per CTA behind mbarriers;
//   * every thread builds START/END bit masks of its V consecutive positions from registers;
//   * look-back #1 propagates the "cluster still open, started at s" state across tiles
//     (a tile that contains any event publishes its inclusive state at once, so chains are short);
//   * look-back #2 is the classic exclusive sum of kept-record counts giving the output offset;
//   * records are written compacted, in position order.
// Algorithmic traffic: 4 B/position read + 10 B/record written.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"

namespace e2s {

constexpr int CL_THREADS = 256;
constexpr int CL_WARPS = CL_THREADS / 32;

// payload encoding of the open-cluster state
constexpr uint64_t OPEN_NONE = 0;
constexpr uint64_t OPEN_UNKNOWN = 1;  // open, but started before this shard
constexpr uint64_t OPEN_BIAS = 2;     // payload = global start + 2

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int V, int STAGES>
__global__ void __launch_bounds__(CL_THREADS) k_cluster(const __grid_constant__ CUtensorMap tmap, ClusterParams p) {
    constexpr int T = CL_THREADS * V;
    constexpr int CH = V / 4;  // 16-byte chunks per thread row
    constexpr uint32_t FULL = 0xffffffffu;
    static_assert(V == 16 || V == 32, "row = 64B or 128B swizzle atom");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint32_t s_wstate[CL_WARPS];
    __shared__ uint32_t s_wcnt[CL_WARPS];
    __shared__ uint64_t s_x;       // incoming open state of the tile (payload encoding)
    __shared__ uint64_t s_prefix;  // records kept before this tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = p.k;
    const uint32_t xr = (V == 16) ? ((tid >> 1) & 3) : (tid & 7);               // swizzle of my row
    const uint32_t xr_prev = (V == 16) ? (((tid - 1) >> 1) & 3) : ((tid - 1) & 7);
    const uint32_t xr_next = (V == 16) ? (((tid + 1) >> 1) & 3) : ((tid + 1) & 7);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint32_t num_tiles = p.num_tiles;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            uint64_t t = uint64_t(blockIdx.x) + uint64_t(s) * gridDim.x;
            if (t < num_tiles) {
                mbar_expect_tx(&full_bar[s], T * 4);
                tma_load_2d(tiles + size_t(s) * T * 4, &tmap, 0, int(t * CL_THREADS), &full_bar[s]);
            }
        }
    }

    unsigned long long acc_end = 0;
    uint32_t acc_any = 0;

    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        const uint64_t tile_base = t * T;                      // local position of the tile
        const uint64_t my_base = tile_base + uint64_t(tid) * V;  // local position of my first element
        const uint8_t* tile = tiles + size_t(stage) * T * 4;

        // halo values that live outside the tile: fetch from global early
        uint32_t g_m2 = 0, g_m1 = 0, g_p = 0;
        if (tid == 0) {
            uint2 h = *reinterpret_cast<const uint2*>(p.lcp + (int64_t(tile_base) - 2));
            g_m2 = h.x;
            g_m1 = h.y;
        }
        if (tid == CL_THREADS - 1) g_p = p.lcp[tile_base + T];

        mbar_wait(&full_bar[stage], parity);

        uint32_t v[V];
        {
            const uint8_t* row = tile + size_t(tid) * V * 4;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                uint4 c = lds128(row + ((uint32_t(j) ^ xr) << 4));
                v[4 * j + 0] = c.x;
                v[4 * j + 1] = c.y;
                v[4 * j + 2] = c.z;
                v[4 * j + 3] = c.w;
            }
        }
        uint32_t v_m2 = __shfl_up_sync(FULL, v[V - 2], 1);
        uint32_t v_m1 = __shfl_up_sync(FULL, v[V - 1], 1);
        uint32_t v_p = __shfl_down_sync(FULL, v[0], 1);
        if (lane == 0) {
            if (tid == 0) {
                v_m2 = g_m2;
                v_m1 = g_m1;
            } else {
                const uint8_t* prow = tile + size_t(tid - 1) * V * 4 + ((uint32_t(CH - 1) ^ xr_prev) << 4);
                uint2 h = *reinterpret_cast<const uint2*>(prow + 8);
                v_m2 = h.x;
                v_m1 = h.y;
            }
        }
        if (lane == 31) {
            if (tid == CL_THREADS - 1) v_p = g_p;
            else v_p = *reinterpret_cast<const uint32_t*>(tile + size_t(tid + 1) * V * 4 + ((0u ^ xr_next) << 4));
        }

        // ---- flags --------------------------------------------------------------------------
        uint32_t G = 0, A = 0;  // G_j = ge(j); A_j = lcp[j-1] > lcp[j]
#pragma unroll
        for (int j = 0; j < V; ++j) {
            G |= uint32_t(v[j] >= k) << j;
            A |= uint32_t((j == 0 ? v_m1 : v[j - 1]) > v[j]) << j;
        }
        const uint32_t g_m1b = v_m1 >= k, g_pb = v_p >= k;
        const uint32_t a_V = v[V - 1] > v_p;
        const uint32_t Gn = (G >> 1) | (g_pb << (V - 1));   // ge(j+1)
        const uint32_t An = (A >> 1) | (a_V << (V - 1));    // lcp[j] > lcp[j+1]
        uint32_t E = G & ((A & ~An) | ~Gn);
        uint32_t e_prev = g_m1b & ((uint32_t(v_m2 > v_m1) & ((~A) & 1u)) | ((~G) & 1u));

        const uint64_t gpos = p.global_off + my_base;  // global position of my first element
        if (gpos == 0) {  // the init special cases of ref:ebwt2clust.cpp:83-86
            E &= ~1u;
            e_prev = 0;
            if ((G & 1u) && !(G & 2u)) E |= 2u;
        }
        // positions that exist, and END(n_global-1) which belongs to the tail rule
        uint32_t vm;
        {
            int64_t nvalid = int64_t(p.n_local) - int64_t(my_base);
            vm = nvalid >= V ? ((V == 32) ? FULL : ((1u << V) - 1u)) : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
            int64_t last = int64_t(p.n_global) - 1 - int64_t(gpos);
            if (last >= 0 && last < V) E &= ~(1u << last);
        }
        E &= vm;
        const uint32_t Gp = (G << 1) | g_m1b;
        const uint32_t Ep = (E << 1) | e_prev;
        const uint32_t S = G & (~Gp | Ep) & vm;

        // ---- open-state scan (who is the START of an END that has none before it in my range) ---
        const uint32_t ev = E | S;
        uint32_t w = 0;
        if (ev) {
            int hs = 31 - __clz(S), he = 31 - __clz(E);  // -1 when empty
            uint32_t open = (hs > he) ? uint32_t(tid * V + hs + 1) : 0u;
            w = 0x80000000u | open;
        }
        acc_end += __popc(E);
        acc_any |= ev;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, wi, d);
            if (lane >= d && !(wi >> 31)) wi = o;
        }
        uint32_t wx = __shfl_up_sync(FULL, wi, 1);
        if (lane == 0) wx = 0;
        if (lane == 31) s_wstate[warp] = wi;
        __syncthreads();  // (1) all threads are done reading the stage; warp states visible

        if (tid == 0) {  // refill this stage for the tile STAGES rounds ahead
            uint64_t tn = t + uint64_t(STAGES) * gridDim.x;
            if (tn < num_tiles) {
                mbar_expect_tx(&full_bar[stage], T * 4);
                tma_load_2d(tiles + size_t(stage) * T * 4, &tmap, 0, int(tn * CL_THREADS), &full_bar[stage]);
            }
        }

        // my block-exclusive state: nearest earlier thread in the tile that has an event
        uint32_t bx = wx;
#pragma unroll
        for (int q = CL_WARPS - 1; q >= 0; --q) {
            uint32_t ws = s_wstate[q];
            if (q < warp && !(bx >> 31)) bx = ws;
        }

        if (tid == 0) {  // look-back #1: incoming open state
            uint32_t tot = 0;
#pragma unroll
            for (int q = 0; q < CL_WARPS; ++q) {
                uint32_t ws = s_wstate[q];
                if (ws >> 31) tot = ws;
            }
            const bool has = tot >> 31;
            if (has) {
                uint32_t op = tot & 0x7fffffffu;
                desc_store(&p.desc_state[t], ST_INCLUSIVE, op ? (p.global_off + tile_base + (op - 1) + OPEN_BIAS) : OPEN_NONE);
            } else {
                desc_store(&p.desc_state[t], ST_AGGREGATE, 0);
            }
            uint64_t X = p.global_off == 0 ? OPEN_NONE : OPEN_UNKNOWN;
            for (int64_t j = int64_t(t) - 1; j >= 0; --j) {
                uint64_t d;
                while (((d = desc_load(&p.desc_state[j])) >> ST_SHIFT) == ST_INVALID) __nanosleep(32);
                if ((d >> ST_SHIFT) == ST_INCLUSIVE) {
                    X = d & ST_PAYLOAD;
                    break;
                }
            }
            if (!has) desc_store(&p.desc_state[t], ST_INCLUSIVE, X);
            s_x = X;
            if (t == num_tiles - 1) {  // state after the whole shard
                uint64_t fin = has ? ((tot & 0x7fffffffu) ? (p.global_off + tile_base + ((tot & 0x7fffffffu) - 1) + 1) : 0)
                                   : (X >= OPEN_BIAS ? X - OPEN_BIAS + 1 : 0);
                p.res->open_start = fin;
            }
        }
        __syncthreads();  // (2) X visible

        // ---- count the records I keep ----------------------------------------------------------
        uint64_t cur;  // payload encoding
        if (bx >> 31) {
            uint32_t op = bx & 0x7fffffffu;
            cur = op ? (p.global_off + tile_base + (op - 1) + OPEN_BIAS) : OPEN_NONE;
        } else {
            cur = s_x;
        }
        uint32_t cnt = 0;
        {
            uint32_t m = E;
            while (m) {
                int e = __ffs(m) - 1;
                m &= m - 1;
                uint32_t below = S & (uint32_t(2u << e) - 1u);
                uint64_t st = below ? (gpos + (31 - __clz(below)) + OPEN_BIAS) : cur;
                if (st >= OPEN_BIAS) {
                    uint32_t len = uint32_t((gpos + e) - (st - OPEN_BIAS) + 1) & 0xffffu;
                    cnt += (int(len) >= p.min_len);
                }
            }
        }
        uint32_t ci = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, ci, d);
            if (lane >= d) ci += o;
        }
        if (lane == 31) s_wcnt[warp] = ci;
        __syncthreads();  // (3) warp counts visible
        uint32_t off = ci - cnt;
#pragma unroll
        for (int q = 0; q < CL_WARPS; ++q) {
            uint32_t wc = s_wcnt[q];
            if (q < warp) off += wc;
        }
        if (tid == 0) {  // look-back #2: exclusive sum of kept records
            uint32_t tot = 0;
#pragma unroll
            for (int q = 0; q < CL_WARPS; ++q) tot += s_wcnt[q];
            uint64_t prefix = 0;
            if (t > 0) {
                desc_store(&p.desc_cnt[t], ST_AGGREGATE, tot);
                for (int64_t j = int64_t(t) - 1; j >= 0; --j) {
                    uint64_t d;
                    while (((d = desc_load(&p.desc_cnt[j])) >> ST_SHIFT) == ST_INVALID) __nanosleep(32);
                    prefix += d & ST_PAYLOAD;
                    if ((d >> ST_SHIFT) == ST_INCLUSIVE) break;
                }
            }
            desc_store(&p.desc_cnt[t], ST_INCLUSIVE, prefix + tot);
            s_prefix = prefix;
            if (t == num_tiles - 1) p.res->n_written = prefix + tot;
        }
        __syncthreads();  // (4) prefix visible

        // ---- write my records -----------------------------------------------------------------
        if (E) {
            uint64_t o = s_prefix + off;
            uint32_t m = E;
            while (m) {
                int e = __ffs(m) - 1;
                m &= m - 1;
                uint32_t below = S & (uint32_t(2u << e) - 1u);
                uint64_t st = below ? (gpos + (31 - __clz(below)) + OPEN_BIAS) : cur;
                const uint64_t ge_pos = gpos + e;
                if (st >= OPEN_BIAS) {
                    st -= OPEN_BIAS;
                    uint32_t len = uint32_t(ge_pos - st + 1) & 0xffffu;
                    if (int(len) >= p.min_len) {
                        if (o < p.cap) {
                            p.out_start[o] = st;
                            p.out_len[o] = uint16_t(len);
                        } else {
                            p.res->overflow = 1;
                        }
                        ++o;
                    }
                    if (ge_pos + 2 == p.n_global) p.res->end_nm2_start = st + 1;
                } else {  // OPEN_UNKNOWN: the shard's head END (OPEN_NONE cannot happen)
                    p.res->head_end = ge_pos + 1;
                    if (ge_pos + 2 == p.n_global) p.res->end_nm2_start = ~0ull;
                }
            }
        }
    }

    // ---- per-shard totals ------------------------------------------------------------------------
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc_end += __shfl_xor_sync(FULL, acc_end, d);
        acc_any |= __shfl_xor_sync(FULL, acc_any, d);
    }
    if (lane == 0) {
        if (acc_end) atomicAdd(&p.res->n_end, acc_end);
        if (acc_any) atomicOr(&p.res->any_event, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

template <int V, int STAGES>
static cudaError_t launch_cluster_t(const ClusterParams& p, uint64_t rows_alloc, int sm_count, cudaStream_t stream,
                                    int* grid_out) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {cuuint64_t(V), cuuint64_t(rows_alloc)};
    cuuint64_t gstride[1] = {cuuint64_t(V) * 4};
    cuuint32_t box[2] = {cuuint32_t(V), cuuint32_t(CL_THREADS)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(p.lcp), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, V == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

    constexpr int T = CL_THREADS * V;
    const size_t smem = size_t(STAGES) * T * 4 + 1024;
    auto kern = k_cluster<V, STAGES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CL_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    // persistent grid: every CTA must be resident (tiles spin on their predecessors' descriptors)
    uint64_t grid = uint64_t(sm_count) * occ;
    if (grid > p.num_tiles) grid = p.num_tiles;
    if (grid_out) *grid_out = int(grid);
    kern<<<dim3(unsigned(grid)), dim3(CL_THREADS), smem, stream>>>(tmap, p);
    return cudaGetLastError();
}

cudaError_t launch_cluster(const ClusterParams& p, uint64_t rows_alloc16, int sm_count, cudaStream_t stream,
                           int variant, int* grid_out) {
    // rows_alloc16 = number of 16-element rows available in the allocation starting at p.lcp
    switch (variant) {
        case 1: return launch_cluster_t<32, 3>(p, rows_alloc16 / 2, sm_count, stream, grid_out);
        case 2: return launch_cluster_t<16, 3>(p, rows_alloc16, sm_count, stream, grid_out);
        case 3: return launch_cluster_t<32, 2>(p, rows_alloc16 / 2, sm_count, stream, grid_out);
        default: return launch_cluster_t<16, 4>(p, rows_alloc16, sm_count, stream, grid_out);
    }
}

int cluster_tile_positions(int variant) { return (variant == 1 || variant == 3) ? CL_THREADS * 32 : CL_THREADS * 16; }

}  // namespace e2s
