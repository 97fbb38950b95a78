// Phase 1 (ebwt2clust): K1 = streaming LCP boundary stencil, K2 = decoupled look-back scan + compaction.
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
// K1  k_lcp_flags: pure streaming map LCP -> two bit masks (START, END), one bit per position.
//     Persistent CTAs; each 8192-position tile (256 rows x 32 u32) is brought to shared memory by TMA
//     (cp.async.bulk.tensor.2d, hardware 128B swizzle => the blocked per-thread 128-bit reads are bank
//     conflict free), STAGES tiles in flight per CTA behind mbarriers; every thread owns 32 consecutive
//     positions in registers and writes one 32-bit word per mask (a warp writes 128 contiguous bytes).
//     Traffic: 4 B/position read + 0.25 B/position written.  No inter-CTA dependency.
// K2  k_cluster_emit: decoupled look-back over the bit masks (0.25 B/position read; 262144 positions per
//     tile, so 32x fewer descriptors than K1's tile size would give).  Two chained scans:
//       #1 "cluster still open, started at s" state (a tile containing any event publishes its inclusive
//          state immediately, so chains stop at the nearest tile with an event);
//       #2 exclusive sum of kept-record counts = output offset (classic aggregate/inclusive descriptors,
//          warp-parallel windows of 64 predecessors).
//     Records are written compacted in position order: 10 B per record; their length histogram
//     (statistics(), ref:clust2snp.cpp:899-907) is accumulated on the way.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"

namespace e2s {

constexpr uint32_t FULL = 0xffffffffu;

// =============================================================================================
// K1: LCP -> START / END bit masks
// =============================================================================================
constexpr int FL_THREADS = 256;
constexpr int FL_V = 32;                      // positions per thread = one 128-byte swizzle row
constexpr int FL_T = FL_THREADS * FL_V;       // 8192 positions per tile

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(FL_THREADS) k_lcp_flags(const __grid_constant__ CUtensorMap tmap, FlagParams p) {
    constexpr int V = FL_V, T = FL_T, CH = V / 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[STAGES];

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t k = p.k;
    const uint32_t xr = tid & 7, xr_prev = (tid - 1) & 7, xr_next = (tid + 1) & 7;  // 128B swizzle: chunk ^= row & 7

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
                tma_load_2d(tiles + size_t(s) * T * 4, &tmap, 0, int(t * FL_THREADS), &full_bar[s]);
            }
        }
    }

    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        const uint64_t tile_base = t * T;
        const uint64_t my_base = tile_base + uint64_t(tid) * V;
        const uint8_t* tile = tiles + size_t(stage) * T * 4;

        uint32_t g_m2 = 0, g_m1 = 0, g_p = 0;  // halo values outside the tile: issue the global loads early
        if (tid == 0) {
            uint2 h = *reinterpret_cast<const uint2*>(p.lcp + (int64_t(tile_base) - 2));
            g_m2 = h.x;
            g_m1 = h.y;
        }
        if (tid == FL_THREADS - 1) g_p = p.lcp[tile_base + T];

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
            if (tid == FL_THREADS - 1) v_p = g_p;
            else v_p = *reinterpret_cast<const uint32_t*>(tile + size_t(tid + 1) * V * 4 + ((0u ^ xr_next) << 4));
        }
        __syncthreads();  // every thread has its tile data in registers: the stage can be refilled
        if (tid == 0) {
            uint64_t tn = t + uint64_t(STAGES) * gridDim.x;
            if (tn < num_tiles) {
                mbar_expect_tx(&full_bar[stage], T * 4);
                tma_load_2d(tiles + size_t(stage) * T * 4, &tmap, 0, int(tn * FL_THREADS), &full_bar[stage]);
            }
        }

        uint32_t G = 0, A = 0;  // G_j = ge(j); A_j = lcp[j-1] > lcp[j]
#pragma unroll
        for (int j = 0; j < V; ++j) {
            G |= uint32_t(v[j] >= k) << j;
            A |= uint32_t((j == 0 ? v_m1 : v[j - 1]) > v[j]) << j;
        }
        const uint32_t g_m1b = v_m1 >= k, g_pb = v_p >= k;
        const uint32_t a_V = v[V - 1] > v_p;
        const uint32_t Gn = (G >> 1) | (g_pb << (V - 1));  // ge(j+1)
        const uint32_t An = (A >> 1) | (a_V << (V - 1));   // lcp[j] > lcp[j+1]
        uint32_t E = G & ((A & ~An) | ~Gn);
        uint32_t e_prev = g_m1b & ((uint32_t(v_m2 > v_m1) & ((~A) & 1u)) | ((~G) & 1u));

        const uint64_t gpos = p.global_off + my_base;
        if (gpos == 0) {  // the init special cases of ref:ebwt2clust.cpp:83-86
            E &= ~1u;
            e_prev = 0;
            if ((G & 1u) && !(G & 2u)) E |= 2u;
        }
        uint32_t vm;
        {
            const int64_t nvalid = int64_t(p.n_local) - int64_t(my_base);
            vm = nvalid >= V ? FULL : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
            const int64_t last = int64_t(p.n_global) - 1 - int64_t(gpos);  // END(n_global-1): host tail rule
            if (last >= 0 && last < V) E &= ~(1u << last);
        }
        E &= vm;
        const uint32_t Gp = (G << 1) | g_m1b;
        const uint32_t Ep = (E << 1) | e_prev;
        const uint32_t S = G & (~Gp | Ep) & vm;

        const uint64_t w = t * FL_THREADS + tid;
        p.s_words[w] = S;
        p.e_words[w] = E;
    }
}

// =============================================================================================
// K2: look-back scan over the bit masks + compaction
// =============================================================================================
constexpr int EM_THREADS = 512;
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int EM_WPT = 16;                             // 32-bit words per thread
constexpr int EM_TILE_WORDS = EM_THREADS * EM_WPT;     // 8192 words = 262144 positions

// payload encoding of the open-cluster state
constexpr uint64_t OPEN_NONE = 0;
constexpr uint64_t OPEN_UNKNOWN = 1;  // open, but started before this shard
constexpr uint64_t OPEN_BIAS = 2;     // payload = global start + 2

__device__ __forceinline__ uint64_t desc_wait(const uint64_t* d) {
    uint64_t v;
    while (((v = desc_load(d)) >> ST_SHIFT) == ST_INVALID) __nanosleep(20);
    return v;
}

// nearest predecessor whose state is final (warp-parallel windows of 64 descriptors, 2 per lane)
__device__ __forceinline__ uint64_t lookback_state(const uint64_t* desc, int64_t t, uint64_t init, int lane) {
    int64_t j = t - 1;
    const uint64_t virt = (ST_INCLUSIVE << ST_SHIFT) | init;  // virtual tile -1
    while (true) {
        const int64_t i0 = j - lane, i1 = j - 32 - lane;
        const uint64_t d0 = i0 >= 0 ? desc_wait(desc + i0) : virt;
        const uint32_t inc0 = __ballot_sync(FULL, (d0 >> ST_SHIFT) == ST_INCLUSIVE);
        if (inc0) return __shfl_sync(FULL, d0, __ffs(inc0) - 1) & ST_PAYLOAD;
        const uint64_t d1 = i1 >= 0 ? desc_wait(desc + i1) : virt;
        const uint32_t inc1 = __ballot_sync(FULL, (d1 >> ST_SHIFT) == ST_INCLUSIVE);
        if (inc1) return __shfl_sync(FULL, d1, __ffs(inc1) - 1) & ST_PAYLOAD;
        j -= 64;
    }
}

// exclusive sum of the predecessors' aggregates (windows of 64 descriptors, both loads in flight together)
__device__ __forceinline__ uint64_t lookback_sum(const uint64_t* desc, int64_t t, int lane) {
    uint64_t prefix = 0;
    int64_t j = t - 1;
    const uint64_t virt = ST_INCLUSIVE << ST_SHIFT;  // virtual tile -1: inclusive 0
    while (true) {
        const int64_t i0 = j - lane, i1 = j - 32 - lane;
        uint64_t d0 = virt, d1 = virt;
        if (i0 >= 0) d0 = desc_load(desc + i0);
        if (i1 >= 0) d1 = desc_load(desc + i1);
        if (i0 >= 0 && (d0 >> ST_SHIFT) == ST_INVALID) d0 = desc_wait(desc + i0);
        const uint32_t inc0 = __ballot_sync(FULL, (d0 >> ST_SHIFT) == ST_INCLUSIVE);
        const int f0 = inc0 ? __ffs(inc0) - 1 : 31;
        uint64_t c = lane <= f0 ? (d0 & ST_PAYLOAD) : 0;
        uint32_t inc1 = 0;
        if (!inc0) {  // uniform branch
            if (i1 >= 0 && (d1 >> ST_SHIFT) == ST_INVALID) d1 = desc_wait(desc + i1);
            inc1 = __ballot_sync(FULL, (d1 >> ST_SHIFT) == ST_INCLUSIVE);
            const int f1 = inc1 ? __ffs(inc1) - 1 : 31;
            c += lane <= f1 ? (d1 & ST_PAYLOAD) : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        prefix += c;
        if (inc0 | inc1) return prefix;
        j -= 64;
    }
}

__global__ void __launch_bounds__(EM_THREADS) k_cluster_emit(EmitParams p) {
    __shared__ uint32_t s_wstate[EM_WARPS];
    __shared__ uint32_t s_wcnt[EM_WARPS];
    __shared__ uint64_t s_x;
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_tot;
    __shared__ unsigned int s_hist[E2S_HIST_BINS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < E2S_HIST_BINS; i += EM_THREADS) s_hist[i] = 0;
    __syncthreads();
    unsigned long long acc_end = 0, acc_bases = 0;
    uint32_t acc_any = 0;

    for (uint64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const uint64_t word0 = t * EM_TILE_WORDS + uint64_t(tid) * EM_WPT;
        uint32_t S[EM_WPT], E[EM_WPT];
        {
            const uint4* ps = reinterpret_cast<const uint4*>(p.s_words + word0);
            const uint4* pe = reinterpret_cast<const uint4*>(p.e_words + word0);
#pragma unroll
            for (int q = 0; q < EM_WPT / 4; ++q) {
                uint4 a = __ldg(ps + q), b = __ldg(pe + q);
                S[4 * q] = a.x; S[4 * q + 1] = a.y; S[4 * q + 2] = a.z; S[4 * q + 3] = a.w;
                E[4 * q] = b.x; E[4 * q + 1] = b.y; E[4 * q + 2] = b.z; E[4 * q + 3] = b.w;
            }
        }
        const uint64_t tile_lbase = t * uint64_t(EM_TILE_WORDS) * 32;      // local position of the tile
        const uint64_t gbase = p.global_off + tile_lbase + uint64_t(tid) * EM_WPT * 32;  // global position of my first bit

        // ---- my state: is a cluster open after my range, and where did it start ------------------
        uint32_t w = 0;
        uint32_t n_e = 0;
#pragma unroll
        for (int j = 0; j < EM_WPT; ++j) {
            n_e += __popc(E[j]);
            if (S[j] | E[j]) {
                const int hs = 31 - __clz(S[j]), he = 31 - __clz(E[j]);
                w = 0x80000000u | (hs > he ? uint32_t((tid * EM_WPT + j) * 32 + hs + 1) : 0u);
            }
        }
        acc_end += n_e;
        acc_any |= w;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, wi, d);
            if (lane >= d && !(wi >> 31)) wi = o;
        }
        uint32_t wx = __shfl_up_sync(FULL, wi, 1);
        if (lane == 0) wx = 0;
        if (lane == 31) s_wstate[warp] = wi;
        __syncthreads();  // (1)

        uint32_t bx = wx;  // nearest earlier thread of the tile with an event
#pragma unroll
        for (int q = EM_WARPS - 1; q >= 0; --q) {
            uint32_t ws = s_wstate[q];
            if (q < warp && !(bx >> 31)) bx = ws;
        }
        if (warp == 0) {  // look-back #1: the open state entering the tile
            uint32_t tot = 0;
#pragma unroll
            for (int q = 0; q < EM_WARPS; ++q) {
                uint32_t ws = s_wstate[q];
                if (ws >> 31) tot = ws;
            }
            const bool has = tot >> 31;
            const uint32_t op = tot & 0x7fffffffu;
            if (lane == 0) {
                if (has) desc_store(&p.desc_state[t], ST_INCLUSIVE, op ? (p.global_off + tile_lbase + (op - 1) + OPEN_BIAS) : OPEN_NONE);
                else desc_store(&p.desc_state[t], ST_AGGREGATE, 0);
            }
            const uint64_t X = lookback_state(p.desc_state, int64_t(t), p.global_off == 0 ? OPEN_NONE : OPEN_UNKNOWN, lane);
            if (lane == 0) {
                if (!has) desc_store(&p.desc_state[t], ST_INCLUSIVE, X);
                s_x = X;
                if (t == p.num_tiles - 1)  // state after the whole shard
                    p.res->open_start = has ? (op ? (p.global_off + tile_lbase + (op - 1) + 1) : 0)
                                            : (X >= OPEN_BIAS ? X - OPEN_BIAS + 1 : 0);
            }
        }
        __syncthreads();  // (2)

        uint64_t cur0;  // open state entering my range (payload encoding)
        if (bx >> 31) {
            const uint32_t op = bx & 0x7fffffffu;
            cur0 = op ? (p.global_off + tile_lbase + (op - 1) + OPEN_BIAS) : OPEN_NONE;
        } else {
            cur0 = s_x;
        }

        // ---- count the records I keep --------------------------------------------------------------
        uint32_t cnt = 0;
        if (n_e) {
            uint64_t cur = cur0;
#pragma unroll
            for (int j = 0; j < EM_WPT; ++j) {
                const uint64_t gp = gbase + uint64_t(j) * 32;
                uint32_t m = E[j];
                while (m) {
                    const int e = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t below = S[j] & (uint32_t(2u << e) - 1u);
                    const uint64_t st = below ? (gp + (31 - __clz(below)) + OPEN_BIAS) : cur;
                    if (st >= OPEN_BIAS) {
                        const uint32_t len = uint32_t((gp + e) - (st - OPEN_BIAS) + 1) & 0xffffu;
                        cnt += (int(len) >= p.min_len);
                    }
                }
                if (S[j] | E[j]) {
                    const int hs = 31 - __clz(S[j]), he = 31 - __clz(E[j]);
                    cur = hs > he ? (gp + hs + OPEN_BIAS) : OPEN_NONE;
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
        __syncthreads();  // (3)
        uint32_t off = ci - cnt;
#pragma unroll
        for (int q = 0; q < EM_WARPS; ++q) {
            uint32_t wc = s_wcnt[q];
            if (q < warp) off += wc;
        }
        if (warp == 0) {  // look-back #2: records kept before this tile
            uint32_t tot = 0;
#pragma unroll
            for (int q = 0; q < EM_WARPS; ++q) tot += s_wcnt[q];
            if (lane == 0 && t > 0) desc_store(&p.desc_cnt[t], ST_AGGREGATE, tot);
            const uint64_t prefix = t > 0 ? lookback_sum(p.desc_cnt, int64_t(t), lane) : 0;
            if (lane == 0) {
                desc_store(&p.desc_cnt[t], ST_INCLUSIVE, prefix + tot);
                s_prefix = prefix;
                s_tot = tot;
                if (t == p.num_tiles - 1) p.res->n_written = prefix + tot;
            }
        }
        __syncthreads();  // (4)

        // ---- write my records ------------------------------------------------------------------------
        if (n_e) {
            uint64_t o = s_prefix + off;
            uint64_t cur = cur0;
            uint32_t last_len = 0;
#pragma unroll
            for (int j = 0; j < EM_WPT; ++j) {
                const uint64_t gp = gbase + uint64_t(j) * 32;
                uint32_t m = E[j];
                while (m) {
                    const int e = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t below = S[j] & (uint32_t(2u << e) - 1u);
                    uint64_t st = below ? (gp + (31 - __clz(below)) + OPEN_BIAS) : cur;
                    const uint64_t ge_pos = gp + e;
                    if (st >= OPEN_BIAS) {
                        st -= OPEN_BIAS;
                        const uint32_t len = uint32_t(ge_pos - st + 1) & 0xffffu;
                        if (int(len) >= p.min_len) {
                            if (o < p.cap) {
                                p.out_start[o] = st;
                                p.out_len[o] = uint16_t(len);
                            } else {
                                p.res->overflow = 1;
                            }
                            ++o;
                            last_len = len;
                            acc_bases += len;
                            if (len <= MAX_C_LEN) atomicAdd(&s_hist[len], 1u);
                        }
                        if (ge_pos + 2 == p.n_global) p.res->end_nm2_start = st + 1;
                    } else {  // OPEN_UNKNOWN: the shard's head END (OPEN_NONE cannot happen)
                        p.res->head_end = ge_pos + 1;
                        if (ge_pos + 2 == p.n_global) p.res->end_nm2_start = ~0ull;
                    }
                }
                if (S[j] | E[j]) {
                    const int hs = 31 - __clz(S[j]), he = 31 - __clz(E[j]);
                    cur = hs > he ? (gp + hs + OPEN_BIAS) : OPEN_NONE;
                }
            }
            // the thread holding the tile's last kept record reports it (o = its index + 1)
            if (cnt && o == s_prefix + s_tot) atomicMax(&p.res->last_rec, (unsigned long long)((o << 16) | last_len));
        }
    }

#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc_end += __shfl_xor_sync(FULL, acc_end, d);
        acc_bases += __shfl_xor_sync(FULL, acc_bases, d);
        acc_any |= __shfl_xor_sync(FULL, acc_any, d);
    }
    if (lane == 0) {
        if (acc_end) atomicAdd(&p.res->n_end, acc_end);
        if (acc_bases) atomicAdd(&p.res->n_bases, acc_bases);
        if (acc_any) atomicOr(&p.res->any_event, 1ull);
    }
    __syncthreads();
    for (int i = tid; i < E2S_HIST_BINS; i += EM_THREADS)
        if (s_hist[i]) atomicAdd(&p.res->hist[i], (unsigned long long)s_hist[i]);
    if (blockIdx.x == 0 && tid == 0 && p.tail_lcp) {
        p.res->tail_lcp_nm2 = p.tail_lcp[0];
        p.res->tail_lcp_nm1 = p.tail_lcp[1];
        p.res->tail_bwt_nm1 = p.tail_bwt[0];
    }
}

// ---- tiny list helpers -----------------------------------------------------------------------
struct PutArgs {
    uint64_t st[3];
    uint64_t ln[3];
    int n;
};
__global__ void k_put_records(uint64_t* d_start, uint16_t* d_len, uint64_t at, PutArgs a) {
    if (int(threadIdx.x) < a.n) {
        d_start[at + threadIdx.x] = a.st[threadIdx.x];
        d_len[at + threadIdx.x] = uint16_t(a.ln[threadIdx.x]);
    }
}
cudaError_t launch_put_records(uint64_t* d_start, uint16_t* d_len, uint64_t at, const uint64_t* st, const uint64_t* ln,
                               int n, cudaStream_t stream) {
    PutArgs a;
    a.n = n;
    for (int i = 0; i < 3; ++i) {
        a.st[i] = i < n ? st[i] : 0;
        a.ln[i] = i < n ? ln[i] : 0;
    }
    k_put_records<<<1, 32, 0, stream>>>(d_start, d_len, at, a);
    return cudaGetLastError();
}

// .clusters file layout: u64 start LE + u16 length LE = 10 bytes, no padding (ref:ebwt2clust.cpp:58-59)
__global__ void k_pack_records(const uint64_t* __restrict__ st, const uint16_t* __restrict__ ln, uint64_t m,
                               uint16_t* __restrict__ out) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t s = st[i];
        uint16_t* o = out + i * 5;
        o[0] = uint16_t(s);
        o[1] = uint16_t(s >> 16);
        o[2] = uint16_t(s >> 32);
        o[3] = uint16_t(s >> 48);
        o[4] = ln[i];
    }
}
cudaError_t launch_pack_records(const uint64_t* d_start, const uint16_t* d_len, uint64_t m, uint8_t* d_out,
                                cudaStream_t stream, int sm_count) {
    if (m == 0) return cudaSuccess;
    uint64_t blocks = (m + 255) / 256;
    if (blocks > uint64_t(sm_count) * 16) blocks = uint64_t(sm_count) * 16;
    k_pack_records<<<unsigned(blocks), 256, 0, stream>>>(d_start, d_len, m, reinterpret_cast<uint16_t*>(d_out));
    return cudaGetLastError();
}

// =============================================================================================
// host side
// =============================================================================================
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

uint64_t flags_words_needed(uint64_t n_local) {
    // K1 writes whole 8192-position tiles; K2 reads whole 65536-position tiles
    const uint64_t w = (n_local + 31) / 32;
    return (w + EM_TILE_WORDS - 1) / EM_TILE_WORDS * EM_TILE_WORDS;
}

uint64_t emit_num_tiles(uint64_t n_local) { return flags_words_needed(n_local) / EM_TILE_WORDS; }

template <int STAGES>
static cudaError_t launch_flags_t(const FlagParams& p0, uint64_t rows_alloc32, int sm_count, cudaStream_t stream) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return cudaErrorNotSupported;
    FlagParams p = p0;
    p.num_tiles = uint32_t((p.n_local + FL_T - 1) / FL_T);
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {cuuint64_t(FL_V), cuuint64_t(rows_alloc32)};
    cuuint64_t gstride[1] = {cuuint64_t(FL_V) * 4};
    cuuint32_t box[2] = {cuuint32_t(FL_V), cuuint32_t(FL_THREADS)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(p.lcp), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const size_t smem = size_t(STAGES) * FL_T * 4 + 1024;
    auto kern = k_lcp_flags<STAGES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FL_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    uint64_t grid = uint64_t(sm_count) * occ;
    if (grid > p.num_tiles) grid = p.num_tiles;
    kern<<<dim3(unsigned(grid)), dim3(FL_THREADS), smem, stream>>>(tmap, p);
    return cudaGetLastError();
}

cudaError_t launch_flags(const FlagParams& p, uint64_t rows_alloc32, int sm_count, cudaStream_t stream, int variant) {
    switch (variant) {
        case 1: return launch_flags_t<2>(p, rows_alloc32, sm_count, stream);
        case 2: return launch_flags_t<4>(p, rows_alloc32, sm_count, stream);
        default: return launch_flags_t<3>(p, rows_alloc32, sm_count, stream);
    }
}

cudaError_t launch_emit(const EmitParams& p0, int sm_count, cudaStream_t stream) {
    EmitParams p = p0;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cluster_emit, EM_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (occ > 1) occ = 1;  // persistent and fully resident: tiles spin on their predecessors' descriptors
    uint64_t grid = uint64_t(sm_count) * occ;
    if (grid > p.num_tiles) grid = p.num_tiles;
    k_cluster_emit<<<dim3(unsigned(grid)), dim3(EM_THREADS), 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace e2s
