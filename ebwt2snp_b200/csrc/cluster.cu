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
// START and END bits alternate (S <= E < S' <= E' ...), so inside a tile the j-th END pairs with the
// (j - open_in)-th START, where open_in says whether a cluster is still open when the tile begins (its
// START then comes from the look-back).  The kernel therefore never walks bits cluster by cluster:
//   1. every thread loads 4+4 mask words (128 positions), popcounts them, and a block scan gives each
//      thread the ordinal of its first START / END;
//   2. the bit positions are scattered by ordinal into two shared lists (16-bit tile-local positions);
//   3. a dense, END-centric pass (one lane per record, all lanes busy) forms (start, len), applies the
//      min_len test on the wrapped length, ranks the kept records with ballot/popc and writes them
//      coalesced at the offset the second look-back provides.
// Tiles with more than EM_CAP ENDs (adversarial inputs only) take the same steps in windows of EM_CAP.
constexpr int EM_THREADS = 256;
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int EM_WPT = 4;                               // 32-bit words per thread and mask (one uint4)
constexpr int EM_TILE_WORDS = EM_THREADS * EM_WPT;      // 1024 words = 32768 positions
constexpr int EM_TILE_POS = EM_TILE_WORDS * 32;
constexpr int EM_CAP = 2048;                            // ENDs per window
constexpr int EM_EPT = EM_CAP / EM_THREADS;             // dense entries per thread and window
constexpr int EM_ROWS = EM_CAP / 32;

// payload encoding of the open-cluster state
constexpr uint64_t OPEN_NONE = 0;
constexpr uint64_t OPEN_UNKNOWN = 1;  // open, but started before this shard
constexpr uint64_t OPEN_BIAS = 2;     // payload = global start + 2

// Look-back windows: every lane keeps LB_R descriptor loads in flight, so one round trip to L2 covers
// 32 * LB_R predecessors.  (The tile rate of the whole grid is bounded by window size / window latency x
// tiles in flight: with 64-wide windows the scan was latency bound, profiles/r1_v4_emit.)  A lane only
// polls a descriptor that is still INVALID when it lies before the nearest INCLUSIVE one.
constexpr int LB_R = 4;

__device__ __forceinline__ uint32_t desc_status(uint64_t d) { return uint32_t(d >> ST_SHIFT); }

// nearest predecessor whose open-cluster state is final
__device__ __forceinline__ uint64_t lookback_state(const uint64_t* desc, int64_t t, uint64_t init, int lane) {
    int64_t j = t - 1;
    const uint64_t virt = (ST_INCLUSIVE << ST_SHIFT) | init;  // virtual tile -1
    while (true) {
        uint64_t d[LB_R];
#pragma unroll
        for (int r = 0; r < LB_R; ++r) {
            const int64_t i = j - 32 * r - lane;
            d[r] = i >= 0 ? desc_load(desc + i) : virt;
        }
#pragma unroll
        for (int r = 0; r < LB_R; ++r) {
            const int64_t i = j - 32 * r - lane;
            while (true) {
                const uint32_t inc = __ballot_sync(FULL, desc_status(d[r]) == ST_INCLUSIVE);
                const uint32_t inv = __ballot_sync(FULL, desc_status(d[r]) == ST_INVALID);
                const uint32_t before = inc ? ((1u << (__ffs(inc) - 1)) - 1u) : FULL;  // lanes nearer than the first INCLUSIVE
                if (!(inv & before)) {
                    if (inc) return __shfl_sync(FULL, d[r], __ffs(inc) - 1) & ST_PAYLOAD;
                    break;  // 32 pass-through tiles: next round
                }
                if (desc_status(d[r]) == ST_INVALID && ((before >> lane) & 1u)) {
                    __nanosleep(20);
                    d[r] = desc_load(desc + i);
                }
            }
        }
        j -= 32 * LB_R;
    }
}

// exclusive sum of the predecessors' aggregates back to the nearest INCLUSIVE descriptor
__device__ __forceinline__ uint64_t lookback_sum(const uint64_t* desc, int64_t t, int lane) {
    uint64_t acc = 0;  // per-lane partial sum, reduced once at the end
    int64_t j = t - 1;
    const uint64_t virt = ST_INCLUSIVE << ST_SHIFT;  // virtual tile -1: inclusive 0
    bool done = false;
    while (!done) {
        uint64_t d[LB_R];
#pragma unroll
        for (int r = 0; r < LB_R; ++r) {
            const int64_t i = j - 32 * r - lane;
            d[r] = i >= 0 ? desc_load(desc + i) : virt;
        }
#pragma unroll
        for (int r = 0; r < LB_R; ++r) {
            if (done) break;
            const int64_t i = j - 32 * r - lane;
            while (true) {
                const uint32_t inc = __ballot_sync(FULL, desc_status(d[r]) == ST_INCLUSIVE);
                const uint32_t inv = __ballot_sync(FULL, desc_status(d[r]) == ST_INVALID);
                const int f = inc ? __ffs(inc) - 1 : 32;
                const uint32_t upto = f >= 31 ? FULL : ((2u << f) - 1u);  // lanes 0..f (f = 32: all)
                if (!(inv & upto)) {
                    if ((upto >> lane) & 1u) acc += d[r] & ST_PAYLOAD;
                    done = inc != 0;
                    break;
                }
                if (desc_status(d[r]) == ST_INVALID && ((upto >> lane) & 1u)) {
                    __nanosleep(20);
                    d[r] = desc_load(desc + i);
                }
            }
        }
        j -= 32 * LB_R;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    return acc;
}

struct EmitShared {
    uint16_t spos[EM_CAP + 2];  // spos[1 + q] = tile-local position of START ordinal (win + q); spos[0]: ordinal win - 1
    uint16_t epos[EM_CAP];      // epos[q]     = tile-local position of END   ordinal (win + q)
    uint16_t erank[EM_CAP];     // FAST: rank of that END among the tile's kept records, 0xFFFF = dropped (shorter than min_len)
    uint32_t wsum[EM_WARPS];    // per warp: #START | #END << 16
    uint32_t wlast[EM_WARPS];   // per warp: 0 = no event, else 0x80000000 | (1 + tile-local position of a START left open)
    uint32_t wfirst[EM_WARPS];  // per warp: 0 = no event, else 0x80000000 | (first event is an END without START)
    uint32_t wfirst_end[EM_WARPS];  // per warp: tile-local position of its first END, ~0 = none
    uint32_t wdrop[EM_WARPS];   // per warp: ENDs whose cluster is shorter than min_len (bit-parallel count)
    uint32_t rowcnt[EM_ROWS];
    uint32_t rowbase[EM_ROWS];
    uint32_t round_total;
    uint32_t adj;               // FAST: 1 if the tile's first END (cluster carried in) turned out not to be kept
    unsigned long long tile;    // ticket of the tile being processed
    uint64_t x;                 // open state entering the tile (payload encoding)
    uint64_t prefix;            // records kept before this tile
    unsigned int hist[E2S_HIST_BINS];
};

// the wrapped length test of append_entry (ref:ebwt2clust.cpp:56,104)
__device__ __forceinline__ bool keep_len(uint64_t st, uint64_t gend, int32_t min_len, uint32_t* len) {
    *len = uint32_t(gend - st + 1) & 0xffffu;
    return int(*len) >= min_len;
}

template <bool FAST>
__global__ void __launch_bounds__(EM_THREADS, FAST ? 4 : 1) k_cluster_emit(EmitParams p) {
    __shared__ EmitShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // min_len <= 33: which ENDs are dropped is a bit-parallel function of the masks (a START within
    // min_len - 1 positions), so the tile's kept count is known right after the load and the offset
    // look-back overlaps the rest of the tile.  Larger min_len: count in the dense pass first.
    constexpr bool fast_count = FAST;  // the launcher picks FAST iff min_len <= 33
    const int spread = p.min_len >= 2 ? p.min_len - 2 : -1;  // extra positions a START shadows; -1: nothing is dropped

    for (int i = tid; i < E2S_HIST_BINS; i += EM_THREADS) sh.hist[i] = 0;
    if (tid == 0) sh.tile = atomicAdd(&p.res->ticket, 1ull);
    unsigned long long acc_end = 0, acc_bases = 0;
    uint32_t acc_any = 0;

    // Tiles are handed out by a ticket counter, so a tile is only ever waited for by CTAs that started
    // after its owner did: the look-back spin-waits cannot deadlock whatever else occupies the SMs.
    for (;;) {
        __syncthreads();  // (T) ticket visible; shared state of the previous tile is no longer read
        const uint64_t t = sh.tile;
        if (t >= p.num_tiles) break;
        const uint64_t word0 = t * EM_TILE_WORDS + uint64_t(tid) * EM_WPT;
        const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(p.s_words + word0));
        const uint4 e4 = __ldg(reinterpret_cast<const uint4*>(p.e_words + word0));
        const uint32_t S[EM_WPT] = {s4.x, s4.y, s4.z, s4.w};
        const uint32_t E[EM_WPT] = {e4.x, e4.y, e4.z, e4.w};
        const uint64_t tile_gbase = p.global_off + t * uint64_t(EM_TILE_POS);

        // ---- 1. counts, ordinals, first / last event --------------------------------------------------
        uint32_t cS = 0, cE = 0, cD = 0, last = 0, first = 0, first_end = ~0u;
        uint32_t D[EM_WPT] = {0, 0, 0, 0};  // FAST: ENDs of clusters shorter than min_len
#pragma unroll
        for (int j = EM_WPT - 1; j >= 0; --j) {
            if (S[j] | E[j]) {
                const uint32_t fs = __ffs(S[j]), fe = __ffs(E[j]);  // 1-based, 0 = none
                first = 0x80000000u | uint32_t(fe && (!fs || fe < fs));
            }
            if (E[j]) first_end = uint32_t((tid * EM_WPT + j) * 32 + __ffs(E[j]) - 1);
        }
        uint32_t s_prev = __shfl_up_sync(FULL, S[EM_WPT - 1], 1);
        if (lane == 0) s_prev = (fast_count && spread > 0 && tid > 0) ? __ldg(p.s_words + word0 - 1) : 0u;
#pragma unroll
        for (int j = 0; j < EM_WPT; ++j) {
            cS += __popc(S[j]);
            cE += __popc(E[j]);
            if (S[j] | E[j]) {
                const int hs = 31 - __clz(S[j]), he = 31 - __clz(E[j]);  // -1 when the word has none
                last = 0x80000000u | (hs > he ? uint32_t((tid * EM_WPT + j) * 32 + hs + 1) : 0u);
            }
            if (fast_count && spread >= 0) {
                uint32_t sp = S[j];
                const uint32_t pv = j ? S[j - 1] : s_prev;
                for (int d = 1; d <= spread; ++d) sp |= __funnelshift_l(pv, S[j], d);
                D[j] = E[j] & sp;
                cD += __popc(D[j]);
            }
        }
        const uint32_t pk = cS | (cE << 16);
        uint32_t inc = pk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        const uint32_t evm = __ballot_sync(FULL, last != 0);
        const uint32_t wl = evm ? __shfl_sync(FULL, last, 31 - __clz(evm)) : 0u;
        const uint32_t wf = evm ? __shfl_sync(FULL, first, __ffs(evm) - 1) : 0u;
        const uint32_t enm = __ballot_sync(FULL, cE != 0);
        const uint32_t wfe = enm ? __shfl_sync(FULL, first_end, __ffs(enm) - 1) : ~0u;
        uint32_t dinc = cD;  // inclusive scan of the dropped counts (FAST only)
        if (fast_count && spread >= 0) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, dinc, d);
                if (lane >= d) dinc += o;
            }
        }
        const uint32_t wd = dinc;  // lane 31 holds the warp total
        if (lane == 31) {
            sh.wsum[warp] = inc;
            sh.wlast[warp] = wl;
            sh.wfirst[warp] = wf;
            sh.wfirst_end[warp] = wfe;
            sh.wdrop[warp] = wd;
        }
        __syncthreads();  // (1)
        uint32_t base = inc - pk, tot = 0, tlast = 0, tfirst = 0, tfirst_end = ~0u, ndrop = 0, myDbase = dinc - cD;
#pragma unroll
        for (int q = EM_WARPS - 1; q >= 0; --q) {
            if (sh.wfirst[q]) tfirst = sh.wfirst[q];
            if (sh.wfirst_end[q] != ~0u) tfirst_end = sh.wfirst_end[q];
        }
#pragma unroll
        for (int q = 0; q < EM_WARPS; ++q) {
            const uint32_t ws = sh.wsum[q], wq = sh.wlast[q];
            if (q < warp) base += ws;
            tot += ws;
            ndrop += sh.wdrop[q];
            if (q < warp) myDbase += sh.wdrop[q];
            if (wq) tlast = wq;
        }
        const uint32_t nE = tot >> 16;
        const uint32_t mySbase = base & 0xffffu, myEbase = base >> 16;
        const bool has = tlast != 0;
        if (tid == 0) {
            acc_end += nE;
            acc_any |= tlast;
        }

        // ---- 2. warp 0: open state entering the tile (look-back #1) and, when the kept count is already
        //         known, the output offset (look-back #2); the other warps go on to the scatter ------------
        if (warp == 0) {
            const uint32_t op = tlast & 0x7fffffffu;
            if (lane == 0) {
                if (has) desc_store(&p.desc_state[t], ST_INCLUSIVE, op ? (tile_gbase + (op - 1) + OPEN_BIAS) : OPEN_NONE);
                else desc_store(&p.desc_state[t], ST_AGGREGATE, 0);
            }
            uint64_t X = lookback_state(p.desc_state, int64_t(t), p.global_off == 0 ? OPEN_NONE : OPEN_UNKNOWN, lane);
            if (lane == 0) {
                if (!has) desc_store(&p.desc_state[t], ST_INCLUSIVE, X);
                if (t == p.num_tiles - 1)  // state after the whole shard
                    p.res->open_start = has ? (op ? (tile_gbase + (op - 1) + 1) : 0) : (X >= OPEN_BIAS ? X - OPEN_BIAS + 1 : 0);
            }
            // nothing seen since the start of a later shard: open iff this tile begins with an END
            if (X == OPEN_UNKNOWN && !(tfirst & 1u)) X = OPEN_NONE;
            if (lane == 0) sh.x = X;
            if (fast_count || nE == 0) {
                uint32_t total = nE - ndrop;
                uint32_t adj = 0;
                if (nE && X != OPEN_NONE) {  // the first END closes a cluster that started before the tile: exact test
                    uint32_t len;
                    const bool k = X >= OPEN_BIAS && keep_len(X - OPEN_BIAS, tile_gbase + tfirst_end, p.min_len, &len);
                    adj = k ? 0u : 1u;
                    total -= adj;
                }
                if (lane == 0) sh.adj = adj;
                if (lane == 0 && t > 0) desc_store(&p.desc_cnt[t], ST_AGGREGATE, total);
                const uint64_t prefix = t > 0 ? lookback_sum(p.desc_cnt, int64_t(t), lane) : 0;
                if (lane == 0) {
                    desc_store(&p.desc_cnt[t], ST_INCLUSIVE, prefix + total);
                    sh.prefix = prefix;
                    sh.round_total = total;
                    if (t == p.num_tiles - 1) p.res->n_written = prefix + total;
                }
            }
        }
        if (nE == 0) {  // nothing closes in this tile (block-uniform)
            if (tid == 0) sh.tile = atomicAdd(&p.res->ticket, 1ull);
            continue;
        }

        // scatter of the bit positions of window [win, win + EM_CAP) into the shared lists (independent of open_in)
        auto extract = [&](uint32_t win) {
            if (myEbase < win + EM_CAP && myEbase + cE > win) {
                uint32_t ord = myEbase - win;  // may wrap below 0: the unsigned compare rejects those
                uint32_t krank = myEbase - myDbase;  // kept ENDs of the tile before mine
#pragma unroll
                for (int j = 0; j < EM_WPT; ++j) {
                    uint32_t m = E[j];
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t dropped = (D[j] >> b) & 1u;
                        if (ord < EM_CAP) {
                            sh.epos[ord] = uint16_t((tid * EM_WPT + j) * 32 + b);
                            if (fast_count) sh.erank[ord] = dropped ? uint16_t(0xFFFF) : uint16_t(krank);
                        }
                        krank += 1u - dropped;
                        ++ord;
                    }
                }
            }
            if (mySbase < win + EM_CAP && mySbase + cS + 1 > win) {
                uint32_t ord = mySbase + 1 - win;  // slot 0 = START ordinal win - 1
#pragma unroll
                for (int j = 0; j < EM_WPT; ++j) {
                    uint32_t m = S[j];
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        if (ord < EM_CAP + 1) sh.spos[ord] = uint16_t((tid * EM_WPT + j) * 32 + b);
                        ++ord;
                    }
                }
            }
        };

        // dense pass over one window: one lane per END; records in registers (start | len << 48, ~0 = head END),
        // kept mask, per-row kept counts and their exclusive scan.  Returns the kept count of the window.
        uint64_t rec[EM_EPT];
        uint32_t kept;
        auto count_round = [&](uint32_t win, uint64_t X) -> uint32_t {
            const uint32_t cntw = nE - win < EM_CAP ? nE - win : EM_CAP;
            const uint32_t open_in = X != OPEN_NONE;
            kept = 0;
#pragma unroll
            for (int i = 0; i < EM_EPT; ++i) {
                const uint32_t q = uint32_t(tid) + uint32_t(i) * EM_THREADS;  // row = warp + EM_WARPS * i
                rec[i] = 0;
                if (uint32_t(i) * EM_THREADS + uint32_t(warp) * 32 < cntw) {  // warp-uniform
                    bool k = false;
                    if (q < cntw) {
                        const uint64_t gend = tile_gbase + sh.epos[q];
                        if (q == 0 && win == 0 && open_in) {
                            if (X >= OPEN_BIAS) {
                                uint32_t len;
                                k = keep_len(X - OPEN_BIAS, gend, p.min_len, &len);
                                rec[i] = (X - OPEN_BIAS) | (uint64_t(len) << 48);
                            } else {
                                rec[i] = ~0ull;  // the shard's head END: its START is in an earlier shard
                            }
                        } else {
                            const uint64_t st = tile_gbase + sh.spos[q + 1 - open_in];
                            uint32_t len;
                            k = keep_len(st, gend, p.min_len, &len);
                            rec[i] = st | (uint64_t(len) << 48);
                        }
                    }
                    const uint32_t bal = __ballot_sync(FULL, k);
                    if (lane == 0) sh.rowcnt[warp + EM_WARPS * i] = __popc(bal);
                    kept |= uint32_t(k) << i;
                } else if (lane == 0) {
                    sh.rowcnt[warp + EM_WARPS * i] = 0;
                }
            }
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the row counts
                static_assert(EM_ROWS == 64, "two rows per lane");
                const uint32_t c0 = sh.rowcnt[2 * lane], c1 = sh.rowcnt[2 * lane + 1];
                const uint32_t s2 = c0 + c1;
                uint32_t in2 = s2;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(FULL, in2, d);
                    if (lane >= d) in2 += o;
                }
                sh.rowbase[2 * lane] = in2 - s2;
                sh.rowbase[2 * lane + 1] = in2 - s2 + c0;
                if (lane == 31 && !fast_count) sh.round_total = in2;
            }
            __syncthreads();
            return fast_count ? 0u : sh.round_total;
        };

        // writes the kept records of the window held in rec[] at out index obase + rank
        auto write_round = [&](uint32_t win, uint64_t obase, uint64_t tile_end_index) -> uint32_t {
            const uint32_t cntw = nE - win < EM_CAP ? nE - win : EM_CAP;
            uint32_t my_kept = 0;
#pragma unroll
            for (int i = 0; i < EM_EPT; ++i) {
                if (uint32_t(i) * EM_THREADS + uint32_t(warp) * 32 >= cntw) continue;  // warp-uniform
                const bool k = (kept >> i) & 1u;
                const uint32_t bal = __ballot_sync(FULL, k);
                const uint32_t q = uint32_t(tid) + uint32_t(i) * EM_THREADS;
                if (q < cntw) {
                    const uint64_t gend = tile_gbase + sh.epos[q];
                    if (rec[i] == ~0ull) {  // head END of the shard
                        p.res->head_end = gend + 1;
                        if (gend + 2 == p.n_global) p.res->end_nm2_start = ~0ull;
                    } else {
                        const uint64_t st = rec[i] & 0xffffffffffffull;
                        const uint32_t len = uint32_t(rec[i] >> 48);
                        if (gend + 2 == p.n_global) p.res->end_nm2_start = st + 1;
                        if (k) {
                            const uint64_t o = obase + sh.rowbase[warp + EM_WARPS * i] + __popc(bal & lt_mask);
                            if (o < p.cap) {
                                p.out_start[o] = st;
                                p.out_len[o] = uint16_t(len);
                            } else {
                                p.res->overflow = 1;
                            }
                            ++my_kept;
                            acc_bases += len;
                            if (len <= MAX_C_LEN) atomicAdd(&sh.hist[len], 1u);
                            if (o + 1 == tile_end_index) atomicMax(&p.res->last_rec, (unsigned long long)(((o + 1) << 16) | len));
                        }
                    }
                }
            }
            return my_kept;
        };

        extract(0);
        __syncthreads();  // (2) lists, sh.x (and, when the count was bit-parallel, sh.prefix / sh.round_total) are ready
        const uint64_t X = sh.x;
        // The next ticket is taken only now: a claimed tile stays INVALID for its successors until its owner gets to
        // it, so claiming before the look-backs are done would make every successor wait for this whole tile.
        unsigned long long next_ticket = 0;
        if (tid == 0) next_ticket = atomicAdd(&p.res->ticket, 1ull);
        if constexpr (fast_count) {
            // dense write: the rank of every kept END is already in the list, the offset in sh.prefix
            const uint64_t prefix = sh.prefix, tile_end = prefix + sh.round_total;
            const uint32_t open_in = X != OPEN_NONE, adj = sh.adj;
            for (uint32_t win = 0; win < nE; win += EM_CAP) {  // one iteration unless the tile has > EM_CAP ENDs
                if (win) {
                    __syncthreads();  // the previous window's lists are no longer read
                    extract(win);
                    __syncthreads();
                }
                const uint32_t cntw = nE - win < EM_CAP ? nE - win : EM_CAP;
                for (uint32_t q = tid; q < cntw; q += EM_THREADS) {
                    const uint64_t gend = tile_gbase + sh.epos[q];
                    uint32_t r = sh.erank[q];
                    uint64_t st;
                    uint32_t len;
                    bool k;
                    if (q == 0 && win == 0 && open_in) {
                        if (X < OPEN_BIAS) {  // the shard's head END: its START is in an earlier shard
                            p.res->head_end = gend + 1;
                            if (gend + 2 == p.n_global) p.res->end_nm2_start = ~0ull;
                            continue;
                        }
                        st = X - OPEN_BIAS;
                        k = keep_len(st, gend, p.min_len, &len);
                    } else {
                        st = tile_gbase + sh.spos[q + 1 - open_in];
                        len = uint32_t(gend - st + 1);  // <= 32768: no wrap inside a tile
                        k = r != 0xFFFFu;
                        r -= adj;
                    }
                    if (gend + 2 == p.n_global) p.res->end_nm2_start = st + 1;
                    if (k) {
                        const uint64_t o = prefix + r;
                        if (o < p.cap) {
                            p.out_start[o] = st;
                            p.out_len[o] = uint16_t(len);
                        } else {
                            p.res->overflow = 1;
                        }
                        acc_bases += len;
                        if (len <= MAX_C_LEN) atomicAdd(&sh.hist[len], 1u);
                        if (o + 1 == tile_end) atomicMax(&p.res->last_rec, (unsigned long long)(((o + 1) << 16) | len));
                    }
                }
            }
        } else {
            // count all windows, publish, look back, then recompute and write
            uint32_t total = count_round(0, X);
            for (uint32_t win = EM_CAP; win < nE; win += EM_CAP) {
                extract(win);
                __syncthreads();
                total += count_round(win, X);
            }
            if (warp == 0) {
                if (lane == 0 && t > 0) desc_store(&p.desc_cnt[t], ST_AGGREGATE, total);
                const uint64_t prefix = t > 0 ? lookback_sum(p.desc_cnt, int64_t(t), lane) : 0;
                if (lane == 0) {
                    desc_store(&p.desc_cnt[t], ST_INCLUSIVE, prefix + total);
                    sh.prefix = prefix;
                    if (t == p.num_tiles - 1) p.res->n_written = prefix + total;
                }
            }
            __syncthreads();
            const uint64_t prefix = sh.prefix;
            uint64_t obase = prefix;
            for (uint32_t win = 0; win < nE; win += EM_CAP) {
                if (nE > EM_CAP) {  // the lists hold the last window: rebuild (single-window tiles still hold window 0)
                    __syncthreads();
                    extract(win);
                    __syncthreads();
                    const uint32_t c = count_round(win, X);
                    write_round(win, obase, prefix + total);
                    obase += c;
                } else {
                    write_round(win, obase, prefix + total);
                }
            }
        }
        if (tid == 0) sh.tile = next_ticket;
    }

#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc_bases += __shfl_xor_sync(FULL, acc_bases, d);
    if (lane == 0 && acc_bases) atomicAdd(&p.res->n_bases, acc_bases);
    if (tid == 0) {
        if (acc_end) atomicAdd(&p.res->n_end, acc_end);
        if (acc_any) atomicOr(&p.res->any_event, 1ull);
    }
    __syncthreads();
    for (int i = tid; i < E2S_HIST_BINS; i += EM_THREADS)
        if (sh.hist[i]) atomicAdd(&p.res->hist[i], (unsigned long long)sh.hist[i]);
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
    const bool fast = p.min_len <= 33;
    int occ = 0;
    cudaError_t e = fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cluster_emit<true>, EM_THREADS, 0)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cluster_emit<false>, EM_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    uint64_t grid = uint64_t(sm_count) * occ;
    if (grid > p.num_tiles) grid = p.num_tiles;
    if (fast) k_cluster_emit<true><<<dim3(unsigned(grid)), dim3(EM_THREADS), 0, stream>>>(p);
    else k_cluster_emit<false><<<dim3(unsigned(grid)), dim3(EM_THREADS), 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace e2s
