// Phase 1 (ebwt2clust) in ONE pass over the resident one-byte LCP: k_cluster_scan.
//
// cluster_lm / append_entry (ref:ebwt2clust.cpp:54-139) in their stencil form (cluster.cu): the
// kernel reads every LCP byte once, keeps the START / END bit masks of a tile in REGISTERS (they are never written to
// memory), and goes straight on to the records:
//
//   tiles     16 384 positions; persistent CTAs take tile numbers from a ticket (so every tile with a smaller number is
//             owned by a CTA that is already running) and keep SC_STAGES tiles in flight each: the LCP bytes come in as
//             one TMA tensor box (128-byte rows, hardware 128B swizzle => each thread's four 16-byte reads of its 64
//             bytes are bank-conflict free), the tile's window of the base-code bit planes (fused prefilter) as one
//             bulk copy, both signalling the same mbarrier.
//   masks     packed byte compares, four positions per instruction (k_lcp_flags8's arithmetic); each thread owns 64
//             consecutive positions = one 64-bit word per mask.
//   ranks     START and END bits alternate, so the START that pairs with the tile's q-th END is the START that has
//             exactly q ENDs before it: ONE block scan of the per-thread END counts ranks both; every START writes its
//             position to list slot [ENDs before it], every END writes (position, kept rank) to slot [its rank].
//   min_len   ENDs of clusters shorter than min_len are a bit-parallel function of the masks (a START at most
//             min_len - 2 positions before, min_len <= 33).  The only END whose START is not in its tile is the tile's
//             first event ("carried" END): that one is tested exactly, wrapped 16-bit length included, against the
//             carried-in START -- a tile is shorter than 65 503 positions, so no other END can wrap.
//   look-back two 64-bit words per tile, each validated by an epoch (no memset between launches):
//               A = state after the tile (closed / open at position p / nothing happened): local, published at once;
//                   a tile without events re-publishes the resolved state once it knows it
//               B = kept-record count: aggregate (after the carried END has been tested against A of the tiles before),
//                   then inclusive prefix -- the classic decoupled look-back, 32 predecessors per poll.
//   records   one lane per END of the tile: START from the list (or the carried-in START), length mod 2^16, stores
//             (start u64, len u16) in position order, length histogram + n_bases of statistics()
//             (ref:clust2snp.cpp:899-907), and -- fused mode -- the one-popcount bound of the BWT prefilter
//             (planes.cuh) on the plane window in shared memory; survivors are appended to the list K3x reads.
//
// Algorithmic bytes: 1 B/position read + 10 B per written record (+ 0.25 B/position of bit planes in fused mode).
// Not handled here (callers fall back to k_lcp_flags + k_cluster_emit): min_len > 33, shards whose LCP does not fit a byte.

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "internal.h"
#include "planes.cuh"

namespace e2s {

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_V = 64;                       // positions per thread
constexpr int SC_T = SC_THREADS * SC_V;        // 16384 positions per tile
constexpr int SC_W = SC_V / 4;                 // 16 packed words per thread
constexpr int SC_STAGES = 2;                   // LCP tiles in flight / being read per CTA (a stage is free once its bytes are in registers)
constexpr int SC_PSLOTS = SC_STAGES + 1;       // plane windows: a window is read until its tile's records are written, after the stage was refilled
constexpr int SC_OCC = 4;                      // resident CTAs per SM the register / shared-memory budget is sized for
constexpr int SC_PF_QUADS = SC_T / 64 + PL_PAD / 64;  // plane quads of one tile and of the PL_PAD positions before it
constexpr int SC_PF_BYTES = SC_PF_QUADS * 16;  // 4144
constexpr int SC_PF_STRIDE = 4224;             // slot stride (128-byte multiple)
constexpr int SC_DYN_SMEM = SC_STAGES * SC_T + SC_PSLOTS * SC_PF_STRIDE + 1024;  // + slack for the 1024-byte alignment of the TMA boxes
constexpr int SC_CAP = 1024;                   // ENDs per list window (a typical tile lists ~300)

// open-cluster state (as in cluster.cu)
constexpr uint64_t OPEN_NONE = 0, OPEN_UNKNOWN = 1, OPEN_BIAS = 2;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u8(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ uint32_t gather_b7(uint32_t x) {  // bit 7 of bytes 0..3 -> bits 28..31
    return (x & 0x80808080u) * 0x00204081u;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    return (uint64_t(__shfl_sync(FULL, uint32_t(v >> 32), src)) << 32) | __shfl_sync(FULL, uint32_t(v), src);
}
__device__ __forceinline__ uint64_t below(int b) { return (uint64_t(1) << b) - 1; }  // bits [0, b), b < 64

struct ScanShared {
    uint64_t full_bar[SC_STAGES];
    uint32_t wsum[2][SC_WARPS];   // per warp: #END | #dropped << 16 (double buffered by tile parity: a tile without ENDs has no barrier (B))
    int wls[2][SC_WARPS], wle[2][SC_WARPS], wfs[2][SC_WARPS], wfe[2][SC_WARPS];  // last / first START / END of the warp (tile-local)
    uint64_t wS[SC_WARPS];        // START word of the warp's last thread (min_len >= 3 only)
    uint64_t prev_S[2];           // [it & 1] = START word of the last thread of the chunk's tile it - 1 (min_len >= 3 only)
    unsigned int hist[E2S_HIST_BINS];
    uint16_t s_pos[SC_CAP];       // [q] = tile-local position of the START that pairs with the window's q-th END
    uint32_t e_ent[SC_CAP];       // [q] = END position | kept rank << 14 | kept << 28
};

constexpr int NO_POS = 0x7fffffff;

}  // namespace

// One CTA per CHUNK of consecutive tiles, no communication between CTAs: the open-cluster state and the record count are
// carried from tile to tile in registers (every thread derives them from the same shared summaries), the records go to the
// chunk's own segment of the record arrays.  What a chunk cannot know -- whether a cluster is open when it starts -- only
// matters for its first event: if that is an END ("head" of the chunk) the record is left to k_chunk_resolve, which sees all
// chunks' summaries (scan.cu, below).
__global__ void __launch_bounds__(SC_THREADS, SC_OCC) k_cluster_scan(const __grid_constant__ CUtensorMap tmap, Scan8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* plane_slots = stages + size_t(SC_STAGES) * SC_T;  // SC_PSLOTS windows of SC_PF_STRIDE bytes
    __shared__ ScanShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool pf = p.pf_mcov != 0;
    const int spread = p.min_len >= 2 ? p.min_len - 2 : -1;  // extra positions a START shadows; -1: nothing is dropped
    const uint32_t kk = p.k > 128u ? 128u : p.k;             // bytes are <= 127: k >= 128 never matches
    const uint32_t kadd = (128u - kk) * 0x01010101u;
    const uint32_t c = blockIdx.x;
    const uint32_t t_lo = c * p.tiles_per_chunk < p.num_tiles ? c * p.tiles_per_chunk : p.num_tiles;
    const uint32_t t_hi = t_lo + p.tiles_per_chunk < p.num_tiles ? t_lo + p.tiles_per_chunk : p.num_tiles;
    const uint32_t n_my = t_hi - t_lo;
    const uint64_t seg_base = uint64_t(c) * p.seg_cap;

    for (int i = tid; i < E2S_HIST_BINS; i += SC_THREADS) sh.hist[i] = 0;
    // thread 0: the it-th tile of the chunk -> LCP box to stage it % SC_STAGES, plane window to slot it % SC_PSLOTS
    auto issue = [&](uint32_t it) {
        if (it >= n_my) return;
        const int stage = int(it % SC_STAGES);
        const uint32_t t = t_lo + it;
        mbar_expect_tx(&sh.full_bar[stage], SC_T + (pf ? SC_PF_BYTES : 0));
        tma_load_2d_u8(stages + size_t(stage) * SC_T, &tmap, 0, int(t * (SC_T / 128)), &sh.full_bar[stage]);
        if (pf) bulk_g2s(plane_slots + size_t(it % SC_PSLOTS) * SC_PF_STRIDE, p.planes + uint64_t(t) * (SC_T / 64), SC_PF_BYTES, &sh.full_bar[stage]);
    };
    if (tid == 0) {
        sh.prev_S[0] = 0;  // before the chunk: 0 (the chunk's head END is tested in k_chunk_resolve)
        for (int s = 0; s < SC_STAGES; ++s) mbar_init(&sh.full_bar[s], 1);
        fence_mbar_init();
        for (uint32_t s = 0; s < uint32_t(SC_STAGES); ++s) issue(s);
    }
    __syncthreads();

    // chunk state, identical in every thread
    uint64_t X = OPEN_UNKNOWN;   // open-cluster state; unknown until the chunk's first event
    uint64_t cnt = 0;            // records written to the segment so far
    uint64_t n_end = 0;          // ENDs seen (head included)
    uint64_t head_end = 0;       // 1 + global position of the chunk's head END, 0 = none
    bool seen = false;           // the chunk has had an event
    unsigned long long acc_bases = 0;  // (per thread) sum of the lengths I wrote
    uint64_t my_last_o = ~0ull;        // (per thread) segment index and length of the last record I wrote
    uint32_t my_last_len = 0;

    for (uint32_t it = 0; it < n_my; ++it) {
        const int stage = it % SC_STAGES;
        const uint32_t parity = (it / SC_STAGES) & 1;
        const int pb = it & 1;
        const uint32_t t = t_lo + it;
        const uint64_t tile_base = uint64_t(t) * SC_T;
        const uint64_t tile_gbase = p.global_off + tile_base;
        const uint8_t* tile = stages + size_t(stage) * SC_T;
        const uint4* pf_win = reinterpret_cast<const uint4*>(plane_slots + size_t(it % SC_PSLOTS) * SC_PF_STRIDE);
        const bool interior = tile_gbase != 0 && tile_base + SC_T <= p.n_local && tile_gbase + SC_T < p.n_global;

        uint32_t g_prev = 0, g_next = 0;  // the bytes around the tile: issue the global loads before waiting for the tile
        if (tid == 0) g_prev = *reinterpret_cast<const uint32_t*>(p.lcp8 + (int64_t(tile_base) - 4));
        if (tid == SC_THREADS - 1) g_next = p.lcp8[tile_base + SC_T];

        mbar_wait(&sh.full_bar[stage], parity);

        // ---- my 64 bytes: row r = tid / 2 of 128 bytes, logical 16-byte chunks 4 (tid & 1) + j at physical chunk ^ (r & 7)
        uint32_t w[SC_W], pw, nw;
        {
            const uint32_t r = uint32_t(tid) >> 1, c0 = (uint32_t(tid) & 1u) * 4u, x = r & 7u;
            const uint8_t* row = tile + r * 128u;
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                const uint4 v = lds128(row + (((c0 + j) ^ x) << 4));
                w[4 * j + 0] = v.x;
                w[4 * j + 1] = v.y;
                w[4 * j + 2] = v.z;
                w[4 * j + 3] = v.w;
            }
            // the word before my bytes and the byte after them
            if (tid == 0) pw = g_prev;
            else if (c0) pw = *reinterpret_cast<const uint32_t*>(row + ((3u ^ x) << 4) + 12);
            else pw = *reinterpret_cast<const uint32_t*>(row - 128 + ((7u ^ ((r - 1u) & 7u)) << 4) + 12);
            if (tid == SC_THREADS - 1) nw = g_next;
            else if (c0) nw = *reinterpret_cast<const uint32_t*>(row + 128 + ((0u ^ ((r + 1u) & 7u)) << 4));
            else nw = *reinterpret_cast<const uint32_t*>(row + ((4u ^ x) << 4));
        }

        // ---- START / END masks of my 64 positions (k_lcp_flags8's packed compares)
        uint64_t S, E;
        {
            uint32_t Gh[2], Ah[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t G = 0, A = 0;
#pragma unroll
                for (int i = 7; i >= 0; --i) {
                    const int q = 8 * h + i;
                    const uint32_t cur = w[q];
                    const uint32_t carry = __umulhi(q == 0 ? pw : w[q - 1], 256u) + 0x7f7f7f7fu;  // (prev >> 24) + 0x7f7f7f7f
                    G = __funnelshift_l(gather_b7(cur + kadd), G, 4);
                    A = __funnelshift_l(gather_b7(cur * 255u + carry), A, 4);
                }
                Gh[h] = G;
                Ah[h] = A;
            }
            const uint64_t G = (uint64_t(Gh[1]) << 32) | Gh[0];
            const uint64_t A = (uint64_t(Ah[1]) << 32) | Ah[0];
            const uint32_t v_m2 = (pw >> 16) & 0xffu, v_m1 = pw >> 24, v_p = nw & 0xffu, v_last = w[SC_W - 1] >> 24;
            const uint64_t g_m1b = v_m1 >= p.k, g_pb = v_p >= p.k;
            const uint64_t a_V = v_last > v_p;
            const uint64_t Gn = (G >> 1) | (g_pb << (SC_V - 1));  // ge(j+1)
            const uint64_t An = (A >> 1) | (a_V << (SC_V - 1));   // lcp[j] > lcp[j+1]
            E = G & ((A & ~An) | ~Gn);
            uint64_t e_prev = g_m1b & ((uint64_t(v_m2 > v_m1) & ((~A) & 1u)) | ((~G) & 1u));
            uint64_t vm = ~uint64_t(0);
            if (!interior) {  // first tile of the eBWT, the tile holding position n_global - 1, tiles reaching past n_local
                const uint64_t my_base = tile_base + uint64_t(tid) * SC_V;
                const uint64_t gpos = p.global_off + my_base;
                if (gpos == 0) {  // the init special cases of ref:ebwt2clust.cpp:83-86
                    E &= ~uint64_t(1);
                    e_prev = 0;
                    if ((G & 1u) && !(G & 2u)) E |= 2u;
                }
                const int64_t nvalid = int64_t(p.n_local) - int64_t(my_base);
                vm = nvalid >= SC_V ? ~uint64_t(0) : (nvalid <= 0 ? 0 : ((uint64_t(1) << nvalid) - 1));
                const int64_t last = int64_t(p.n_global) - 1 - int64_t(gpos);  // END(n_global-1): host tail rule
                if (last >= 0 && last < SC_V) E &= ~(uint64_t(1) << last);
                E &= vm;
            }
            const uint64_t Gp = (G << 1) | g_m1b;
            const uint64_t Ep = (E << 1) | e_prev;
            S = G & (~Gp | Ep) & vm;
        }

        // ---- D = ENDs of clusters shorter than min_len: a START at the same position or up to `spread` positions before
        uint64_t D = 0;
        if (spread >= 0) {
            uint64_t sm = S;
            if (spread >= 1) {  // (kernel-uniform) the previous thread's START word: neighbours by shuffle, warps through shared memory
                if (lane == 31) sh.wS[warp] = S;
                if (tid == SC_THREADS - 1) sh.prev_S[pb ^ 1] = S;  // for the next tile
                __syncthreads();
                const uint64_t before_tile = sh.prev_S[pb];
                uint64_t lo = shfl64(S, (lane + 31) & 31);
                if (lane == 0) lo = warp ? sh.wS[warp - 1] : before_tile;
                uint64_t hi = S;
                int width = 1;  // sm = OR of (S << d), d = 0 .. width - 1, over the 128 bits lo:hi
                while (2 * width <= spread + 1) {
                    hi |= (hi << width) | (lo >> (64 - width));
                    lo |= lo << width;
                    width *= 2;
                }
                const int rest = spread + 1 - width;
                if (rest) hi |= (hi << rest) | (lo >> (64 - rest));
                sm = hi;
            }
            D = E & sm;
        }

        // ---- ranks: one block scan of #END | #dropped << 16; last / first events of the tile
        const uint32_t cE = __popcll(E), cD = __popcll(D);
        const uint32_t pk = cE | (cD << 16);
        uint32_t inc = pk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        {
            const int ls = __reduce_max_sync(FULL, S ? tid * SC_V + 63 - __clzll(S) : -1);
            const int le = __reduce_max_sync(FULL, E ? tid * SC_V + 63 - __clzll(E) : -1);
            const int fs = __reduce_min_sync(FULL, S ? tid * SC_V + __ffsll(S) - 1 : NO_POS);
            const int fe = __reduce_min_sync(FULL, E ? tid * SC_V + __ffsll(E) - 1 : NO_POS);
            if (lane == 31) sh.wsum[pb][warp] = inc;
            if (lane == 0) {
                sh.wls[pb][warp] = ls;
                sh.wle[pb][warp] = le;
                sh.wfs[pb][warp] = fs;
                sh.wfe[pb][warp] = fe;
            }
        }
        __syncthreads();  // (A) per-warp summaries.  Every thread has this tile's bytes in registers and has written the records
        // of the tile before: the stage is refilled with the chunk's tile it + SC_STAGES (plane slot: the previous tile's)
        if (tid == 0) issue(it + SC_STAGES);

        uint32_t base = inc - pk, tot = 0;
#pragma unroll
        for (int q = 0; q < SC_WARPS; ++q) {
            const uint32_t ws = sh.wsum[pb][q];
            if (q < warp) base += ws;
            tot += ws;
        }
        const uint32_t nE = tot & 0xffffu, nD = tot >> 16;
        const uint32_t baseE = base & 0xffffu, baseD = base >> 16;
        int t_ls = -1, t_le = -1, t_fs = NO_POS, t_fe = NO_POS;
#pragma unroll
        for (int q = 0; q < SC_WARPS; ++q) {
            t_ls = max(t_ls, sh.wls[pb][q]);
            t_le = max(t_le, sh.wle[pb][q]);
            t_fs = min(t_fs, sh.wfs[pb][q]);
            t_fe = min(t_fe, sh.wfe[pb][q]);
        }
        const bool has_event = t_ls >= 0 || t_le >= 0;
        const bool carried = t_fe != NO_POS && t_fe < t_fs;  // the tile's first event is an END: its START lies before the tile
        // (a START and an END at the same position: the START comes first, t_fe == t_fs is not carried)

        // ---- the carried END: tested exactly on the wrapped length (ref:ebwt2clust.cpp:56,104) against the carried-in START --
        // it is the only END of the tile that can be 65 536 or more positions from its START.  Before the chunk's first
        // event the START is not known here: that END is the chunk's head, left to k_chunk_resolve.
        uint32_t adj = 0;  // 1: the carried END is not written by this tile
        const bool is_head = carried && !seen;
        if (carried) {
            adj = 1;
            if (seen && X >= OPEN_BIAS) {
                const uint32_t len = uint32_t(tile_gbase + uint64_t(t_fe) - (X - OPEN_BIAS) + 1) & 0xffffu;
                adj = int(len) >= p.min_len ? 0u : 1u;
            }
            if (is_head) head_end = tile_gbase + uint64_t(t_fe) + 1;
        }
        const uint64_t x_in = X, prefix = seg_base + cnt;
        cnt += nE - nD - adj;
        n_end += nE;
        if (has_event) {
            X = t_ls > t_le ? tile_gbase + uint64_t(t_ls) + OPEN_BIAS : OPEN_NONE;
            seen = true;
        }

        // ---- lists + records, one window of SC_CAP ENDs at a time (one iteration unless the tile is unusually dense)
        for (uint32_t win = 0; win < nE; win += SC_CAP) {
            if (win) __syncthreads();  // the previous window's lists are no longer read
            uint64_t m = S;
            while (m) {  // a START goes to the slot of the END it pairs with: the one with as many ENDs before it
                const int b = __ffsll(m) - 1;
                m &= m - 1;
                const uint32_t q = baseE + __popcll(E & below(b)) - win;
                if (q < uint32_t(SC_CAP)) sh.s_pos[q] = uint16_t(tid * SC_V + b);
            }
            m = E;
            while (m) {
                const int b = __ffsll(m) - 1;
                m &= m - 1;
                const uint32_t qa = baseE + __popcll(E & below(b));
                const uint32_t q = qa - win;
                if (q < uint32_t(SC_CAP)) {
                    const uint32_t r = qa - (baseD + __popcll(D & below(b)));
                    sh.e_ent[q] = uint32_t(tid * SC_V + b) | (r << 14) | ((uint32_t((D >> b) & 1u) ^ 1u) << 28);
                }
            }
            __syncthreads();  // (B) lists
            const uint32_t n_win = nE - win < uint32_t(SC_CAP) ? nE - win : uint32_t(SC_CAP);
            for (uint32_t i = tid; i < n_win; i += SC_THREADS) {
                const uint32_t ent = sh.e_ent[i];
                const uint32_t e = ent & 0x3fffu, r = (ent >> 14) & 0x3fffu;
                const uint64_t gend = tile_gbase + e;
                const bool first_carried = carried && win + i == 0;
                uint64_t st;
                if (first_carried) {
                    if (is_head) continue;       // k_chunk_resolve writes (or drops) it, and answers for position n_global - 2
                    st = x_in - OPEN_BIAS;       // (x_in is an open state: START and END bits alternate)
                } else {
                    st = tile_gbase + sh.s_pos[i];
                }
                if (gend + 2 == p.n_global) p.res->end_nm2_start = st + 1;  // decides the post-EOF phantom, kept or not
                if (first_carried ? adj != 0 : !((ent >> 28) & 1u)) continue;
                const uint32_t len = uint32_t(gend - st + 1) & 0xffffu;
                const uint64_t o = prefix + r - ((carried && !first_carried) ? adj : 0u);
                if (o - seg_base < p.seg_cap) {
                    p.seg_start[o] = st;
                    p.seg_len[o] = uint16_t(len);
                } else {
                    p.res->overflow = 1;
                }
                acc_bases += len;
                if (len <= uint32_t(MAX_C_LEN)) atomicAdd(&sh.hist[len], 1u);
                my_last_o = o;
                my_last_len = len;
                // fused BWT prefilter of find_variants (ref:clust2snp.cpp:402-429; planes.cuh): the one-popcount bound.  Fewer than
                // mcov positions with a base code other than the first position's => at most one frequent code => the
                // cluster cannot pass; everything else goes to the exact test (K3x).
                if (pf && len >= 2 * p.pf_mcov && len <= uint32_t(MAX_C_LEN)) {
                    bool pass = st + PL_PAD < tile_gbase;  // the 16-bit length wrapped: the analysed range lies far before this window
                    if (!pass) {
                        const uint64_t b_lo = st + PL_PAD - tile_gbase, b_last = b_lo + len - 1;
                        const uint32_t q_lo = uint32_t(b_lo >> 6), q_last = uint32_t(b_last >> 6);
                        unsigned long long f0 = 0, f1 = 0;
                        uint32_t others = 0;
                        for (uint32_t q = q_lo; q <= q_last; ++q) {
                            const uint4 v = pf_win[q];
                            const unsigned long long x0 = (uint64_t(v.y) << 32) | v.x, x1 = (uint64_t(v.w) << 32) | v.z;
                            unsigned long long mask = ~0ull;
                            if (q == q_lo) {
                                mask = ~0ull << (b_lo & 63);
                                f0 = 0ull - ((x0 >> (b_lo & 63)) & 1ull);
                                f1 = 0ull - ((x1 >> (b_lo & 63)) & 1ull);
                            }
                            if (q == q_last) mask &= ~0ull >> (63 - (b_last & 63));
                            others += __popcll(((x0 ^ f0) | (x1 ^ f1)) & mask);
                        }
                        pass = others >= p.pf_mcov;
                    }
                    if (pass) {
                        const unsigned long long at = atomicAdd(&p.res->n_pf, 1ull);
                        if (at < p.pf_cap) p.pf_list[at] = SurvEntry{st, st - p.global_off, len, 0u};
                    }
                }
            }
        }
    }

    // ---- what the chunk leaves for k_chunk_resolve
    __syncthreads();
    for (int i = tid; i < E2S_HIST_BINS; i += SC_THREADS)
        if (sh.hist[i]) atomicAdd(&p.res->hist[i], (unsigned long long)sh.hist[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc_bases += __shfl_xor_sync(FULL, acc_bases, d);
    ChunkRec* cr = p.chunks + c;
    if (lane == 0 && acc_bases) atomicAdd(&cr->n_bases, acc_bases);
    if (cnt && my_last_o + 1 == seg_base + cnt) cr->last_len = my_last_len;  // (one thread: the writer of the chunk's last record)
    if (tid == 0) {
        cr->own_count = cnt;
        cr->head_end = head_end;
        cr->last_state = seen ? (X >= OPEN_BIAS ? X : 1ull) : 0ull;  // 0: no event; 1: closed; >= 2: OPEN_BIAS + global START
        cr->n_end = n_end;
        if (c == 0 && p.tail_lcp) {
            p.res->tail_lcp_nm2 = p.tail_lcp[0];
            p.res->tail_lcp_nm1 = p.tail_lcp[1];
            p.res->tail_bwt_nm1 = p.tail_bwt[0];
        }
    }
}

// =============================================================================================
// k_chunk_resolve: what the chunks could not know (one CTA; a few hundred chunks)
// =============================================================================================
// state entering a chunk = state left by the nearest earlier chunk that had an event (else the state before the shard);
// the chunk's head END closes the cluster that state holds open: exact length test, one record in front of the chunk's own;
// an END without any START in the shard is the SHARD's head (resolved by e2s_cluster_merge across shards).  Then the
// exclusive scan of the chunks' record counts (= where each segment goes in the position-ordered list) and the shard totals.
constexpr int RS_THREADS = 1024;

__global__ void __launch_bounds__(RS_THREADS) k_chunk_resolve(ResolveParams p) {
    __shared__ unsigned long long s_state[RS_THREADS];
    __shared__ unsigned long long s_cnt[RS_THREADS];
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_tot[4];  // records, ENDs, bases, (last chunk with records) + 1
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
    const bool act = uint32_t(c) < p.n_chunks;
    ChunkRec rec;
    rec.own_count = rec.head_end = rec.last_state = rec.n_end = rec.n_bases = rec.last_len = 0;
    if (act) rec = p.chunks[c];
    s_state[c] = rec.last_state;
    if (c < 4) s_tot[c] = 0;
    __syncthreads();
    uint64_t X = p.init_state;
    for (int j = c - 1; j >= 0; --j)
        if (s_state[j]) {
            X = s_state[j] >= OPEN_BIAS ? s_state[j] : OPEN_NONE;
            break;
        }
    uint64_t h_start = 0;
    uint32_t h_len = 0, h_kept = 0;
    if (act && rec.head_end) {
        const uint64_t e = rec.head_end - 1;
        if (X >= OPEN_BIAS) {
            h_start = X - OPEN_BIAS;
            h_len = uint32_t(e - h_start + 1) & 0xffffu;  // append_entry's wrapped length (ref:ebwt2clust.cpp:56,104)
            h_kept = int(h_len) >= p.min_len ? 1u : 0u;
            if (e + 2 == p.n_global) p.res->end_nm2_start = h_start + 1;
        } else {  // no START in this shard: the shard's head
            p.res->head_end = e + 1;
            if (e + 2 == p.n_global) p.res->end_nm2_start = ~0ull;
        }
    }
    const unsigned long long mine = rec.own_count + h_kept;
    unsigned long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = shfl64(inc, lane >= d ? lane - d : lane);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    s_cnt[c] = mine;
    __syncthreads();
    unsigned long long off = inc - mine;
    for (int q = 0; q < warp; ++q) off += s_warp[q];
    if (act) {
        ChunkSeg sg;
        sg.off = off;
        sg.own_count = rec.own_count;
        sg.head_start = h_start;
        sg.head_len = h_len;
        sg.head_kept = h_kept;
        p.segs[c] = sg;
        if (h_kept) {
            if (h_len <= uint32_t(MAX_C_LEN)) atomicAdd(&p.res->hist[h_len], 1ull);
            // the head record's turn at the fused BWT prefilter (its positions lie before the chunk: global planes)
            if (p.pf_mcov && h_len >= 2 * p.pf_mcov && h_len <= uint32_t(MAX_C_LEN)) {
                bool pass = h_start < p.global_off;  // (cannot happen for a record with a known START; kept for safety)
                if (!pass) pass = frequent_bound(p.planes, int64_t(h_start - p.global_off), h_len, p.pf_mcov);
                if (pass) {
                    const unsigned long long at = atomicAdd(&p.res->n_pf, 1ull);
                    if (at < p.pf_cap) p.pf_list[at] = SurvEntry{h_start, h_start - p.global_off, h_len, 0u};
                }
            }
        }
        atomicAdd(&s_tot[0], mine);
        atomicAdd(&s_tot[1], rec.n_end);
        atomicAdd(&s_tot[2], rec.n_bases + (h_kept ? h_len : 0u));
        if (mine) atomicMax(&s_tot[3], (unsigned long long)(c + 1));
    }
    __syncthreads();
    if (act && mine && s_tot[3] == (unsigned long long)(c + 1))  // the list's last record is this chunk's last
        p.res->last_rec = (s_tot[0] << 16) | (rec.own_count ? rec.last_len : h_len);
    if (c == 0) {
        p.res->n_written = s_tot[0];
        p.res->n_end = s_tot[1];
        p.res->n_bases = s_tot[2];
        uint64_t x_out = p.init_state;
        bool any = false;
        for (int j = int(p.n_chunks) - 1; j >= 0; --j)
            if (s_state[j]) {
                x_out = s_state[j] >= OPEN_BIAS ? s_state[j] : OPEN_NONE;
                any = true;
                break;
            }
        p.res->any_event = any ? 1ull : 0ull;
        p.res->open_start = x_out >= OPEN_BIAS ? x_out - OPEN_BIAS + 1 : 0;
    }
}

// segments -> position-ordered contiguous list: as SoA (start u64, len u16) or as the 10-byte records of the .clusters
// file (ref:ebwt2clust.cpp:58-59)
__global__ void __launch_bounds__(256) k_export_records(const ChunkSeg* __restrict__ segs, uint32_t n_chunks, uint64_t seg_cap,
                                                        const uint64_t* __restrict__ seg_start, const uint16_t* __restrict__ seg_len,
                                                        uint64_t* __restrict__ out_start, uint16_t* __restrict__ out_len,
                                                        uint16_t* __restrict__ out_packed) {
    for (uint32_t c = blockIdx.y; c < n_chunks; c += gridDim.y) {
        const ChunkSeg sg = segs[c];
        const uint64_t total = sg.own_count + sg.head_kept;
        const uint64_t src0 = uint64_t(c) * seg_cap;
        for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
            uint64_t st;
            uint16_t ln;
            if (sg.head_kept && i == 0) {
                st = sg.head_start;
                ln = uint16_t(sg.head_len);
            } else {
                st = seg_start[src0 + i - sg.head_kept];
                ln = seg_len[src0 + i - sg.head_kept];
            }
            const uint64_t o = sg.off + i;
            if (out_packed) {
                uint16_t* q = out_packed + o * 5;
                q[0] = uint16_t(st);
                q[1] = uint16_t(st >> 16);
                q[2] = uint16_t(st >> 32);
                q[3] = uint16_t(st >> 48);
                q[4] = ln;
            } else {
                out_start[o] = st;
                out_len[o] = ln;
            }
        }
    }
}

// =============================================================================================
// host side
// =============================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled scan_encode_fn() {
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

uint64_t scan_num_tiles(uint64_t n_local) { return (n_local + SC_T - 1) / SC_T; }

static int scan_occupancy(int* out) {
    static int occ_dev[64] = {0};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = occ_dev[dev & 63];
    if (!occ) {
        cudaError_t e = cudaFuncSetAttribute(k_cluster_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SC_DYN_SMEM));
        if (e != cudaSuccess) return int(e);
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_cluster_scan, SC_THREADS, SC_DYN_SMEM);
        if (e != cudaSuccess) return int(e);
        if (o < 1) return int(cudaErrorLaunchOutOfResources);
        occ = o;
    }
    *out = occ;
    return 0;
}

// chunks of a shard: one per CTA the device holds at once (at most SCAN_MAX_CHUNKS), at least one tile each
cudaError_t scan_plan(uint64_t n_local, int sm_count, uint32_t* n_chunks, uint32_t* tiles_per_chunk) {
    int occ = 0;
    const int rc = scan_occupancy(&occ);
    if (rc) return cudaError_t(rc);
    const uint64_t nt = scan_num_tiles(n_local);
    uint64_t g = uint64_t(sm_count) * occ;
    if (g > SCAN_MAX_CHUNKS) g = SCAN_MAX_CHUNKS;
    if (g > nt) g = nt;
    if (g < 1) g = 1;
    const uint64_t tpc = (nt + g - 1) / g;
    *tiles_per_chunk = uint32_t(tpc < 1 ? 1 : tpc);
    *n_chunks = uint32_t(nt ? (nt + *tiles_per_chunk - 1) / *tiles_per_chunk : 1);
    return cudaSuccess;
}

cudaError_t launch_scan(const Scan8Params& p0, uint64_t alloc_r, cudaStream_t stream) {
    PFN_encodeTiled enc = scan_encode_fn();
    if (!enc) return cudaErrorNotSupported;
    Scan8Params p = p0;
    p.num_tiles = uint32_t(scan_num_tiles(p.n_local));
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {128, cuuint64_t(alloc_r / 128)};  // the padded byte array from local position 0 on, as rows of 128 bytes
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {128, cuuint32_t(SC_T / 128)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(p.lcp8), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    int occ = 0;
    const int rc = scan_occupancy(&occ);
    if (rc) return cudaError_t(rc);
    k_cluster_scan<<<dim3(p.n_chunks), dim3(SC_THREADS), SC_DYN_SMEM, stream>>>(tmap, p);
    return cudaGetLastError();
}

cudaError_t launch_chunk_resolve(const ResolveParams& p, cudaStream_t stream) {
    if (p.n_chunks > uint32_t(RS_THREADS)) return cudaErrorInvalidValue;
    k_chunk_resolve<<<1, RS_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_export_records(const ChunkSeg* segs, uint32_t n_chunks, uint64_t seg_cap, const uint64_t* seg_start,
                                  const uint16_t* seg_len, uint64_t* out_start, uint16_t* out_len, uint8_t* out_packed,
                                  cudaStream_t stream) {
    if (!n_chunks) return cudaSuccess;
    uint64_t bx = (seg_cap + 255) / 256;
    if (bx > 16) bx = 16;
    if (bx < 1) bx = 1;
    k_export_records<<<dim3(unsigned(bx), n_chunks > 1024 ? 1024 : n_chunks), 256, 0, stream>>>(
        segs, n_chunks, seg_cap, seg_start, seg_len, out_start, out_len, reinterpret_cast<uint16_t*>(out_packed));
    return cudaGetLastError();
}

}  // namespace e2s
