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
constexpr int SC_OCC = 4;                      // resident CTAs per SM the register / shared-memory budget is sized for
constexpr int SC_PSLOTS = SC_STAGES + 1;       // plane windows: a window is read until its tile's records are written, after the stage was refilled
constexpr int SC_PF_QUADS = SC_T / 64 + PL_PAD / 64;  // plane quads of one tile and of the PL_PAD positions before it
constexpr int SC_PF_BYTES = SC_PF_QUADS * 16;  // 4144
constexpr int SC_PF_STRIDE = 4224;             // slot stride (128-byte multiple)
constexpr int SC_DYN_SMEM = SC_STAGES * SC_T + SC_PSLOTS * SC_PF_STRIDE + 1024;  // + slack for the 1024-byte alignment of the TMA boxes
constexpr int SC_CAP = 512;                    // kept ENDs per list window (a typical tile lists ~300)

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
    // per warp, double buffered by tile parity (a tile without ENDs has no barrier (B)):
    uint32_t wsum[2][SC_WARPS];   // #kept ENDs | #ENDs << 16
    int wls[2][SC_WARPS], wlc[2][SC_WARPS];  // last START / last base-code change of the warp (tile-local position), -1 = none
    int wle[2][SC_WARPS];                    // last END of the warp
    int wfs[2][SC_WARPS], wfe[2][SC_WARPS];  // first START / END of the warp, NO_POS = none
    uint64_t wS[SC_WARPS];        // START word of the warp's last thread (min_len >= 3 only)
    // per tile, written by warp 0 between the barriers (A) and (B):
    uint64_t x_in;                // open-cluster state entering the tile
    uint64_t c_in;                // 1 + global position of the last base-code change before the tile (0: none in this chunk so far)
    uint64_t prefix;              // segment index of the tile's first record
    uint32_t adj;                 // 1: the carried END is not written by this tile
    uint32_t carried;             // the tile's first event is an END
    unsigned int hist[E2S_HIST_BINS];
    uint64_t smask[SC_THREADS];   // START words of the tile (only in the tile that holds position n_global - 2)
    uint32_t n2[2];               // entries of list2 (by tile parity: the other one is reset while this one is in use)
    uint32_t list2[SC_CAP];       // records of the window with a base-code change inside: (start - tile start + PL_PAD) | len << 16 (counted after the records are out)
    uint32_t e_ent[SC_CAP];       // [r] = kept END: position | len << 14 | survivor << 30, or position | change-seen << 30 | 1 << 31 (START before its warp)
};

constexpr int NO_POS = 0x7fffffff;

}  // namespace

// One CTA per CHUNK of consecutive tiles, no communication between CTAs: the open-cluster state and the record count are
// carried from tile to tile by warp 0, the records go to the chunk's own segment of the record arrays.  What a chunk
// cannot know -- whether a cluster is open when it starts -- only matters for its first event: if that is an END ("head"
// of the chunk) the record is left to k_chunk_resolve, which sees all chunks' summaries.
//
// Per tile: masks (S, E) and the change word C of my 64 positions; kept ENDs K = E minus the ENDs of clusters shorter than
// min_len (bit-parallel: a START at most min_len - 2 positions before); one block scan ranks the kept ENDs.  A cluster is
// [nearest START at or before its END, END]: each END finds that START in its own thread's word or, by one ballot + one
// shuffle, in an earlier lane of its warp -- and the last base-code change before the END the same way; only ENDs whose
// START lies before their warp (about one per warp) are resolved when the records are written, from the per-warp
// summaries.  BWT prefilter (fused mode): a cluster without a base-code change inside it has ONE base code and can never
// pass find_variants (ref:clust2snp.cpp:402-429: both samples' frequent sets would be that one letter): it is dropped
// here for the price of a compare; the others go to the list K3x tests exactly.
__global__ void __launch_bounds__(SC_THREADS, SC_OCC) k_cluster_scan(const __grid_constant__ CUtensorMap tmap, Scan8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* plane_slots = stages + size_t(SC_STAGES) * SC_T;  // SC_PSLOTS windows of SC_PF_STRIDE bytes
    __shared__ ScanShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
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
    auto issue = [&](uint32_t it) {  // thread 0: the it-th tile of the chunk -> stage it % SC_STAGES
        if (it >= n_my) return;
        const int stage = int(it % SC_STAGES);
        mbar_expect_tx(&sh.full_bar[stage], SC_T + (pf ? SC_PF_BYTES : 0));
        tma_load_2d_u8(stages + size_t(stage) * SC_T, &tmap, 0, int((t_lo + it) * (SC_T / 128)), &sh.full_bar[stage]);
        if (pf) bulk_g2s(plane_slots + size_t(it % SC_PSLOTS) * SC_PF_STRIDE, p.planes + uint64_t(t_lo + it) * (SC_T / 64), SC_PF_BYTES, &sh.full_bar[stage]);
    };
    if (tid == 0) {
        sh.n2[0] = sh.n2[1] = 0;
        for (int s = 0; s < SC_STAGES; ++s) mbar_init(&sh.full_bar[s], 1);
        fence_mbar_init();
        for (uint32_t s = 0; s < uint32_t(SC_STAGES); ++s) issue(s);
    }
    __syncthreads();

    // chunk state (warp 0; identical in its lanes)
    uint64_t X = OPEN_UNKNOWN;   // open-cluster state; unknown until the chunk's first event
    uint64_t c_last = 0;         // 1 + global position of the chunk's last base-code change so far
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

        // issued before waiting for the tile: the bytes around it and my word of the change plane
        uint32_t g_prev = 0, g_next = 0;
        if (tid == 0) g_prev = *reinterpret_cast<const uint32_t*>(p.lcp8 + (int64_t(tile_base) - 4));
        if (tid == SC_THREADS - 1) g_next = p.lcp8[tile_base + SC_T];
        uint64_t C = pf ? __ldg(p.chg + ((tile_base + uint64_t(tid) * SC_V + PL_PAD) >> 6)) : 0;

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
                C &= vm;
            }
            const uint64_t Gp = (G << 1) | g_m1b;
            const uint64_t Ep = (E << 1) | e_prev;
            S = G & (~Gp | Ep) & vm;
        }

        // ---- K = kept ENDs: E minus the ENDs of clusters shorter than min_len (a START at the same position or up to `spread`
        // positions before).  Before the tile: no START assumed -- the one END that can pair with a START of an earlier tile
        // is the tile's first event, which warp 0 tests exactly below.
        uint64_t K = E;
        if (spread >= 0) {
            uint64_t sm = S;
            if (spread >= 1) {  // (kernel-uniform) the previous lane's START word; across warps through shared memory
                if (lane == 31) sh.wS[warp] = S;
                __syncthreads();
                uint64_t lo = shfl64(S, (lane + 31) & 31);
                if (lane == 0) lo = warp ? sh.wS[warp - 1] : 0;
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
            K = E & ~sm;
        }

        // ---- ranks of the kept ENDs (one block scan of #kept | #ENDs << 16) and the per-warp summaries
        const uint32_t pk = uint32_t(__popcll(K)) | (uint32_t(__popcll(E)) << 16);
        uint32_t inc = pk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        const int my_ls = S ? tid * SC_V + 63 - __clzll(S) : -1;  // my last START / base-code change (tile-local)
        const int my_lc = C ? tid * SC_V + 63 - __clzll(C) : -1;
        int prev_s = -1, prev_c = -1;  // the nearest one in an EARLIER lane of my warp
        {
            const uint32_t ms = __ballot_sync(FULL, S != 0) & lt_mask, mc = __ballot_sync(FULL, C != 0) & lt_mask;
            const int vs = __shfl_sync(FULL, my_ls, ms ? 31 - __clz(ms) : lane);
            const int vc = __shfl_sync(FULL, my_lc, mc ? 31 - __clz(mc) : lane);
            if (ms) prev_s = vs;
            if (mc) prev_c = vc;
        }
        {
            const int ls = __reduce_max_sync(FULL, my_ls), lc = __reduce_max_sync(FULL, my_lc);
            const int le = __reduce_max_sync(FULL, E ? tid * SC_V + 63 - __clzll(E) : -1);
            const int fs = __reduce_min_sync(FULL, S ? tid * SC_V + __ffsll(S) - 1 : NO_POS);
            const int fe = __reduce_min_sync(FULL, E ? tid * SC_V + __ffsll(E) - 1 : NO_POS);
            if (lane == 31) sh.wsum[pb][warp] = inc;
            if (lane == 0) {
                sh.wls[pb][warp] = ls;
                sh.wlc[pb][warp] = lc;
                sh.wle[pb][warp] = le;
                sh.wfs[pb][warp] = fs;
                sh.wfe[pb][warp] = fe;
            }
        }
        __syncthreads();  // (A) per-warp summaries.  Every thread has this tile's bytes in registers: the stage is refilled
        if (tid == 0) {
            issue(it + SC_STAGES);
            sh.n2[pb ^ 1] = 0;  // (the previous tile's count: every thread is done with it)
        }

        uint32_t base = inc - pk, tot = 0;
#pragma unroll
        for (int q = 0; q < SC_WARPS; ++q) {
            const uint32_t ws = sh.wsum[pb][q];
            if (q < warp) base += ws;
            tot += ws;
        }
        const uint32_t nK = tot & 0xffffu, nE = tot >> 16;
        const uint32_t baseK = base & 0xffffu;

        // ---- warp 0: the tile's first / last events, the carried END, the chunk state
        if (warp == 0) {
            const int q8 = lane & (SC_WARPS - 1);
            const int t_ls = __reduce_max_sync(FULL, sh.wls[pb][q8]), t_le = __reduce_max_sync(FULL, sh.wle[pb][q8]);
            const int t_lc = __reduce_max_sync(FULL, sh.wlc[pb][q8]);
            const int t_fs = __reduce_min_sync(FULL, sh.wfs[pb][q8]), t_fe = __reduce_min_sync(FULL, sh.wfe[pb][q8]);
            const bool has_event = t_ls >= 0 || t_le >= 0;
            // the tile's first event is an END: its START lies before the tile (a START and an END at the same position: the
            // START comes first, t_fe == t_fs is not carried)
            const bool carried = t_fe != NO_POS && t_fe < t_fs;
            // The carried END is tested exactly on the wrapped length (ref:ebwt2clust.cpp:56,104) against the carried-in START: it is
            // the only END of the tile that can be 65 536 or more positions from its START.  Before the chunk's first event that
            // START is not known here: the END is the chunk's head, left to k_chunk_resolve.
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
            if (lane == 0) {
                sh.x_in = X;
                sh.c_in = c_last;
                sh.prefix = seg_base + cnt;
                sh.adj = adj | (is_head ? 2u : 0u);
                sh.carried = carried ? 1u : 0u;
            }
            cnt += nK - adj;
            n_end += nE;
            if (has_event) {
                X = t_ls > t_le ? tile_gbase + uint64_t(t_ls) + OPEN_BIAS : OPEN_NONE;
                seen = true;
            }
            if (t_lc >= 0) c_last = tile_gbase + uint64_t(t_lc) + 1;
        }

        // ---- the tile that holds position n_global - 2 (one per eBWT): an END there decides the reference's post-EOF phantom
        // value whether its record is kept or not -- its START, by a backward search in the tile's START words
        if (!interior && p.n_global - 2 - tile_gbase < uint64_t(SC_T)) {
            sh.smask[tid] = S;
            __syncthreads();
            const uint32_t e_loc = uint32_t(p.n_global - 2 - tile_gbase);
            if (uint32_t(tid) == (e_loc >> 6) && ((E >> (e_loc & 63)) & 1u)) {
                int wi = tid;
                uint64_t m = S & ((uint64_t(2) << (e_loc & 63)) - 1);
                while (!m && wi > 0) m = sh.smask[--wi];
                if (m) p.res->end_nm2_start = tile_gbase + uint64_t(wi) * SC_V + uint64_t(63 - __clzll(m)) + 1;
                else if (!(sh.adj & 2u)) p.res->end_nm2_start = sh.x_in >= OPEN_BIAS ? sh.x_in - OPEN_BIAS + 1 : ~0ull;
                // (the chunk's head: k_chunk_resolve answers)
            }
        }

        // ---- list of the kept ENDs + records, one window of SC_CAP at a time (one iteration unless the tile is unusually dense)
        for (uint32_t win = 0; win < nK; win += SC_CAP) {
            if (win) {  // the previous window's lists are no longer read
                __syncthreads();
                if (tid == 0) sh.n2[pb] = 0;
            }
            {
                uint64_t m = K;
                uint32_t r = baseK - win;  // (mod 2^32: ranks below the window fail the bound check)
                while (m) {
                    const int b = __ffsll(m) - 1;
                    m &= m - 1;
                    if (r < uint32_t(SC_CAP)) {
                        const uint64_t upto = (uint64_t(2) << b) - 1;  // bits 0 .. b
                        const uint64_t sb = S & upto, cb = C & upto;
                        const int s_loc = sb ? tid * SC_V + 63 - __clzll(sb) : prev_s;
                        const uint32_t e_loc = uint32_t(tid * SC_V + b);
                        uint32_t ent;
                        if (s_loc >= 0) {  // the cluster is [s_loc, e_loc]: a base-code change at a position in (s_loc, e_loc]?
                            const int c_loc = cb ? tid * SC_V + 63 - __clzll(cb) : prev_c;
                            ent = e_loc | ((e_loc - uint32_t(s_loc) + 1u) << 14) | (uint32_t(c_loc > s_loc) << 30);
                        } else {           // START before my warp: resolved when the record is written; a change seen here is inside
                            ent = e_loc | (uint32_t(cb != 0 || prev_c >= 0) << 30) | (1u << 31);
                        }
                        sh.e_ent[r] = ent;
                    }
                    ++r;
                }
            }
            __syncthreads();  // (B) list; warp 0's tile words
            const uint64_t x_in = sh.x_in, prefix = sh.prefix;
            const uint32_t adj = sh.adj & 1u;
            const bool carried = sh.carried != 0;
            const uint32_t n_win = nK - win < uint32_t(SC_CAP) ? nK - win : uint32_t(SC_CAP);
            for (uint32_t i = tid; i < n_win; i += SC_THREADS) {
                const uint32_t ent = sh.e_ent[i];
                const uint32_t e = ent & 0x3fffu;
                const uint64_t gend = tile_gbase + e;
                uint64_t st;
                uint32_t len;
                bool chg = (ent >> 30) & 1u, carried_end = false;
                if (!(ent >> 31)) {
                    len = (ent >> 14) & 0xffffu;
                    st = gend - len + 1;
                } else {  // nearest START / change in the warps before the END's
                    const int wq = int(e >> 11);
                    int s_loc = -1, c_loc = -1;
                    for (int q = wq - 1; q >= 0 && s_loc < 0; --q) s_loc = sh.wls[pb][q];
                    for (int q = wq - 1; q >= 0 && c_loc < 0; --q) c_loc = sh.wlc[pb][q];
                    if (s_loc >= 0) {
                        st = tile_gbase + uint64_t(s_loc);
                        len = e - uint32_t(s_loc) + 1u;
                        chg = chg || c_loc > s_loc;
                    } else {  // no START in the tile before it: the tile's carried END
                        carried_end = true;
                        if (sh.adj) continue;  // the chunk's head (k_chunk_resolve writes it) / dropped by the exact test
                        st = x_in - OPEN_BIAS;
                        len = uint32_t(gend - st + 1) & 0xffffu;
                        // a change in the tile before it, or after the START in an earlier tile; a wrapped length: the analysed
                        // range [st, st + len) is not this cluster's -- the exact test decides
                        chg = chg || c_loc >= 0 || sh.c_in > st + 1 || gend - st + 1 != uint64_t(len);
                    }
                }
                const uint64_t o = prefix + win + i - ((carried && !carried_end) ? adj : 0u);
                if (o - seg_base < p.seg_cap) {
                    p.seg_start[o] = st;
                    p.seg_len[o] = uint16_t(len);
                } else {
                    p.res->overflow |= 1;
                }
                acc_bases += len;
                if (len <= uint32_t(MAX_C_LEN)) atomicAdd(&sh.hist[len], 1u);
                my_last_o = o;
                my_last_len = len;
                // BWT prefilter, first level: no base-code change inside the cluster = one base code = it cannot pass.  The others are
                // counted below, densely; a record whose analysed range lies before the plane window (wrapped length) goes
                // straight to the exact test.
                if (pf && chg && len >= 2 * p.pf_mcov && len <= uint32_t(MAX_C_LEN)) {
                    if (st + PL_PAD < tile_gbase) {
                        const unsigned long long at = atomicAdd(&p.res->n_pf, 1ull);
                        if (at < p.pf_cap) p.pf_list[at] = SurvEntry{st, st - p.global_off, len, 0u};
                        else p.res->overflow |= 2;  // (seen by every rank in the exchange rows: all of them repeat the round)
                    } else {
                        sh.list2[atomicAdd(&sh.n2[pb], 1u)] = uint32_t(st + PL_PAD - tile_gbase) | (len << 16);
                    }
                }
            }
            if (pf) {  // second level: the one-popcount bound (planes.cuh) on the records that have a change inside, one lane each
                __syncthreads();
                const uint32_t n2 = sh.n2[pb];
                for (uint32_t j = tid; j < n2; j += SC_THREADS) {
                    const uint32_t v = sh.list2[j];
                    const uint32_t b_lo = v & 0xffffu, len = v >> 16, b_last = b_lo + len - 1;
                    const uint32_t q_lo = b_lo >> 6, q_last = b_last >> 6;
                    unsigned long long f0 = 0, f1 = 0;
                    uint32_t others = 0;
                    for (uint32_t q = q_lo; q <= q_last; ++q) {
                        const uint4 pv = pf_win[q];
                        const unsigned long long x0 = (uint64_t(pv.y) << 32) | pv.x, x1 = (uint64_t(pv.w) << 32) | pv.z;
                        unsigned long long mask = ~0ull;
                        if (q == q_lo) {
                            mask = ~0ull << (b_lo & 63);
                            f0 = 0ull - ((x0 >> (b_lo & 63)) & 1ull);
                            f1 = 0ull - ((x1 >> (b_lo & 63)) & 1ull);
                        }
                        if (q == q_last) mask &= ~0ull >> (63 - (b_last & 63));
                        others += __popcll(((x0 ^ f0) | (x1 ^ f1)) & mask);
                    }
                    if (others >= p.pf_mcov) {  // at least mcov records differ from the first one's base code: to the exact test
                        const uint64_t st = tile_gbase + b_lo - PL_PAD;
                        const unsigned long long at = atomicAdd(&p.res->n_pf, 1ull);
                        if (at < p.pf_cap) p.pf_list[at] = SurvEntry{st, st - p.global_off, len, 0u};
                        else p.res->overflow |= 2;  // (seen by every rank in the exchange rows: all of them repeat the round)
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
    if (tid == 0) {
        cr->own_count = cnt;
        cr->head_end = head_end;
        cr->last_state = seen ? (X >= OPEN_BIAS ? X : 1ull) : 0ull;  // 0: no event; 1: closed; >= 2: OPEN_BIAS + global START
        cr->n_end = n_end;
        sh.prefix = seg_base + cnt;  // (for the thread that wrote the chunk's last record)
        if (c == 0 && p.tail_lcp) {
            p.res->tail_lcp_nm2 = p.tail_lcp[0];
            p.res->tail_lcp_nm1 = p.tail_lcp[1];
            p.res->tail_bwt_nm1 = p.tail_bwt[0];
        }
    }
    __syncthreads();
    if (my_last_o + 1 == sh.prefix && sh.prefix != seg_base) cr->last_len = my_last_len;
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
                    else p.res->overflow |= 2;
                }
            }
        }
    }
    {   // block totals: warp sums by shuffles, one shared atomic per warp
        unsigned long long w0 = act ? mine : 0, w1 = act ? rec.n_end : 0, w2 = act ? rec.n_bases + (h_kept ? h_len : 0u) : 0;
        unsigned long long w3 = (act && mine) ? (unsigned long long)(c + 1) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            w0 += shfl64(w0, lane ^ d);
            w1 += shfl64(w1, lane ^ d);
            w2 += shfl64(w2, lane ^ d);
            const unsigned long long o = shfl64(w3, lane ^ d);
            w3 = o > w3 ? o : w3;
        }
        if (lane == 0) {
            if (w0) atomicAdd(&s_tot[0], w0);
            if (w1) atomicAdd(&s_tot[1], w1);
            if (w2) atomicAdd(&s_tot[2], w2);
            if (w3) atomicMax(&s_tot[3], w3);
        }
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
