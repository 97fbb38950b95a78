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
constexpr int SC_PSLOTS = 4;                   // plane windows: a tile's window lives until its records are written, one tile later
constexpr int SC_OCC = 3;                      // resident CTAs per SM the register / shared-memory budget is sized for
constexpr int SC_PF_QUADS = SC_T / 64 + PL_PAD / 64;  // plane quads of one tile and of the PL_PAD positions before it
constexpr int SC_PF_BYTES = SC_PF_QUADS * 16;  // 4144
constexpr int SC_PF_STRIDE = 4224;             // slot stride (128-byte multiple)
constexpr int SC_DYN_SMEM = SC_STAGES * SC_T + SC_PSLOTS * SC_PF_STRIDE + 1024;  // + slack for the 1024-byte alignment of the TMA boxes
constexpr int SC_CAP = 1024;                   // ENDs per list window (a typical tile lists ~300)

// open-cluster state (as in cluster.cu)
constexpr uint64_t OPEN_NONE = 0, OPEN_UNKNOWN = 1, OPEN_BIAS = 2;

// descriptor words: [63:62] status  [61:20] payload  [19:0] epoch
constexpr uint64_t D_AGG = 1, D_INC = 2;
constexpr uint64_t D_EPOCH_MASK = (uint64_t(1) << 20) - 1;
constexpr uint64_t A_NONE = 0, A_CLOSED = 1, A_OPEN = 2, A_UNKNOWN = 3;  // payload bits [41:40] of A; [39:0] = local START
constexpr uint64_t A_POS_MASK = (uint64_t(1) << 40) - 1;

__device__ __forceinline__ uint64_t make_a(uint64_t status, uint64_t kind, uint64_t pos, uint32_t epoch) {
    return (status << 62) | (kind << 60) | ((pos & A_POS_MASK) << 20) | epoch;
}
__device__ __forceinline__ uint64_t make_b(uint64_t status, uint64_t count, uint32_t epoch) {
    return (status << 62) | (count << 20) | epoch;
}
__device__ __forceinline__ bool desc_valid(uint64_t v, uint32_t epoch) { return (v & D_EPOCH_MASK) == epoch && (v >> 62) != 0; }

__device__ __forceinline__ void desc_store_raw(uint64_t* d, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(d), "l"(v) : "memory");
}
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

// One look-back window: lane l polls desc[j0 - l] (tiles before tile 0 count as resolved) until the NEAREST resolved word
// (status INC) is known and every word nearer than it is valid.  true: lane `first` holds that word; false: all 32 words are
// valid aggregates (first = 32), go on with the next window.
__device__ __forceinline__ bool poll_window(const uint64_t* desc, int64_t j0, int lane, uint32_t epoch, uint64_t& w, int& first) {
    const int64_t j = j0 - lane;
    w = 0;
    bool ok = j < 0;
    while (true) {
        if (!ok) {
            w = desc_load(desc + j);
            ok = desc_valid(w, epoch);
        }
        const uint32_t okm = __ballot_sync(FULL, ok);
        const uint32_t stopm = __ballot_sync(FULL, ok && (j < 0 || (w >> 62) == D_INC));
        if (stopm) {
            first = __ffs(stopm) - 1;
            const uint32_t need = (1u << first) - 1u;
            if ((okm & need) == need) return true;
        } else if (okm == FULL) {
            first = 32;
            return false;
        }
        __nanosleep(20);
    }
}

struct TileMeta {       // what P2 needs to know about a tile whose lists were filled by P1
    uint64_t tile_gbase;
    uint64_t x_in;      // open state entering the tile
    uint64_t prefix;    // records kept before the tile (written by P2's look-back)
    uint32_t tile;      // tile number
    uint32_t nE;        // ENDs in the tile
    uint32_t count;     // records it keeps
    uint32_t adj;       // 1: the carried END is not written
    uint32_t carried;   // the tile's first event is an END whose START lies before the tile
    uint32_t pslot;     // plane-window slot
};

struct ScanShared {
    uint64_t full_bar[SC_STAGES];
    uint32_t tile_of[SC_STAGES];
    uint32_t wsum[2][SC_WARPS];   // per warp: #END | #dropped << 16 (double buffered by tile parity)
    int wls[2][SC_WARPS], wle[2][SC_WARPS], wfs[2][SC_WARPS], wfe[2][SC_WARPS];  // last / first START / END of the warp (tile-local)
    uint64_t wS[SC_WARPS];        // START word of the warp's last thread (min_len >= 3 only)
    TileMeta meta[2];
    unsigned int hist[E2S_HIST_BINS];
    uint16_t s_pos[2][SC_CAP];    // [slot][q] = tile-local position of the START that pairs with the window's q-th END
    uint32_t e_ent[2][SC_CAP];    // [slot][q] = END position | kept rank << 14 | kept << 28
};

constexpr int NO_POS = 0x7fffffff;

}  // namespace

// The tile loop is software-pipelined so that no CTA ever blocks while it still owes the grid an aggregate:
//   P1(tile)  wait for the tile, masks, ranks, lists; warp 0: publish A, resolve the state entering the tile, test the carried
//             END, publish the kept count (B aggregate)
//   P2(tile)  warp 0: look back over the B words -> records kept before the tile, publish B inclusive; all: write the records
// run as P1(t0) P1(t1) P2(t0) P1(t2) P2(t1) ...: the aggregate of a CTA's NEXT tile is out before the CTA waits for the
// prefix of its current one (with P1, P2 back to back a waiting CTA would hold up every tile after its next one).
__global__ void __launch_bounds__(SC_THREADS, SC_OCC) k_cluster_scan(const __grid_constant__ CUtensorMap tmap, Scan8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* plane_slots = stages + size_t(SC_STAGES) * SC_T;  // SC_PSLOTS windows of SC_PF_STRIDE bytes
    __shared__ ScanShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t num_tiles = p.num_tiles, epoch = p.epoch;
    const bool pf = p.pf_mcov != 0;
    const int spread = p.min_len >= 2 ? p.min_len - 2 : -1;  // extra positions a START shadows; -1: nothing is dropped
    const uint32_t kk = p.k > 128u ? 128u : p.k;             // bytes are <= 127: k >= 128 never matches
    const uint32_t kadd = (128u - kk) * 0x01010101u;

    for (int i = tid; i < E2S_HIST_BINS; i += SC_THREADS) sh.hist[i] = 0;
    // thread 0: the it-th tile of this CTA = next number from the ticket; its LCP box goes to stage it % SC_STAGES, its plane
    // window to slot it % SC_PSLOTS, both signalling the stage's barrier (a bare arrival when no tile is left)
    auto issue = [&](uint32_t it) {
        const int stage = int(it % SC_STAGES);
        const uint32_t t = uint32_t(atomicAdd(&p.res->ticket, 1ull));
        sh.tile_of[stage] = t < num_tiles ? t : 0xffffffffu;
        if (t < num_tiles) {
            mbar_expect_tx(&sh.full_bar[stage], SC_T + (pf ? SC_PF_BYTES : 0));
            tma_load_2d_u8(stages + size_t(stage) * SC_T, &tmap, 0, int(t * (SC_T / 128)), &sh.full_bar[stage]);
            if (pf) bulk_g2s(plane_slots + size_t(it % SC_PSLOTS) * SC_PF_STRIDE, p.planes + uint64_t(t) * (SC_T / 64), SC_PF_BYTES, &sh.full_bar[stage]);
        } else {
            mbar_arrive(&sh.full_bar[stage]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < SC_STAGES; ++s) mbar_init(&sh.full_bar[s], 1);
        fence_mbar_init();
        for (uint32_t s = 0; s < uint32_t(SC_STAGES); ++s) issue(s);
    }
    __syncthreads();

    unsigned long long acc_bases = 0, my_last = 0;  // sum of kept lengths; (index of my last record + 1) << 16 | its length
    unsigned long long cta_ends = 0;                // thread 0: ENDs of the tiles this CTA processed
    bool cta_any = false;                           // thread 0: one of them had an event

    // ---- P2: the records of one list window of the tile described by sh.meta[slot] (after a barrier that published lists + meta)
    auto emit_window = [&](int slot, uint32_t win) {
        const TileMeta& mt = sh.meta[slot];
        const uint64_t X = mt.x_in, prefix = mt.prefix, tile_gbase = mt.tile_gbase;
        const uint32_t adj = mt.adj, nE = mt.nE;
        const bool carried = mt.carried != 0;
        const uint4* pf_win = reinterpret_cast<const uint4*>(plane_slots + size_t(mt.pslot) * SC_PF_STRIDE);
        const uint32_t cnt = nE - win < uint32_t(SC_CAP) ? nE - win : uint32_t(SC_CAP);
        for (uint32_t i = tid; i < cnt; i += SC_THREADS) {
            const uint32_t ent = sh.e_ent[slot][i];
            const uint32_t e = ent & 0x3fffu, r = (ent >> 14) & 0x3fffu;
            const uint64_t gend = tile_gbase + e;
            const bool first_carried = carried && win + i == 0;
            uint64_t st;
            bool known = true;
            if (first_carried) {
                known = X >= OPEN_BIAS;
                st = X - OPEN_BIAS;
            } else {
                st = tile_gbase + sh.s_pos[slot][i];
            }
            if (gend + 2 == p.n_global) p.res->end_nm2_start = known ? st + 1 : ~0ull;  // decides the post-EOF phantom
            if (first_carried) {
                if (!known) {  // the shard's head END: its START is in an earlier shard
                    p.res->head_end = gend + 1;
                    continue;
                }
                if (adj) continue;
            } else if (!((ent >> 28) & 1u)) {
                continue;
            }
            const uint32_t len = uint32_t(gend - st + 1) & 0xffffu;
            const uint64_t o = prefix + r - ((carried && !first_carried) ? adj : 0u);
            if (o < p.cap) {
                p.out_start[o] = st;
                p.out_len[o] = uint16_t(len);
            } else {
                p.res->overflow = 1;
            }
            acc_bases += len;
            if (len <= uint32_t(MAX_C_LEN)) atomicAdd(&sh.hist[len], 1u);
            const unsigned long long mark = ((o + 1) << 16) | len;
            my_last = mark > my_last ? mark : my_last;
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
    };
    // warp 0: records kept before the tile of sh.meta[slot] (decoupled look-back over the B words), B inclusive published
    auto lookback_prefix = [&](int slot) {
        __syncwarp();  // (lane 0 may have written the meta words just now)
        TileMeta& mt = sh.meta[slot];
        const uint32_t t = mt.tile;
        uint64_t prefix = 0;
        if (t > 0) {
            int64_t j0 = int64_t(t) - 1;
            while (true) {
                uint64_t b;
                int first;
                const bool done = poll_window(p.descB, j0, lane, epoch, b, first);
                uint64_t v = (lane <= first && j0 - lane >= 0) ? (b >> 20) & ((uint64_t(1) << 42) - 1) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += shfl64(v, lane ^ o);
                prefix += v;
                if (done) break;
                j0 -= 32;
            }
        }
        if (lane == 0) {
            desc_store_raw(p.descB + t, make_b(D_INC, prefix + mt.count, epoch));
            mt.prefix = prefix;
            if (t == num_tiles - 1) p.res->n_written = prefix + mt.count;  // total of kept records
        }
    };

    bool pending = false;  // the previous tile's P2 is still to run (block-uniform)
    uint32_t it = 0;
    for (;; ++it) {
        const int stage = it % SC_STAGES;
        const uint32_t parity = (it / SC_STAGES) & 1;
        const int pb = it & 1;  // slot of the lists / meta / per-warp summaries
        const uint32_t t = sh.tile_of[stage];  // (written before a barrier every thread has passed since)
        if (t == 0xffffffffu) break;
        const uint64_t tile_base = uint64_t(t) * SC_T;
        const uint64_t tile_gbase = p.global_off + tile_base;
        const uint8_t* tile = stages + size_t(stage) * SC_T;
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
                const uint4 c = lds128(row + (((c0 + j) ^ x) << 4));
                w[4 * j + 0] = c.x;
                w[4 * j + 1] = c.y;
                w[4 * j + 2] = c.z;
                w[4 * j + 3] = c.w;
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
                __syncthreads();
                uint64_t lo = shfl64(S, (lane + 31) & 31);
                if (lane == 0) lo = warp ? sh.wS[warp - 1] : 0;  // before the tile: 0 (a carried END is tested exactly below)
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
        __syncthreads();  // (A) per-warp summaries.  Every thread has this tile's bytes in registers and is done with the tile
        // before the previous one (its plane window included): the stage is refilled with the CTA's tile it + SC_STAGES
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

        // ---- warp 0: state entering the tile, the carried END's fate, B aggregate -- everything the tiles after this one wait for
        if (warp == 0) {
            // A: the state after this tile is local whenever the tile has an event
            if (lane == 0 && has_event)
                desc_store_raw(p.descA + t, t_ls > t_le ? make_a(D_INC, A_OPEN, tile_base + uint64_t(t_ls), epoch) : make_a(D_INC, A_CLOSED, 0, epoch));
            else if (lane == 0)
                desc_store_raw(p.descA + t, make_a(D_AGG, A_NONE, 0, epoch));
            uint64_t X = p.global_off == 0 ? OPEN_NONE : OPEN_UNKNOWN;  // state before the shard
            if (t > 0 && (carried || !has_event)) {
                int64_t j0 = int64_t(t) - 1;
                while (true) {
                    uint64_t a;
                    int first;
                    if (poll_window(p.descA, j0, lane, epoch, a, first)) {
                        const uint64_t av = shfl64(a, first);
                        if (j0 - first >= 0) {
                            const uint64_t kind = (av >> 60) & 3u;
                            X = kind == A_OPEN ? OPEN_BIAS + p.global_off + ((av >> 20) & A_POS_MASK)
                                               : (kind == A_UNKNOWN ? OPEN_UNKNOWN : OPEN_NONE);
                        }
                        break;
                    }
                    j0 -= 32;
                }
            }
            if (lane == 0 && !has_event && t > 0)  // nothing happened here: pass the resolved state on
                desc_store_raw(p.descA + t, X >= OPEN_BIAS ? make_a(D_INC, A_OPEN, X - OPEN_BIAS - p.global_off, epoch)
                                                           : make_a(D_INC, X == OPEN_UNKNOWN ? A_UNKNOWN : A_CLOSED, 0, epoch));
            uint32_t adj = 0;
            if (carried) {  // exact test of append_entry on the wrapped length (ref:ebwt2clust.cpp:56,104); unknown START: the shard's head
                const uint64_t gend = tile_gbase + uint64_t(t_fe);
                adj = 1;
                if (X >= OPEN_BIAS) {
                    const uint32_t len = uint32_t(gend - (X - OPEN_BIAS) + 1) & 0xffffu;
                    adj = int(len) >= p.min_len ? 0u : 1u;
                }
            }
            if (lane == 0) {
                const uint32_t count = nE - nD - adj;
                if (t > 0) desc_store_raw(p.descB + t, make_b(D_AGG, count, epoch));
                TileMeta& mt = sh.meta[pb];
                mt.tile_gbase = tile_gbase;
                mt.x_in = X;
                mt.tile = t;
                mt.nE = nE;
                mt.count = count;
                mt.adj = adj;
                mt.carried = carried ? 1u : 0u;
                mt.pslot = it % SC_PSLOTS;
                cta_ends += nE;
                cta_any |= has_event;
                if (t == num_tiles - 1) {  // state after the whole shard
                    const uint64_t x_out = has_event ? (t_ls > t_le ? tile_gbase + uint64_t(t_ls) + OPEN_BIAS : OPEN_NONE) : X;
                    p.res->open_start = x_out >= OPEN_BIAS ? x_out - OPEN_BIAS + 1 : 0;
                }
            }
        }

        // ---- lists of the tile's first SC_CAP ENDs (all of them unless the tile is unusually dense)
        auto scatter = [&](uint32_t win) {
            uint64_t m = S;
            while (m) {  // a START goes to the slot of the END it pairs with: the one with as many ENDs before it
                const int b = __ffsll(m) - 1;
                m &= m - 1;
                const uint32_t q = baseE + __popcll(E & below(b)) - win;
                if (q < uint32_t(SC_CAP)) sh.s_pos[pb][q] = uint16_t(tid * SC_V + b);
            }
            m = E;
            while (m) {
                const int b = __ffsll(m) - 1;
                m &= m - 1;
                const uint32_t qa = baseE + __popcll(E & below(b));
                const uint32_t q = qa - win;
                if (q < uint32_t(SC_CAP)) {
                    const uint32_t r = qa - (baseD + __popcll(D & below(b)));
                    sh.e_ent[pb][q] = uint32_t(tid * SC_V + b) | (r << 14) | ((uint32_t((D >> b) & 1u) ^ 1u) << 28);
                }
            }
        };
        scatter(0);

        // ---- P2 of the previous tile: its aggregate and ours are out, now this CTA may wait
        if (pending) {
            if (warp == 0) lookback_prefix(pb ^ 1);
            __syncthreads();  // (B) prefix of the previous tile (its lists and meta were published by barrier (A) above)
            emit_window(pb ^ 1, 0);
        }
        pending = true;
        if (nE > uint32_t(SC_CAP)) {  // (block-uniform, rare) more ENDs than one list window holds: finish this tile here, window by window
            if (warp == 0) lookback_prefix(pb);
            __syncthreads();
            emit_window(pb, 0);
            for (uint32_t win = SC_CAP; win < nE; win += SC_CAP) {
                __syncthreads();  // the previous window's lists are no longer read
                scatter(win);
                __syncthreads();
                emit_window(pb, win);
            }
            pending = false;
        }
    }
    if (pending) {  // P2 of the CTA's last tile (the loop left at iteration `it`: that tile's slot is the other one)
        const int slot = int((it - 1) & 1);
        if (warp == 0) lookback_prefix(slot);
        __syncthreads();
        emit_window(slot, 0);
    }

    // ---- per-CTA totals
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc_bases += __shfl_xor_sync(FULL, acc_bases, d);
        const unsigned long long o = __shfl_xor_sync(FULL, my_last, d);
        my_last = o > my_last ? o : my_last;
    }
    if (lane == 0) {
        if (acc_bases) atomicAdd(&p.res->n_bases, acc_bases);
        if (my_last) atomicMax(&p.res->last_rec, my_last);
    }
    if (tid == 0) {
        if (cta_ends) atomicAdd(&p.res->n_end, cta_ends);
        if (cta_any) atomicOr(&p.res->any_event, 1ull);
    }
    __syncthreads();
    for (int i = tid; i < E2S_HIST_BINS; i += SC_THREADS)
        if (sh.hist[i]) atomicAdd(&p.res->hist[i], (unsigned long long)sh.hist[i]);
    if (blockIdx.x == 0 && tid == 0 && p.tail_lcp) {
        p.res->tail_lcp_nm2 = p.tail_lcp[0];
        p.res->tail_lcp_nm1 = p.tail_lcp[1];
        p.res->tail_bwt_nm1 = p.tail_bwt[0];
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

cudaError_t launch_scan(const Scan8Params& p0, uint64_t alloc_r, int sm_count, cudaStream_t stream) {
    PFN_encodeTiled enc = scan_encode_fn();
    if (!enc) return cudaErrorNotSupported;
    Scan8Params p = p0;
    p.num_tiles = uint32_t(scan_num_tiles(p.n_local));
    if (p.num_tiles == 0) return cudaSuccess;
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {128, cuuint64_t(alloc_r / 128)};  // the padded byte array from local position 0 on, as rows of 128 bytes
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {128, cuuint32_t(SC_T / 128)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(p.lcp8), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const size_t smem = size_t(SC_DYN_SMEM);
    static int occ_dev[64] = {0};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = occ_dev[dev & 63];
    if (!occ) {
        cudaError_t e = cudaFuncSetAttribute(k_cluster_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_cluster_scan, SC_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (o < 1) return cudaErrorLaunchOutOfResources;
        occ = o;
    }
    uint64_t grid = uint64_t(sm_count) * occ;
    if (grid > p.num_tiles) grid = p.num_tiles;
    k_cluster_scan<<<dim3(unsigned(grid)), dim3(SC_THREADS), smem, stream>>>(tmap, p);
    return cudaGetLastError();
}

}  // namespace e2s
