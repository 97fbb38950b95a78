// EGSA construction on the GPU: eBWT + LCP + generalized suffix array of a collection of
// reads, the input the two tools expect an external `egsa` / BCR run to have produced (ref:README.md:46-60,
// ref:pipeline.sh:98-109).  Reads may have any lengths (the reference accepts any FASTA, ref:clust2snp.cpp:147-212).
// Conventions = the ones this repo's synthetic data has used from the start (ebwt2snp_b200/synth.py; the reference pins
// none of them, SURVEY.md §8(b) last row):
//   one record per suffix of every read INCLUDING the terminator suffix (n = sum of (len_r + 1));
//   `$` < A < C < G < T; equal suffixes ordered by read id;
//   lcp[i] = common prefix with record i - 1, never extending over a terminator, lcp[0] = 0;
//   text = read id, suff = offset of the suffix in its read (len_r for the terminator suffix);
//   bwt = preceding character, `$` for whole-read suffixes.
//
// Suffixes of short reads are short strings, so there is no doubling: every suffix is a fixed-width key and the sort is
// a least-significant-digit radix sort written for exactly this key shape -- no library sort.
//   k_pack_reads   2 bits per base (A=0 C=1 G=2 T=3, most significant first), zero padded; read r's row starts at word
//                  (start_r >> 5) + 2 r (rows cannot overlap, no prefix sum over the reads needed)
//   k_init_ids     the initial order.  Equal lengths: id = (L - p) R + r, ascending = shortest suffix first, then read id.
//                  Ragged: id = (r << shift) | p in (read, offset) order, then one or two radix passes on the suffix
//                  LENGTH bring it to the same shortest-first order.
//   per key word w, last word first (zero padding orders a suffix before every longer one it prefixes because the sort is
//   stable and starts shortest-first):
//     k_keys_hist  symbols [p + 32 w, p + 32 w + 32) of every suffix as one 64-bit word (funnel shift of two packed
//                  words) in the current id order, AND the 256-bin histograms of every 8-bit digit of that word in the
//                  same pass (warp-aggregated shared-memory atomics: most suffixes are shorter than 32 w and share digit 0)
//     k_scan_hist  exclusive scans -> first output slot of every (digit place, bin)
//     k_radix_pass one launch per digit place that can differ (only the bits of symbols that exist in the longest read):
//                  tiles of 4096 pairs dealt by a ticket, stable warp-level multisplit (match.any), chained-scan
//                  decoupled look-back per bin (one 64-bit tagged word per tile and bin: no clearing between passes),
//                  keys then ids staged through one shared buffer so every bin leaves as a coalesced run.  The last place
//                  of a word moves the ids only.
//   k_egsa_finish  decode (r, p); text, suff, bwt; lcp with the previous record = min(len_a - p_a, len_b - p_b, first
//                  differing symbol) by clz on XORed key words.
// Ids are 32-bit when they fit (equal lengths: n < 2^32; ragged: R << shift <= 2^32), else 64-bit.
// Outputs go straight to device arrays (a shard's resident SoA arrays via e2s_shard_load_soa_dev, or the caller's).

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace e2s {

namespace {

constexpr uint32_t LEN_WORD = 0xffffffffu;  // k_keys_hist: key = suffix length instead of a symbol word

struct ReadsView {
    const uint8_t* bases;  // all reads, no separators
    const uint64_t* off;   // R + 1 offsets into bases, or nullptr: equal lengths (read r starts at r * L)
    uint64_t R;
    uint32_t L;      // the read length (equal) / the longest read (ragged)
    uint32_t shift;  // ragged ids: (r << shift) | p; 0 = equal-length ids (L - p) R + r
    __device__ __forceinline__ uint64_t start(uint64_t r) const { return off ? off[r] : r * L; }
    __device__ __forceinline__ uint32_t len(uint64_t r) const { return off ? uint32_t(off[r + 1] - off[r]) : L; }
    __device__ __forceinline__ uint64_t row(uint64_t r, uint64_t st) const { return (st >> 5) + 2 * r; }
    template <typename IdT>
    __device__ __forceinline__ void decode(IdT id, uint64_t& r, uint32_t& p) const {
        if (shift) {
            r = uint64_t(id) >> shift;
            p = uint32_t(id) & ((1u << shift) - 1u);
        } else if (sizeof(IdT) == 4) {
            const uint32_t q = uint32_t(id) / uint32_t(R);
            r = uint32_t(id) - q * uint32_t(R);
            p = L - q;
        } else {
            const uint64_t q = uint64_t(id) / R;
            r = uint64_t(id) - q * R;
            p = L - uint32_t(q);
        }
    }
};

__device__ __forceinline__ uint32_t code2(uint32_t c) {  // ACGT / acgt -> 0..3 (anything else is refused by k_pack_reads)
    const uint32_t u = c & 0xDFu;
    return uint32_t(u == 'C') + 2u * uint32_t(u == 'G') + 3u * uint32_t(u == 'T');
}

// one warp per read, lanes over the row's words (ceil(len / 32) + 1 of them, the last one all padding)
// *bad is raised when a base is not one of ACGT / acgt: the 2-bit keys have no code for it (N would sort and compare as A)
__global__ void k_pack_reads(ReadsView v, uint64_t* __restrict__ packed, uint32_t* __restrict__ bad) {
    const uint64_t r = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= v.R) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t st = v.start(r);
    const uint32_t len = v.len(r), W = (len + 31) >> 5;
    uint64_t* row = packed + v.row(r, st);
    bool other = false;
    for (uint32_t w = lane; w <= W; w += 32) {
        uint64_t x = 0;
        for (uint32_t j = 0; j < 32; ++j) {
            const uint32_t s = 32 * w + j;
            x <<= 2;
            if (s < len) {
                const uint32_t c = v.bases[st + s], u = c & 0xDFu;
                other |= !(u == 'A' || u == 'C' || u == 'G' || u == 'T');
                x |= code2(c);
            }
        }
        row[w] = x;
    }
    if (other) *bad = 1u;
}

// symbols [p + 32 w, +32) of a read as one word (zero past its end); row = the read's first packed word
__device__ __forceinline__ uint64_t suffix_word(const uint64_t* __restrict__ row, uint32_t len, uint32_t p, uint32_t w) {
    const uint32_t s = p + 32 * w;
    if (s >= len) return 0;
    const uint32_t q = s >> 5, sh = (s & 31) * 2;  // q < ceil(len / 32): row[q + 1] exists
    const uint64_t a = row[q];
    return sh ? (a << sh) | (row[q + 1] >> (64 - sh)) : a;
}

template <typename IdT>
__global__ void k_init_ids_equal(IdT* __restrict__ ids, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = IdT(i);
}

// ragged: one warp per read; suffix (r, p) sits at start_r + r + p in the initial order
template <typename IdT>
__global__ void k_init_ids_ragged(ReadsView v, IdT* __restrict__ ids) {
    const uint64_t r = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= v.R) return;
    const uint64_t st = v.start(r);
    const uint32_t len = v.len(r);
    for (uint32_t p = threadIdx.x & 31; p <= len; p += 32) ids[st + r + p] = IdT((r << v.shift) | p);
}

// ---- the radix sort ------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 16;                      // pairs per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;    // 4096 pairs per tile
constexpr int RS_WARP_SPAN = 32 * RS_IPT;       // a warp's contiguous piece of the tile
constexpr int MAX_PLACES = 8;                   // 8-bit digit places in a 64-bit word

// tile status word of the chained scan: [63:45] tag of the pass (never 0), [44] inclusive, [43:0] count
constexpr int TAG_SHIFT = 45;
constexpr uint64_t INCL_BIT = uint64_t(1) << 44;
constexpr uint64_t CNT_MASK = INCL_BIT - 1;

__device__ __forceinline__ void st_status(uint64_t* d, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(d), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_status(const uint64_t* d) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(d) : "memory");
    return v;
}

// keys of word w in the current id order + the histograms of its `places` digit places (place j = bits [begin_bit + 8 j, +8))
template <typename IdT>
__global__ void __launch_bounds__(256) k_keys_hist(ReadsView v, const uint64_t* __restrict__ packed, const IdT* __restrict__ ids,
                                                   uint64_t n, uint32_t w, uint32_t begin_bit, uint32_t places,
                                                   uint64_t* __restrict__ keys, unsigned long long* __restrict__ hist) {
    __shared__ uint32_t s_hist[MAX_PLACES][256];
    for (uint32_t b = threadIdx.x; b < MAX_PLACES * 256; b += blockDim.x) (&s_hist[0][0])[b] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t base = uint64_t(blockIdx.x) * blockDim.x; base < n; base += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const bool valid = i < n;
        uint64_t key = 0;
        if (valid) {
            uint64_t r;
            uint32_t p;
            v.decode(ids[i], r, p);
            const uint64_t st = v.start(r);
            const uint32_t len = v.len(r);
            key = w == LEN_WORD ? uint64_t(len - p) : suffix_word(packed + v.row(r, st), len, p, w);
            keys[i] = key;
        }
        for (uint32_t j = 0; j < places; ++j) {
            const uint32_t d = valid ? uint32_t(key >> (begin_bit + 8 * j)) & 0xffu : 256u + lane;
            const uint32_t m = __match_any_sync(0xffffffffu, d);
            if (valid && lane == uint32_t(__ffs(int(m)) - 1)) atomicAdd(&s_hist[j][d], uint32_t(__popc(m)));
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < places * 256; b += blockDim.x) {
        const uint32_t c = (&s_hist[0][0])[b];
        if (c) atomicAdd(hist + b, static_cast<unsigned long long>(c));
    }
}

// hist[place][bin] -> first output slot of the bin (exclusive scan per place); clears hist for the next word
__global__ void k_scan_hist(unsigned long long* __restrict__ hist, uint64_t* __restrict__ bin_base) {
    __shared__ uint64_t s[MAX_PLACES][256];
    for (uint32_t b = threadIdx.x; b < MAX_PLACES * 256; b += blockDim.x) {
        (&s[0][0])[b] = hist[b];
        hist[b] = 0;
    }
    __syncthreads();
    if (threadIdx.x < MAX_PLACES) {
        uint64_t sum = 0;
        for (int b = 0; b < 256; ++b) {
            const uint64_t t = s[threadIdx.x][b];
            s[threadIdx.x][b] = sum;
            sum += t;
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < MAX_PLACES * 256; b += blockDim.x) bin_base[b] = (&s[0][0])[b];
}

// one stable pass on the digit (key >> shift) & 255
template <typename IdT, bool WRITE_KEYS>
__global__ void __launch_bounds__(RS_THREADS, 3) k_radix_pass(const uint64_t* __restrict__ kin, const IdT* __restrict__ iin,
                                                           uint64_t* __restrict__ kout, IdT* __restrict__ iout, uint64_t n, uint32_t shift,
                                                           const uint64_t* __restrict__ bin_base, uint64_t* __restrict__ status,
                                                           uint32_t* __restrict__ ticket, uint32_t tag) {
    __shared__ uint64_t s_buf[RS_TILE];            // the tile's keys, then its ids, in output order
    __shared__ uint8_t s_digit[RS_TILE];           // digit of every staged slot
    __shared__ uint32_t s_whist[RS_WARPS][256];    // per-warp digit counts -> exclusive over the warps
    __shared__ uint32_t s_binstart[256];           // first staged slot of every bin
    __shared__ uint64_t s_gofs[256];               // output index of staged slot s of bin b = s_gofs[b] + s
    __shared__ uint32_t s_wsum[RS_WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);  // tiles in ticket order: every earlier tile is running or done
    for (uint32_t b = tid; b < RS_WARPS * 256; b += RS_THREADS) (&s_whist[0][0])[b] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = uint64_t(tile) * RS_TILE;
    const uint32_t tile_n = n - tile_base < uint64_t(RS_TILE) ? uint32_t(n - tile_base) : uint32_t(RS_TILE);
    const uint32_t first = warp * RS_WARP_SPAN + lane;  // item j of this thread = tile slot first + 32 j (warp-striped)

    uint64_t key[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const uint32_t li = first + 32 * j;
        key[j] = li < tile_n ? kin[tile_base + li] : ~uint64_t(0);
    }
    // rank of every item among the items of its warp with the same digit, in (item, lane) order
    uint16_t rnk[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const bool valid = first + 32 * j < tile_n;
        const uint32_t d = valid ? uint32_t(key[j] >> shift) & 0xffu : 256u + lane;
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        const uint32_t leader = uint32_t(__ffs(int(m)) - 1);
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = s_whist[warp][d];
            s_whist[warp][d] = prev + uint32_t(__popc(m));
        }
        prev = __shfl_sync(0xffffffffu, prev, int(leader));
        rnk[j] = uint16_t(prev + uint32_t(__popc(m & ((1u << lane) - 1u))));
        __syncwarp();
    }
    __syncthreads();

    // thread b owns bin b: counts exclusive over the warps, the tile's count published for the tiles behind
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t t = s_whist[w][tid];
        s_whist[w][tid] = cnt;
        cnt += t;
    }
    uint64_t* st = status + uint64_t(tile) * 256 + tid;
    const uint64_t tagw = uint64_t(tag) << TAG_SHIFT;
    st_status(st, tagw | (tile == 0 ? INCL_BIT : 0) | cnt);
    // first staged slot of the bin: block-wide exclusive scan of cnt
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= uint32_t(o)) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) wbase += uint32_t(w) < warp ? s_wsum[w] : 0u;
    const uint32_t binstart = wbase + incl - cnt;
    s_binstart[tid] = binstart;
    // chained scan: pairs of this bin in all earlier tiles
    uint64_t excl = 0;
    if (tile > 0) {
        for (int64_t t = int64_t(tile) - 1;; --t) {
            const uint64_t* pw = status + uint64_t(t) * 256 + tid;
            uint64_t wv;
            do {
                wv = ld_status(pw);
            } while ((wv >> TAG_SHIFT) != tag);
            excl += wv & CNT_MASK;
            if (wv & INCL_BIT) break;
        }
        st_status(st, tagw | INCL_BIT | (excl + cnt));
    }
    s_gofs[tid] = bin_base[tid] + excl - binstart;
    __syncthreads();

    // keys into output order (shared), then out as one coalesced run per bin
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        if (first + 32 * j < tile_n) {
            const uint32_t d = uint32_t(key[j] >> shift) & 0xffu;
            const uint32_t slot = s_binstart[d] + s_whist[warp][d] + rnk[j];
            rnk[j] = uint16_t(slot);
            s_digit[slot] = uint8_t(d);
            if (WRITE_KEYS) s_buf[slot] = key[j];
        }
    }
    __syncthreads();
    if (WRITE_KEYS) {
        for (uint32_t i = tid; i < tile_n; i += RS_THREADS) kout[s_gofs[s_digit[i]] + i] = s_buf[i];
        __syncthreads();
    }
    IdT* s_ids = reinterpret_cast<IdT*>(s_buf);
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const uint32_t li = first + 32 * j;
        if (li < tile_n) s_ids[rnk[j]] = iin[tile_base + li];
    }
    __syncthreads();
    for (uint32_t i = tid; i < tile_n; i += RS_THREADS) iout[s_gofs[s_digit[i]] + i] = s_ids[i];
}

template <typename IdT>
__global__ void k_egsa_finish(ReadsView v, const uint64_t* __restrict__ packed, const IdT* __restrict__ ids, uint64_t n,
                              uint32_t* __restrict__ lcp, uint32_t* __restrict__ text, uint32_t* __restrict__ suff,
                              uint8_t* __restrict__ bwt) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r;
    uint32_t p;
    v.decode(ids[i], r, p);
    const uint64_t st = v.start(r);
    const uint32_t len = v.len(r);
    text[i] = uint32_t(r);
    suff[i] = p;
    bwt[i] = p ? v.bases[st + p - 1] : uint8_t('$');
    uint32_t l = 0;
    if (i) {
        uint64_t r0;
        uint32_t p0;
        v.decode(ids[i - 1], r0, p0);
        const uint64_t st0 = v.start(r0);
        const uint32_t len0 = v.len(r0);
        const uint32_t la = len - p, lb = len0 - p0;
        l = la < lb ? la : lb;  // the shorter of the two suffixes
        const uint64_t* row = packed + v.row(r, st);
        const uint64_t* row0 = packed + v.row(r0, st0);
        for (uint32_t w = 0; 32 * w < l; ++w) {
            const uint64_t x = suffix_word(row0, len0, p0, w) ^ suffix_word(row, len, p, w);
            if (x) {
                const uint32_t d = 32 * w + uint32_t(__clzll(static_cast<long long>(x))) / 2;
                l = d < l ? d : l;
                break;
            }
        }
    }
    lcp[i] = l;
}

inline unsigned blocks_for(uint64_t n, int t) { return unsigned((n + uint64_t(t) - 1) / uint64_t(t)); }

template <typename IdT>
cudaError_t build_typed(ReadsView v, uint64_t n, uint64_t total_bases, uint32_t* d_lcp, uint32_t* d_text, uint32_t* d_suff,
                        uint8_t* d_bwt, cudaStream_t stream, uint64_t* launches) {
    const uint32_t W = (v.L + 31) / 32;
    const uint64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    const uint32_t len_places = v.shift ? (v.L >= 256 ? 2u : 1u) : 0u;
    uint32_t passes = len_places;
    for (uint32_t w = 0; w < W; ++w) {
        const uint32_t syms = v.L - 32 * w < 32 ? v.L - 32 * w : 32;
        passes += (2 * syms + 7) / 8;
    }
    if (passes >= (1u << (64 - TAG_SHIFT)) - 1u || tiles >= (uint64_t(1) << 31)) return cudaErrorInvalidConfiguration;
    const uint64_t packed_words = (total_bases >> 5) + 2 * v.R + 2;
    uint64_t *packed = nullptr, *k0 = nullptr, *k1 = nullptr, *status = nullptr, *misc = nullptr;
    IdT *i0 = nullptr, *i1 = nullptr;
    uint32_t* tickets = nullptr;
    cudaError_t e = cudaSuccess;
    auto done = [&](cudaError_t rc) {
        cudaFree(packed); cudaFree(k0); cudaFree(k1); cudaFree(i0); cudaFree(i1); cudaFree(status); cudaFree(misc); cudaFree(tickets);
        return rc;
    };
    // misc: hist[8][256] | bin_base[8][256] | bad flag
    const size_t misc_bytes = size_t(2 * MAX_PLACES * 256 + 1) * 8;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&misc), misc_bytes)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&tickets), size_t(passes + 1) * 4)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&status), tiles * 256 * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&packed), packed_words * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&k0), n * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&k1), n * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&i0), n * sizeof(IdT))) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&i1), n * sizeof(IdT))) != cudaSuccess) return done(e);
    if ((e = cudaMemsetAsync(misc, 0, misc_bytes, stream)) != cudaSuccess) return done(e);
    if ((e = cudaMemsetAsync(tickets, 0, size_t(passes + 1) * 4, stream)) != cudaSuccess) return done(e);
    if ((e = cudaMemsetAsync(status, 0, tiles * 256 * 8, stream)) != cudaSuccess) return done(e);  // tag 0 = no pass
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(misc);
    uint64_t* bin_base = misc + MAX_PLACES * 256;
    uint32_t* d_bad = reinterpret_cast<uint32_t*>(misc + 2 * MAX_PLACES * 256);

    k_pack_reads<<<blocks_for(v.R * 32, 256), 256, 0, stream>>>(v, packed, d_bad);
    *launches += 1;
    {   // a base outside ACGT / acgt has no 2-bit code: refuse before sorting instead of building a wrong index
        uint32_t h_bad = 0;
        if ((e = cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return done(e);
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return done(e);
        if (h_bad) return done(cudaErrorInvalidValue);
    }
    if (v.shift) k_init_ids_ragged<IdT><<<blocks_for(v.R * 32, 256), 256, 0, stream>>>(v, i0);
    else k_init_ids_equal<IdT><<<blocks_for(n, 256), 256, 0, stream>>>(i0, n);
    *launches += 1;

    uint64_t *kc = k0, *ka = k1;  // current / alternate
    IdT *ic = i0, *ia = i1;
    uint32_t seq = 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t hist_blocks_want = (n + 255) / 256;
    const unsigned hist_blocks = unsigned(hist_blocks_want < uint64_t(sms) * 8 ? hist_blocks_want : uint64_t(sms) * 8);
    auto sort_word = [&](uint32_t w, uint32_t begin_bit, uint32_t places) {
        k_keys_hist<IdT><<<hist_blocks, 256, 0, stream>>>(v, packed, ic, n, w, begin_bit, places, kc, hist);
        k_scan_hist<<<1, 256, 0, stream>>>(hist, bin_base);
        *launches += 2;
        for (uint32_t j = 0; j < places; ++j, ++seq) {
            const uint32_t sh = begin_bit + 8 * j;
            if (j + 1 < places)
                k_radix_pass<IdT, true><<<unsigned(tiles), RS_THREADS, 0, stream>>>(kc, ic, ka, ia, n, sh, bin_base + 256 * j, status,
                                                                                    tickets + seq, seq + 1);
            else  // the keys of this word are not looked at again
                k_radix_pass<IdT, false><<<unsigned(tiles), RS_THREADS, 0, stream>>>(kc, ic, ka, ia, n, sh, bin_base + 256 * j, status,
                                                                                     tickets + seq, seq + 1);
            *launches += 1;
            uint64_t* tk = kc; kc = ka; ka = tk;
            IdT* ti = ic; ic = ia; ia = ti;
        }
    };
    if (len_places) sort_word(LEN_WORD, 0, len_places);
    for (int w = int(W) - 1; w >= 0; --w) {
        // symbols of this word that exist in the longest suffix: only those bits can differ
        const uint32_t syms = v.L - 32 * uint32_t(w) < 32 ? v.L - 32 * uint32_t(w) : 32;
        sort_word(uint32_t(w), 64 - 2 * syms, (2 * syms + 7) / 8);
    }
    k_egsa_finish<IdT><<<blocks_for(n, 256), 256, 0, stream>>>(v, packed, ic, n, d_lcp, d_text, d_suff, d_bwt);
    *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return done(e);
    return done(cudaSuccess);
}

}  // namespace

// d_off == nullptr: R reads of L bases each; else R + 1 DEVICE offsets, L = the longest read, total_bases = off[R].
// scratch (freed before returning): two key buffers + two id buffers + tile status = 24.5 (32-bit ids) / 32.5 bytes per suffix
cudaError_t build_egsa(const uint8_t* d_reads, const uint64_t* d_off, uint64_t R, uint32_t L, uint64_t total_bases, uint32_t* d_lcp,
                       uint32_t* d_text, uint32_t* d_suff, uint8_t* d_bwt, cudaStream_t stream, uint64_t* launches) {
    ReadsView v;
    v.bases = d_reads;
    v.off = d_off;
    v.R = R;
    v.L = L;
    v.shift = 0;
    const uint64_t n = total_bases + R;
    bool wide;
    if (d_off) {
        if (L >= 65536) return cudaErrorInvalidConfiguration;
        uint32_t bits = 1;
        while ((L >> bits) != 0) ++bits;  // p in [0, L]
        v.shift = bits;
        wide = R > (uint64_t(1) << (32 - bits));
    } else {
        wide = n > 0xffffffffull;
    }
    if (getenv("E2S_BUILD_IDS64")) wide = true;  // test hook: the 64-bit instantiation on small inputs
    return wide ? build_typed<uint64_t>(v, n, total_bases, d_lcp, d_text, d_suff, d_bwt, stream, launches)
                : build_typed<uint32_t>(v, n, total_bases, d_lcp, d_text, d_suff, d_bwt, stream, launches);
}

}  // namespace e2s
