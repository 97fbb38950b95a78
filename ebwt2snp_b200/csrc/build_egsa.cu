// EGSA construction on the GPU: eBWT + LCP + generalized suffix array of a collection of
// equal-length reads, the input the two tools expect an external `egsa` / BCR run to have produced
// (ref:README.md:46-60, ref:pipeline.sh:98-109).  Conventions = the ones this repo's synthetic data has used from the
// start (ebwt2snp_b200/synth.py; the reference pins none of them, SURVEY.md §8(b) last row):
//   one record per suffix of every read INCLUDING the terminator suffix (n = R (L + 1));
//   `$` < A < C < G < T; equal suffixes ordered by read id;
//   lcp[i] = common prefix with record i - 1, never extending over a terminator, lcp[0] = 0;
//   text = read id, suff = offset of the suffix in its read (L for the terminator suffix);
//   bwt = preceding character, `$` for whole-read suffixes.
//
// Suffixes of short reads are short strings, so there is no doubling: every suffix is a fixed-width key.
//   k_pack_reads   2 bits per base (A=0 C=1 G=2 T=3, most significant first), zero padded: packed[r][W + 1] u64 words
//   k_suffix_keys  key word w of suffix (r, p) = symbols [p + 32 w, p + 32 w + 32) of read r, by a funnel shift of two
//                  packed words.  Zero padding past the end of the read is enough to order a suffix before every longer
//                  one it prefixes IF ties are resolved shortest first: the sort is least-significant-word-first and
//                  stable, and its initial order is (offset descending, read id ascending).
//   radix sort     one stable pass per 64-bit key word, last word first (cub::DeviceRadixSort::SortPairs on exactly the
//                  significant bits of each word -- library code; the key extraction, the order trick and the finish
//                  kernel around it are this file)
//   k_finish       decode (r, p); text, suff, bwt; lcp with the previous record = min(L - p_a, L - p_b, first differing
//                  symbol) by clz on XORed key words.
// Outputs go straight to device arrays (a shard's resident SoA arrays via e2s_shard_load_soa_dev, or the caller's).

#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "internal.h"

namespace e2s {

__device__ __forceinline__ uint32_t code2(uint32_t c) {  // ACGT / acgt -> 0..3 (anything else 0: callers keep reads ACGT-only)
    const uint32_t u = c & 0xDFu;
    return uint32_t(u == 'C') + 2u * uint32_t(u == 'G') + 3u * uint32_t(u == 'T');
}

// one thread per (read, word)
// *bad is raised when a base is not one of ACGT / acgt: the 2-bit keys have no code for it (N would sort and compare as A)
__global__ void k_pack_reads(const uint8_t* __restrict__ reads, uint64_t R, uint32_t L, uint32_t W, uint64_t* __restrict__ packed,
                             uint32_t* __restrict__ bad) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= R * (W + 1)) return;
    const uint64_t r = i / (W + 1);
    const uint32_t w = uint32_t(i % (W + 1));
    uint64_t v = 0;
    bool other = false;
    for (uint32_t j = 0; j < 32; ++j) {
        const uint32_t s = 32 * w + j;
        v <<= 2;
        if (s < L) {
            const uint32_t c = reads[r * L + s], u = c & 0xDFu;
            other |= !(u == 'A' || u == 'C' || u == 'G' || u == 'T');
            v |= code2(c);
        }
    }
    packed[i] = v;
    if (other) *bad = 1u;
}

// symbols [p + 32 w, +32) of read r as one word (zero past the end of the read)
__device__ __forceinline__ uint64_t suffix_word(const uint64_t* __restrict__ packed, uint32_t W, uint64_t r, uint32_t p, uint32_t w) {
    const uint32_t s = p + 32 * w, q = s >> 5, sh = (s & 31) * 2;
    if (q > W) return 0;
    const uint64_t* row = packed + r * (W + 1);
    const uint64_t a = row[q], b = q < W ? row[q + 1] : 0;
    return sh ? (a << sh) | (b >> (64 - sh)) : a;
}

// suffix id = (L - p) * R + r: ascending ids = offset descending, then read id ascending (the sort's initial order)
__global__ void k_suffix_keys(const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ids, uint64_t n, uint64_t R, uint32_t L,
                              uint32_t W, uint32_t w, uint64_t* __restrict__ keys) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t id = ids ? ids[i] : i;
    const uint32_t p = L - uint32_t(id / R);
    keys[i] = suffix_word(packed, W, id % R, p, w);
}

__global__ void k_iota32(uint32_t* a, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) a[i] = uint32_t(i);
}

__global__ void k_egsa_finish(const uint8_t* __restrict__ reads, const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ids,
                              uint64_t n, uint64_t R, uint32_t L, uint32_t W, uint32_t* __restrict__ lcp, uint32_t* __restrict__ text,
                              uint32_t* __restrict__ suff, uint8_t* __restrict__ bwt) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t id = ids[i];
    const uint32_t p = L - uint32_t(id / R);
    const uint64_t r = id % R;
    text[i] = uint32_t(r);
    suff[i] = p;
    bwt[i] = p ? reads[r * L + p - 1] : uint8_t('$');
    uint32_t l = 0;
    if (i) {
        const uint64_t id0 = ids[i - 1];
        const uint32_t p0 = L - uint32_t(id0 / R);
        const uint64_t r0 = id0 % R;
        l = L - (p0 > p ? p0 : p);  // the shorter of the two suffixes
        for (uint32_t w = 0; 32 * w < l; ++w) {
            const uint64_t x = suffix_word(packed, W, r0, p0, w) ^ suffix_word(packed, W, r, p, w);
            if (x) {
                const uint32_t d = 32 * w + uint32_t(__clzll(x)) / 2;
                l = d < l ? d : l;
                break;
            }
        }
    }
    lcp[i] = l;
}

static inline unsigned blocks_for(uint64_t n, int t) { return unsigned((n + t - 1) / t); }

// scratch: packed reads, two key buffers, two id buffers, CUB temp storage -- allocated here, freed before returning
cudaError_t build_egsa(const uint8_t* d_reads, uint64_t R, uint32_t L, uint32_t* d_lcp, uint32_t* d_text, uint32_t* d_suff,
                       uint8_t* d_bwt, cudaStream_t stream, uint64_t* launches) {
    const uint64_t n = R * (uint64_t(L) + 1);
    const uint32_t W = (L + 31) / 32;
    uint64_t *packed = nullptr, *k0 = nullptr, *k1 = nullptr;
    uint32_t *i0 = nullptr, *i1 = nullptr;
    void* tmp = nullptr;
    uint32_t* d_bad = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaSuccess;
    auto done = [&](cudaError_t rc) {
        cudaFree(packed); cudaFree(k0); cudaFree(k1); cudaFree(i0); cudaFree(i1); cudaFree(tmp); cudaFree(d_bad);
        return rc;
    };
    if ((e = cudaMalloc(reinterpret_cast<void**>(&d_bad), 4)) != cudaSuccess) return done(e);
    if ((e = cudaMemsetAsync(d_bad, 0, 4, stream)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&packed), R * (W + 1) * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&k0), n * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&k1), n * 8)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&i0), n * 4)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&i1), n * 4)) != cudaSuccess) return done(e);
    cub::DoubleBuffer<uint64_t> keys(k0, k1);
    cub::DoubleBuffer<uint32_t> ids(i0, i1);
    if ((e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, ids, n, 0, 64, stream)) != cudaSuccess) return done(e);
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return done(e);

    k_pack_reads<<<blocks_for(R * (W + 1), 256), 256, 0, stream>>>(d_reads, R, L, W, packed, d_bad);
    {   // a base outside ACGT / acgt has no 2-bit code: refuse before sorting instead of building a wrong index
        uint32_t h_bad = 0;
        if ((e = cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return done(e);
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return done(e);
        if (h_bad) return done(cudaErrorInvalidValue);
    }
    k_iota32<<<blocks_for(n, 256), 256, 0, stream>>>(ids.Current(), n);
    *launches += 2;
    for (int w = int(W) - 1; w >= 0; --w) {
        k_suffix_keys<<<blocks_for(n, 256), 256, 0, stream>>>(packed, ids.Current(), n, R, L, W, uint32_t(w), keys.Current());
        // symbols of this word that exist in the longest suffix: only those bits can differ
        const uint32_t syms = L - 32 * uint32_t(w) < 32 ? L - 32 * uint32_t(w) : 32;
        const int begin_bit = int(64 - 2 * syms);
        if ((e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, ids, n, begin_bit, 64, stream)) != cudaSuccess) return done(e);
        *launches += 1 + (64 - begin_bit + 7) / 8;  // key kernel + about one sort kernel per digit (library launches, approximate)
    }
    k_egsa_finish<<<blocks_for(n, 256), 256, 0, stream>>>(d_reads, packed, ids.Current(), n, R, L, W, d_lcp, d_text, d_suff, d_bwt);
    *launches += 1;
    if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return done(e);
    return done(cudaSuccess);
}

}  // namespace e2s
