// Phase 2 (clust2snp): per-cluster analysis and SNP/indel calling.
//
//   k_len_hist       statistics(): length histogram                      ref:clust2snp.cpp:889-909
//   k_code_scan      K3a: exact prefilter of find_variants on the BWT base codes alone.  A cluster whose
//                    records show at most ONE base code (base_to_int, ref:include.hpp:265-279) with
//                    >= mcov_out occurrences over both samples cannot pass ref:clust2snp.cpp:402-429:
//                    counts[s][c] <= total[c] < mcov_out for every other code, so both frequent sets are
//                    subsets of one letter and are either empty or equal.  Only clusters with >= 2
//                    frequent codes (true variants, repeats) go on to the exact test.
//   k_cluster_exact  K3x: per-cluster 2x4 nucleotide histogram (sample = gSA text id < nreads1),
//                    first-argmax LCP and the find_variants filters        ref:clust2snp.cpp:377-429
//   k_candidates     K3b: ordered (ballot/popc) selection of the first <= c supporting reads per
//                    sample and allele pair                                ref:clust2snp.cpp:431-496
//   k_events         K4: gSA-driven gather of read contexts, consensus, support, distance()
//                                                  ref:clust2snp.cpp:541-624, 254-302, include.hpp:334-371
//
// K3a is the only phase-2 kernel that touches every position: one thread per cluster record answers the record with
// range popcounts on the shard's resident bit planes (16-byte loads of 64 positions x 2 planes, shared between
// neighbouring threads through L1).  Traffic: 0.25 B/position + 10 B/cluster.
// It does not run when K2 already applied the same prefilter while writing the records (fused mode,
// e2s_cluster_prefilter): snp_run then starts from K2's survivor list.
// The whole phase is enqueued with device-resident counts and synchronises once (snp_run).

#include <cuda_runtime.h>

#include <algorithm>
#include <vector>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "internal.h"
#include "planes.cuh"

namespace e2s {

constexpr uint32_t FULL = 0xffffffffu;

// ref:include.hpp:265-279: ACGT/acgt -> 0..3, everything else (incl. '$') -> 0.  N/n is flagged.
__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    const uint32_t u = c & 0xDFu;
    return uint32_t(u == 'C') + 2u * uint32_t(u == 'G') + 3u * uint32_t(u == 'T');
}
__device__ __forceinline__ bool is_n(uint32_t c) { return (c & 0xDFu) == 'N'; }

// ---------------------------------------------------------------------------------------------
// statistics
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_len_hist(const uint16_t* __restrict__ len, uint64_t m, unsigned long long* out) {
    __shared__ unsigned int h[MAX_C_LEN + 1];
    for (int i = threadIdx.x; i <= MAX_C_LEN; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned long long bases = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += uint64_t(gridDim.x) * blockDim.x) {
        uint32_t l = len[i];
        bases += l;
        if (l <= MAX_C_LEN) atomicAdd(&h[l], 1u);
    }
    for (int d = 16; d > 0; d >>= 1) bases += __shfl_xor_sync(FULL, bases, d);
    if ((threadIdx.x & 31) == 0 && bases) atomicAdd(&out[MAX_C_LEN + 1], bases);
    __syncthreads();
    for (int i = threadIdx.x; i <= MAX_C_LEN; i += blockDim.x)
        if (h[i]) atomicAdd(&out[i], (unsigned long long)h[i]);
}

cudaError_t launch_len_hist(const uint16_t* len, uint64_t m, unsigned long long* hist, cudaStream_t stream, int sm_count) {
    if (m == 0) return cudaSuccess;
    uint64_t blocks = (m + 256 * 16 - 1) / (256 * 16);
    if (blocks > uint64_t(sm_count) * 8) blocks = uint64_t(sm_count) * 8;
    k_len_hist<<<unsigned(blocks), 256, 0, stream>>>(len, m, hist);
    return cudaGetLastError();
}

__global__ void k_check_sorted(const uint64_t* __restrict__ start, const uint16_t* __restrict__ len, uint64_t m, SnpDev* dev) {
    bool bad = false;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i + 1 < m; i += uint64_t(gridDim.x) * blockDim.x)
        bad |= start[i] + len[i] > start[i + 1];
    if (bad) dev->unsorted = 1;
}

cudaError_t launch_check_sorted(const uint64_t* start, const uint16_t* len, uint64_t m, SnpDev* dev, cudaStream_t stream,
                                int sm_count) {
    if (m < 2) return cudaSuccess;
    uint64_t blocks = (m + 256 * 8 - 1) / (256 * 8);
    if (blocks > uint64_t(sm_count) * 8) blocks = uint64_t(sm_count) * 8;
    k_check_sorted<<<unsigned(blocks), 256, 0, stream>>>(start, len, m, dev);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K3a: base-code prefilter on the resident bit planes
// ---------------------------------------------------------------------------------------------
constexpr int PS_THREADS = 256;
constexpr int PS_T = 16384;  // granularity of the analysed range: records starting before ceil(n_local / PS_T) * PS_T

struct ScanParams {
    SnpArrays a;
    uint64_t limit;              // records with global_off <= start < global_off + limit are analysed
    uint32_t min_len, max_len;   // 2*mcov_out, max_clust_length
    uint32_t mcov;
    SurvEntry* survivors;        // out: clusters that need the exact test (unordered; dev->n_survivors counts them)
    uint64_t cap_surv;
    SnpDev* dev;
};

// One thread per cluster record: total count of each base code by range popcounts on the planes (0.25 B/position,
// 16-byte loads that neighbouring threads share through L1) + the 10-byte record.
constexpr int PS_U = 4;  // records per thread in flight: the record loads, then the plane loads they address, overlap

__global__ void __launch_bounds__(PS_THREADS) k_code_scan(ScanParams p) {
    unsigned long long n_analysed = 0;
    const uint64_t stride = uint64_t(gridDim.x) * PS_THREADS * PS_U;
    for (uint64_t c0 = uint64_t(blockIdx.x) * PS_THREADS * PS_U + threadIdx.x; c0 < p.a.m; c0 += stride) {
        uint64_t st[PS_U];
        uint32_t ln[PS_U];
        uint4 q0[PS_U];  // first plane quad of each record (most clusters need one or two)
#pragma unroll
        for (int u = 0; u < PS_U; ++u) {
            const uint64_t c = c0 + uint64_t(u) * PS_THREADS;
            st[u] = c < p.a.m ? p.a.cl_start[c] : 0;
            ln[u] = c < p.a.m ? p.a.cl_len[c] : 0;  // length 0 is never analysed
        }
        bool take[PS_U];
#pragma unroll
        for (int u = 0; u < PS_U; ++u) {
            take[u] = ln[u] >= p.min_len && ln[u] <= p.max_len && st[u] - p.a.global_off < p.limit;
            q0[u] = take[u] ? __ldg(p.a.planes + ((st[u] - p.a.global_off + PL_PAD) >> 6)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < PS_U; ++u) {
            if (!take[u]) continue;
            ++n_analysed;
            if (frequent_codes<true>(p.a.planes, int64_t(st[u] - p.a.global_off), ln[u], p.mcov, &q0[u]) >= 2) {
                // rare (variants, repeats): plain atomic append, the exact test does not need an order
                const unsigned long long at = atomicAdd(&p.dev->n_survivors, 1ull);
                if (at < p.cap_surv) p.survivors[at] = SurvEntry{st[u], st[u] - p.a.global_off, ln[u], 0u};
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_analysed += __shfl_xor_sync(FULL, n_analysed, d);
    if ((threadIdx.x & 31) == 0 && n_analysed) atomicAdd(&p.dev->n_analysed, n_analysed);
}

// K3a alone (streaming clust2snp: the staged records of one chunk -> the list k_capture reads); dev->n_survivors / n_analysed count
cudaError_t launch_code_scan(const SnpArrays& a, uint32_t mcov, uint32_t max_len, SurvEntry* out, uint64_t cap, SnpDev* dev,
                             cudaStream_t stream, int sm_count) {
    if (!a.m) return cudaSuccess;
    ScanParams sp;
    sp.a = a;
    sp.limit = (a.n_local + PS_T - 1) / PS_T * PS_T;
    sp.min_len = 2 * mcov;
    sp.max_len = max_len;
    sp.mcov = mcov;
    sp.survivors = out;
    sp.cap_surv = cap;
    sp.dev = dev;
    uint64_t grid = (a.m + PS_THREADS * PS_U - 1) / (PS_THREADS * PS_U);
    if (grid > uint64_t(sm_count) * 16) grid = uint64_t(sm_count) * 16;
    k_code_scan<<<unsigned(grid), PS_THREADS, 0, stream>>>(sp);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// streaming: the surviving clusters' records -> compact payload (one warp per cluster)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_capture(CaptureParams p) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_avail = *p.n_in;
    const uint64_t n = n_avail < p.n_in_cap ? n_avail : p.n_in_cap;
    const uint64_t warps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t i = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const SurvEntry e = p.in[i];
        const int64_t local = int64_t(e.base);  // (relative to the chunk's local position 0)
        if (local < -int64_t(PAD_L)) {  // a cluster whose analysed range left the device with an earlier chunk (wrapped 16-bit length)
            if (lane == 0) atomicOr(&p.counters[2], 1ull);
            continue;
        }
        unsigned long long off = 0, slot = 0;
        if (lane == 0) {
            off = atomicAdd(&p.counters[0], (unsigned long long)e.len);
            slot = atomicAdd(&p.counters[1], 1ull);
        }
        off = __shfl_sync(FULL, off, 0);
        slot = __shfl_sync(FULL, slot, 0);
        if (off + e.len > p.pay_cap || slot >= p.out_cap) {
            if (lane == 0) atomicOr(&p.counters[2], off + e.len > p.pay_cap ? 2ull : 4ull);
            continue;
        }
        for (uint32_t j = lane; j < e.len; j += 32) {
            p.p_lcp[off + j] = p.lcp[local + j];
            p.p_text[off + j] = p.text[local + j];
            p.p_suff[off + j] = p.suff[local + j];
            p.p_bwt[off + j] = p.bwt[local + j];
        }
        if (lane == 0) p.out[slot] = SurvEntry{e.start, off, e.len, 0u};
    }
}

cudaError_t launch_capture(const CaptureParams& p, cudaStream_t stream, int sm_count) {
    k_capture<<<unsigned(sm_count) * 4, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K3x: exact find_variants filters for the clusters that survived the prefilter
// ---------------------------------------------------------------------------------------------
constexpr int EX_THREADS = 256;
constexpr int EX_G = 8;  // lanes per cluster

struct ExactParams {
    SnpArrays a;
    const SurvEntry* list;     // clusters to test
    const unsigned long long* n_list;  // device-resident length of the list; null: n_list_host
    uint64_t n_list_host;
    uint64_t cap_list;         // capacity of the list (the pass is repeated when it was too small)
    uint32_t min_len, max_len; // 2 * mcov_out, max_clust_length (a fused prefilter only knew max_len <= 150)
    const int32_t* d_max_len;  // != null: max_clust_length comes from device memory (the merge kernel wrote it)
    uint32_t mcov, k_right;
    uint32_t nr1_lo, nr1_big;
    SurvEntry* flagged;        // out: clusters that pass the find_variants filters, unordered (dev->n_flagged counts them)
    uint64_t cap_flagged;
    SnpDev* dev;
};

__global__ void __launch_bounds__(EX_THREADS) k_cluster_exact(ExactParams p) {
    // Appending to the flagged list: ONE atomic per block and round (tens of thousands of same-address atomics, one per
    // cluster, were what the kernel's time consisted of: ~2 ns each, serialised in L2)
    __shared__ uint32_t s_wcnt[EX_THREADS / 32];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (EX_G - 1);
    const uint32_t gmask = ((1u << EX_G) - 1u) << (lane & ~(EX_G - 1));
    const uint64_t n_avail = p.n_list ? *p.n_list : p.n_list_host;
    const uint64_t n_list = n_avail < p.cap_list ? n_avail : p.cap_list;
    const uint64_t groups = uint64_t(gridDim.x) * (EX_THREADS / EX_G);
    const uint32_t max_len = p.d_max_len ? uint32_t(*p.d_max_len) : p.max_len;
    uint32_t saw_n = 0;
    for (uint64_t i0 = uint64_t(blockIdx.x) * (EX_THREADS / EX_G); i0 < n_list; i0 += groups) {  // block-uniform trip count
        const uint64_t i = i0 + uint64_t(threadIdx.x / EX_G);
        bool ok = false;
        SurvEntry fe;
        fe.start = fe.base = 0;
        fe.len = fe.pad = 0;
        if (i < n_list) {
            const SurvEntry ent = p.list[i];
            const uint64_t lp = ent.base;
            const uint32_t len = ent.len;
            // (group-uniform conditions) the analysed lengths; where the shard's bit planes are at hand the one-popcount bound
            // (planes.cuh) once more for the records the scan could not test (adopted ones, ranges before a plane window)
            if (len >= p.min_len && len <= max_len &&
                !(p.a.planes && lp + len <= p.a.n_local + MAX_C_LEN + 1 && !frequent_bound(p.a.planes, int64_t(lp), len, p.mcov))) {
                unsigned long long acc = 0, best = 0;
#pragma unroll 4
                for (uint32_t j = gl; j < len; j += EX_G) {  // (unrolled: the loads of four rounds are in flight together)
                    const uint32_t tx = p.a.text[lp + j];
                    const uint32_t lc = p.a.lcp[lp + j];
                    const uint32_t b = p.a.bwt[lp + j];
                    const uint32_t sample = (tx >= p.nr1_lo) & (p.nr1_big ^ 1u);
                    saw_n |= is_n(b);
                    acc += 1ull << (8 * (sample * 4 + base_code(b)));
                    const unsigned long long key = (uint64_t(lc) << 8) | (255u - j);
                    best = key > best ? key : best;
                }
#pragma unroll
                for (int d = EX_G / 2; d > 0; d >>= 1) {
                    acc += __shfl_xor_sync(gmask, acc, d);
                    unsigned long long o = __shfl_xor_sync(gmask, best, d);
                    best = o > best ? o : best;
                }
                if (gl == 0 && (best >> 8) >= p.k_right) {
                    uint32_t f0 = 0, f1 = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        f0 |= uint32_t(((acc >> (8 * b)) & 0xff) >= p.mcov) << b;
                        f1 |= uint32_t(((acc >> (8 * (b + 4))) & 0xff) >= p.mcov) << b;
                    }
                    ok = f0 && f1 && __popc(f0) <= 2 && __popc(f1) <= 2 && f0 != f1 && __popc(f0 | f1) <= 3;
                    fe = ent;
                    // what K3b would otherwise recompute: the frequent codes of both samples and the first strict LCP maximum
                    fe.pad = f0 | (f1 << 4) | ((255u - uint32_t(best & 0xff)) << 8) | (1u << 31);
                }
            }
        }
        const uint32_t okm = __ballot_sync(FULL, ok);
        if (lane == 0) s_wcnt[warp] = uint32_t(__popc(okm));
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int q = 0; q < EX_THREADS / 32; ++q) tot += s_wcnt[q];
            s_base = tot ? atomicAdd(&p.dev->n_flagged, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (ok) {
            unsigned long long at = s_base + uint32_t(__popc(okm & ((1u << lane) - 1u)));
            for (int q = 0; q < warp; ++q) at += s_wcnt[q];
            if (at < p.cap_flagged) p.flagged[at] = fe;  // the host compares the count with the capacity and retries
        }
        __syncthreads();  // (s_wcnt / s_base are rewritten by the next round)
    }
    if (__any_sync(__activemask(), saw_n) && saw_n) atomicOr(&p.dev->saw_n, 1ull);
}

// ---------------------------------------------------------------------------------------------
// K3b: candidates of the flagged clusters
// ---------------------------------------------------------------------------------------------
struct CandParams {
    SnpArrays a;
    const SurvEntry* flagged;  // clusters that passed the exact test
    const unsigned long long* n_flagged;  // device-resident length of the list
    uint32_t mcov, k_left, k_right, cap;  // cap = min(consensus_reads, 150)
    uint32_t nr1_lo, nr1_big;
    CandSlot* slots;       // 4 per flagged cluster
    uint32_t* slot_text;   // [slot][2][cap]
    uint32_t* slot_pos;    // [slot][2][cap]
    uint64_t* cand;        // out: valid slots, unordered (dev->n_slots_valid counts them)
    SnpDev* dev;
    uint64_t cap_flagged;  // capacity of the flagged list / slot arrays
};

constexpr int CA_WARPS = 8;

__global__ void __launch_bounds__(CA_WARPS * 32) k_candidates(CandParams p) {
    __shared__ uint32_t s_wcnt[CA_WARPS];  // valid slots per warp: one atomic per block and round appends them all
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t n_f = *p.n_flagged;
    if (n_f > p.cap_flagged) n_f = p.cap_flagged;  // overflow is reported by the host, which retries with more room
    const uint64_t warps = uint64_t(gridDim.x) * CA_WARPS;
    for (uint64_t f0 = uint64_t(blockIdx.x) * CA_WARPS; f0 < n_f; f0 += warps) {  // block-uniform trip count
    const uint64_t f = f0 + uint64_t(warp);
    uint32_t vmask = 0;  // valid slots of my cluster
    if (f < n_f) {
    const SurvEntry ent = p.flagged[f];
    const uint64_t start = ent.start;
    const uint32_t len = ent.len;
    const uint64_t lp = ent.base;  // first record of the cluster in the arrays
    const uint32_t* text = p.a.text + lp;
    const uint32_t* lcp = p.a.lcp + lp;
    const uint32_t* suff = p.a.suff + lp;
    const uint8_t* bwt = p.a.bwt + lp;

    uint32_t fr[2][2], nf[2] = {0, 0};
    uint32_t jbest;
    if (ent.pad >> 31) {  // K3x left the frequent codes (ascending, at most two per sample) and the position of the LCP maximum
        for (int s = 0; s < 2; ++s) {
            uint32_t f = (ent.pad >> (4 * s)) & 0xfu;
            while (f && nf[s] < 2) {
                fr[s][nf[s]++] = uint32_t(__ffs(f) - 1);
                f &= f - 1;
            }
        }
        jbest = (ent.pad >> 8) & 0xffu;
    } else {
        unsigned long long acc = 0, best = 0;
        for (uint32_t j = lane; j < len; j += 32) {
            const uint32_t sample = (text[j] >= p.nr1_lo) & (p.nr1_big ^ 1u);
            acc += 1ull << (8 * (sample * 4 + base_code(bwt[j])));
            const unsigned long long key = (uint64_t(lcp[j]) << 8) | (255u - j);
            best = key > best ? key : best;
        }
        for (int d = 16; d > 0; d >>= 1) {
            acc += __shfl_xor_sync(FULL, acc, d);
            unsigned long long o = __shfl_xor_sync(FULL, best, d);
            best = o > best ? o : best;
        }
        for (int s = 0; s < 2; ++s)
            for (int b = 0; b < 4; ++b)
                if (((acc >> (8 * (s * 4 + b))) & 0xff) >= p.mcov && nf[s] < 2) fr[s][nf[s]++] = b;
        jbest = 255u - uint32_t(best & 0xff);
    }
    const uint64_t right_idx = text[jbest], right_pos = suff[jbest];
    // the cluster's records, once, all loads in flight together (a flagged cluster is at most MAX_C_LEN = 150 long: 5 rounds of 32)
    constexpr int CA_R = (MAX_C_LEN + 31) / 32;
    uint32_t rtx[CA_R], rsf[CA_R], rch[CA_R];
    bool rcommon[CA_R];
    {
        uint32_t rlc[CA_R];
#pragma unroll
        for (int it = 0; it < CA_R; ++it) {
            const uint32_t j = uint32_t(it) * 32u + uint32_t(lane);
            const bool in = j < len;
            rtx[it] = in ? text[j] : 0u;
            rsf[it] = in ? suff[j] : 0u;
            rch[it] = in ? uint32_t(bwt[j]) : 0u;
            rlc[it] = in ? lcp[j] : 0u;
        }
#pragma unroll
        for (int it = 0; it < CA_R; ++it)
            rcommon[it] = uint32_t(it) * 32u + uint32_t(lane) < len && rsf[it] >= p.k_left && rlc[it] >= p.k_right;
    }

    for (uint32_t i0 = 0; i0 < 2; ++i0) {
        for (uint32_t i1 = 0; i1 < 2; ++i1) {
            const uint64_t slot = f * 4 + i0 * 2 + i1;
            CandSlot hdr;
            hdr.n0 = hdr.n1 = hdr.valid = hdr.pad = 0;
            hdr.right_idx = right_idx;
            hdr.right_pos = right_pos;
            hdr.cluster_start = start;
            if (i0 < nf[0] && i1 < nf[1] && fr[0][i0] != fr[1][i1]) {
                const uint32_t c0 = "ACGT"[fr[0][i0]], c1 = "ACGT"[fr[1][i1]];
                uint32_t n0 = 0, n1 = 0;
                uint32_t* t0 = p.slot_text + (slot * 2 + 0) * p.cap;
                uint32_t* t1 = p.slot_text + (slot * 2 + 1) * p.cap;
                uint32_t* p0 = p.slot_pos + (slot * 2 + 0) * p.cap;
                uint32_t* p1 = p.slot_pos + (slot * 2 + 1) * p.cap;
#pragma unroll
                for (int it = 0; it < CA_R; ++it) {
                    if (uint32_t(it) * 32u >= len) break;  // (warp-uniform)
                    const uint32_t tx = rtx[it], sf = rsf[it];
                    const uint32_t sample = (tx >= p.nr1_lo) & (p.nr1_big ^ 1u);
                    const bool q0 = rcommon[it] && rch[it] == c0 && sample == 0;  // raw byte compare, ref:clust2snp.cpp:455,466
                    const bool q1 = rcommon[it] && rch[it] == c1 && sample == 1;
                    const uint32_t b0 = __ballot_sync(FULL, q0), b1 = __ballot_sync(FULL, q1);
                    const uint32_t lt = (1u << lane) - 1u;
                    if (q0) {
                        uint32_t r = n0 + __popc(b0 & lt);
                        if (r < p.cap) { t0[r] = tx; p0[r] = sf - p.k_left; }
                    }
                    if (q1) {
                        uint32_t r = n1 + __popc(b1 & lt);
                        if (r < p.cap) { t1[r] = tx; p1[r] = sf - p.k_left; }
                    }
                    n0 += __popc(b0);
                    n1 += __popc(b1);
                }
                hdr.n0 = n0 < p.cap ? n0 : p.cap;
                hdr.n1 = n1 < p.cap ? n1 : p.cap;
                hdr.valid = (n0 > 0 && n1 > 0) ? 1u : 0u;
            }
            if (lane == 0) p.slots[slot] = hdr;
            if (hdr.valid) vmask |= 1u << (i0 * 2 + i1);
        }
    }
    }
    // the valid slots of the block's clusters -> candidate list (at most 4 per flagged cluster: fits)
    if (lane == 0) s_wcnt[warp] = uint32_t(__popc(vmask));
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int q = 0; q < CA_WARPS; ++q) tot += s_wcnt[q];
        s_base = tot ? atomicAdd(&p.dev->n_slots_valid, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (lane == 0 && vmask) {
        unsigned long long at = s_base;
        for (int q = 0; q < warp; ++q) at += s_wcnt[q];
        for (uint32_t b = 0; b < 4; ++b)
            if ((vmask >> b) & 1u) p.cand[at++] = f * 4 + b;
    }
    __syncthreads();  // (s_wcnt / s_base are rewritten by the next round)
    }
}

// ---------------------------------------------------------------------------------------------
// K4: contexts, consensus, support, distance
// ---------------------------------------------------------------------------------------------
struct PackedEventHdr {  // followed by left0[k_left] left1[k_left] right[k_right]
    int32_t D, gap, supp0, supp1, right_len, flags;  // flags bit0 = variant (supp0>0 && supp1>0), bit1 = keep, bits 8-9 = allele pair of the cluster
    uint64_t cluster_start;
};

// The same when the staged reads hold nothing but upper-case ACGT and both context lengths are <= 32: the three strings as two
// bit planes each (bit i of plane p = bit p of the code (char >> 1) & 3 of character i; the host maps the codes back through
// "ACTG").  48 bytes instead of 128 for the default -L 31 -R 30: K4's stores cross PCIe to pinned host memory.
struct CompactEvent {
    uint64_t cluster_start;
    uint8_t D, supp0, supp1, right_len;
    int8_t gap;
    uint8_t flags;  // bit0 = variant, bit1 = keep, bits 2-3 = allele pair of the cluster
    uint16_t pad;
    uint32_t l0[2], l1[2], r[2];
    uint32_t pad2[2];
};
static_assert(sizeof(CompactEvent) == 48, "compact event record");

struct EventParams {
    const CandSlot* slots;
    const uint32_t* slot_text;
    const uint32_t* slot_pos;
    const uint64_t* cand;
    const unsigned long long* n_cand;  // device-resident number of candidates
    uint32_t cap;
    int32_t k_left, k_right, max_gap, max_err, max_snvs;
    const uint8_t* bases;
    const uint64_t* off;
    uint64_t n_reads;
    uint8_t* out;
    uint32_t stride;
    SnpDev* dev;
    const uint32_t* reads_flag;  // device word: bit 0 = the staged reads hold a byte outside ACGTacgt, bit 1 = a lower-case base (null: unknown)
};

constexpr int EV_WARPS = 4;
constexpr int EV_MAXR = 32;  // reads per sample the fast path caches in shared memory
constexpr int EV_B = 8;   // reads gathered per batch (general path)
constexpr int EV_FB = 8;  // ... (fast path)
constexpr int EV_REC_MAX = 32 + 3 * E2S_MAX_K;  // packed record: header + left0 + left1 + right, stride is a multiple of 16

__global__ void __launch_bounds__(EV_WARPS * 32, 12) k_events(EventParams p) {
    __shared__ uint64_t s_base[EV_WARPS][MAX_C_LEN];
    __shared__ char s_cons[EV_WARPS][2][E2S_MAX_K];
    __shared__ __align__(16) uint8_t s_rec[EV_WARPS][EV_REC_MAX];
    __shared__ uint8_t s_ch[EV_WARPS][EV_MAXR][32];  // (fast path) the left contexts of one sample's reads, one byte per lane
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t n_cand = *p.n_cand;
    const uint32_t rflag = p.reads_flag ? p.reads_flag[0] : 1u;
    // every read as long as the first one (k_offsets_check): offsets by arithmetic (UL = 0: look them up)
    const uint64_t UL = (p.reads_flag && p.reads_flag[1] == 0 && p.n_reads) ? p.off[1] - p.off[0] : 0, U0 = UL ? p.off[0] : 0;
    auto off_of = [&](uint64_t r) { return UL ? U0 + r * UL : p.off[r]; };
    unsigned long long nv_local = 0, ne_local = 0;  // (lane 0) variants / kept events of my candidates: one atomic per warp at the end
    const bool fast = p.k_left <= 32 && !(rflag & 1u);
    const bool compact = fast && rflag == 0 && p.k_right <= 32;  // (kernel-uniform: every record of the launch has the same format)
    if (compact && blockIdx.x == 0 && threadIdx.x == 0) p.dev->compact = 1;
    for (uint64_t c = uint64_t(blockIdx.x) * EV_WARPS + w; c < n_cand; c += uint64_t(gridDim.x) * EV_WARPS) {
    __syncwarp();
    const uint64_t slot = p.cand[c];
    const CandSlot hdr = p.slots[slot];
    const int kl = p.k_left;
    bool bad = false;
    uint32_t saw_n = 0;
    int supp[2];

    if (fast && hdr.n0 <= uint32_t(EV_MAXR) && hdr.n1 <= uint32_t(EV_MAXR)) {
        // k_left <= 32 (one context position per lane), reads known to be ACGTacgt only, at most EV_MAXR reads per sample:
        // both samples' read addresses in one go, every read's context gathered ONCE into shared memory (all loads in flight
        // together), consensus and support from there.  The counter slot of a base is (char >> 1) & 3 -- a bijection on ACGT in
        // either case, which is all cons::increment needs (ref:include.hpp:349-358); the stored winner is the raw char.
        {
            const uint32_t n0 = hdr.n0, nt = hdr.n0 + hdr.n1;
            for (uint32_t j = lane; j < nt; j += 32) {
                const uint32_t s = j >= n0, jj = s ? j - n0 : j;
                const uint64_t r = p.slot_text[(slot * 2 + s) * p.cap + jj];
                uint64_t b = 0;
                if (r >= p.n_reads) bad = true;
                else {
                    b = off_of(r) + p.slot_pos[(slot * 2 + s) * p.cap + jj];
                    if (b + uint64_t(kl) > off_of(r + 1)) bad = true;
                }
                s_base[w][s * EV_MAXR + jj] = b;
            }
            bad = __any_sync(FULL, bad);
            __syncwarp();
        }
        if (!bad) {
            const bool act = lane < kl;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const uint32_t nr = s ? hdr.n1 : hdr.n0;
                for (uint32_t j0 = 0; j0 < nr; j0 += EV_FB) {  // EV_FB independent loads in flight, then their stores
                    uint8_t v[EV_FB];
#pragma unroll
                    for (int u = 0; u < EV_FB; ++u) v[u] = (act && j0 + u < nr) ? p.bases[s_base[w][s * EV_MAXR + j0 + u] + lane] : uint8_t(0);
#pragma unroll
                    for (int u = 0; u < EV_FB; ++u)
                        if (j0 + u < nr) s_ch[w][j0 + u][lane] = v[u];
                }
                __syncwarp();
                uint32_t cnt = 0, cur = 'A', cur_b = 0, cur_cnt = 0;  // four 8-bit counters; the winner so far, its slot and its count
                for (uint32_t j = 0; j < nr; ++j) {
                    const uint32_t ch = s_ch[w][j][lane];
                    const uint32_t b8 = (ch << 2) & 24u;  // 8 * ((ch >> 1) & 3)
                    cnt += 1u << b8;
                    const uint32_t cb = (cnt >> b8) & 0xffu;
                    const bool same = b8 == cur_b, lead = cb > cur_cnt;
                    if (!same && lead) {  // strictly more than the current winner: first base to reach the final maximum wins
                        cur = ch;
                        cur_b = b8;
                    }
                    if (same || lead) cur_cnt = cb;  // (the stored char stays the first one that reached the lead)
                }
                if (act) s_cons[w][s][lane] = char(cur);
                // support: reads within max_err mismatches of the consensus, ref:clust2snp.cpp:556-567
                const uint32_t c = act ? cur : 0u;
                int sp = 0;
                for (uint32_t j = 0; j < nr; ++j) sp += int(__popc(__ballot_sync(FULL, uint32_t(s_ch[w][j][lane]) != c))) <= p.max_err;
                supp[s] = sp;
                __syncwarp();
            }
        }
    } else
    for (int s = 0; s < 2; ++s) {
        const uint32_t nr = s ? hdr.n1 : hdr.n0;
        const uint32_t* lt = p.slot_text + (slot * 2 + s) * p.cap;
        const uint32_t* lp = p.slot_pos + (slot * 2 + s) * p.cap;
        for (uint32_t j = lane; j < nr; j += 32) {
            const uint64_t r = lt[j];
            uint64_t b = 0;
            if (r >= p.n_reads) bad = true;
            else {
                b = off_of(r) + lp[j];
                if (b + uint64_t(kl) > off_of(r + 1)) bad = true;
            }
            s_base[w][j] = b;
        }
        bad = __any_sync(FULL, bad);
        __syncwarp();
        if (bad) break;
        // cons::increment, ref:include.hpp:349-358: lane i owns context positions i, i+32, ...
        // (the gathers are issued EV_B reads at a time so that their latencies overlap; the updates stay in list order)
        for (int i = lane; i < kl; i += 32) {
            uint32_t cnt = 0;  // four 8-bit counters (<= 150 reads), one per base code
            uint32_t cur = 'A', cur_b = 0, cur_cnt = 0;  // the winner so far, its code and its count
            for (uint32_t j0 = 0; j0 < nr; j0 += EV_B) {
                uint32_t ch[EV_B];
#pragma unroll
                for (int u = 0; u < EV_B; ++u) ch[u] = j0 + u < nr ? p.bases[s_base[w][j0 + u] + i] : 0u;
#pragma unroll
                for (int u = 0; u < EV_B; ++u) {
                    if (j0 + u < nr) {
                        saw_n |= is_n(ch[u]);
                        const uint32_t b = base_code(ch[u]);
                        cnt += 1u << (8 * b);
                        const uint32_t cb = (cnt >> (8 * b)) & 0xffu;
                        if (b == cur_b) {
                            cur_cnt = cb;  // (the stored char stays the first one that reached the lead)
                        } else if (cb > cur_cnt) {  // strictly more than the current winner: first base to reach the final maximum wins
                            cur = ch[u];
                            cur_b = b;
                            cur_cnt = cb;
                        }
                    }
                }
            }
            s_cons[w][s][i] = char(cur);
        }
        __syncwarp();
        // support: reads within max_err mismatches of the consensus, ref:clust2snp.cpp:556-567
        int sp = 0;
        if (kl <= 32) {  // one context position per lane: a mismatch count is the popcount of a ballot
            const bool act = lane < kl;
            const uint32_t c = act ? uint32_t(uint8_t(s_cons[w][s][lane])) : 0u;
            for (uint32_t j0 = 0; j0 < nr; j0 += EV_B) {
                uint32_t ch[EV_B];
#pragma unroll
                for (int u = 0; u < EV_B; ++u) ch[u] = (act && j0 + u < nr) ? p.bases[s_base[w][j0 + u] + lane] : c;
#pragma unroll
                for (int u = 0; u < EV_B; ++u)
                    if (j0 + u < nr) sp += int(__popc(__ballot_sync(FULL, ch[u] != c))) <= p.max_err;  // (warp-uniform condition)
            }
        } else
        for (uint32_t j0 = 0; j0 < nr; j0 += EV_B) {
            int d[EV_B];
#pragma unroll
            for (int u = 0; u < EV_B; ++u) d[u] = 0;
            for (int i = lane; i < kl; i += 32) {
                const uint8_t c = uint8_t(s_cons[w][s][i]);
#pragma unroll
                for (int u = 0; u < EV_B; ++u)
                    if (j0 + u < nr) d[u] += (c != p.bases[s_base[w][j0 + u] + i]);
            }
#pragma unroll
            for (int u = 0; u < EV_B; ++u) {
                sp += (j0 + u < nr && int(__reduce_add_sync(FULL, uint32_t(d[u]))) <= p.max_err);
            }
        }
        supp[s] = sp;
        __syncwarp();
    }
    if (bad) {
        if (lane == 0) p.dev->bad_ref = 1;
        return;
    }
    if (__any_sync(FULL, saw_n) && lane == 0) atomicOr(&p.dev->saw_n, 1ull);

    // the record is assembled in shared memory and leaves the SM as whole 16-byte vectors: the destination is pinned
    // HOST memory, where byte-sized stores would each become a PCIe transaction
    uint8_t* o = s_rec[w];
    PackedEventHdr* oh = reinterpret_cast<PackedEventHdr*>(o);
    char* o_l0 = reinterpret_cast<char*>(o + sizeof(PackedEventHdr));
    char* o_l1 = o_l0 + kl;
    char* o_r = o_l1 + kl;
    const bool variant = supp[0] > 0 && supp[1] > 0;

    // right context = reads[idx].substr(pos, k_right), ref:clust2snp.cpp:604
    int rl = 0;
    uint32_t r_ch = 0;  // (compact) my character of the right context
    if (variant) {
        if (hdr.right_idx >= p.n_reads) bad = true;
        else {
            const uint64_t rb = off_of(hdr.right_idx), re = off_of(hdr.right_idx + 1);
            if (hdr.right_pos > re - rb) bad = true;
            else {
                uint64_t avail = re - rb - hdr.right_pos;
                rl = int(avail < uint64_t(p.k_right) ? avail : uint64_t(p.k_right));
                if (compact) {
                    if (lane < rl) r_ch = p.bases[rb + hdr.right_pos + lane];
                } else {
                    for (int i = lane; i < rl; i += 32) o_r[i] = char(p.bases[rb + hdr.right_pos + i]);
                }
            }
        }
        if (bad) {
            if (lane == 0) p.dev->bad_ref = 1;
            return;
        }
    }
    uint32_t cm[6] = {0, 0, 0, 0, 0, 0};  // (compact) bit planes of left0, left1, right
    if (compact) {
        const uint32_t a_ch = lane < kl ? uint32_t(uint8_t(s_cons[w][0][lane])) : 0u, b_ch = lane < kl ? uint32_t(uint8_t(s_cons[w][1][lane])) : 0u;
        cm[0] = __ballot_sync(FULL, lane < kl && (a_ch & 2u));
        cm[1] = __ballot_sync(FULL, lane < kl && (a_ch & 4u));
        cm[2] = __ballot_sync(FULL, lane < kl && (b_ch & 2u));
        cm[3] = __ballot_sync(FULL, lane < kl && (b_ch & 4u));
        cm[4] = __ballot_sync(FULL, lane < rl && (r_ch & 2u));
        cm[5] = __ballot_sync(FULL, lane < rl && (r_ch & 4u));
    } else {
        for (int i = lane; i < kl; i += 32) {
            o_l0[i] = s_cons[w][0][i];
            o_l1[i] = s_cons[w][1][i];
        }
    }

    // distance(), ref:clust2snp.cpp:254-302.  g = 0: plain Hamming; g >= 1: drop g chars on the right of a / of b
    const char* a = s_cons[w][0];
    const char* b = s_cons[w][1];
    uint32_t best_ab = 0xffffffffu, best_ba = 0xffffffffu;  // (dist << 8) | g : min => smallest dist, then smallest g
    int d0 = 0;
    if (kl <= 32) {
        // lane t holds a[t] and b[t]: dH of the right-aligned strings with g characters dropped on the right of a
        // (resp. b) = #{t >= g : a[t-g] != b[t]} (resp. a[t] != b[t-g]) = popcount of a ballot; every lane ends with the result
        const uint32_t av = lane < kl ? uint32_t(uint8_t(a[lane])) : 0x100u, bv = lane < kl ? uint32_t(uint8_t(b[lane])) : 0x100u;
        for (int g = 0; g <= p.max_gap; ++g) {
            const uint32_t as = __shfl_up_sync(FULL, av, g), bs = __shfl_up_sync(FULL, bv, g);
            const bool in = lane >= g && lane < kl;
            const int dab = __popc(__ballot_sync(FULL, in && as != bv));
            const int dba = __popc(__ballot_sync(FULL, in && av != bs));
            if (g == 0) d0 = dab;
            else {
                const uint32_t kab = (uint32_t(dab + g) << 8) | uint32_t(g), kba = (uint32_t(dba + g) << 8) | uint32_t(g);
                best_ab = kab < best_ab ? kab : best_ab;
                best_ba = kba < best_ba ? kba : best_ba;
            }
        }
    } else {
    for (int g = lane; g <= p.max_gap; g += 32) {
        int dab = 0, dba = 0;
        const int n = kl - g;  // compared length, right aligned
        for (int i = 0; i < n; ++i) {
            dab += a[kl - g - 1 - i] != b[kl - 1 - i];
            dba += a[kl - 1 - i] != b[kl - g - 1 - i];
        }
        if (g == 0) d0 = dab;
        else {
            uint32_t kab = (uint32_t(dab + g) << 8) | uint32_t(g), kba = (uint32_t(dba + g) << 8) | uint32_t(g);
            best_ab = kab < best_ab ? kab : best_ab;
            best_ba = kba < best_ba ? kba : best_ba;
        }
    }
    d0 = __shfl_sync(FULL, d0, 0);
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
        uint32_t x = __shfl_xor_sync(FULL, best_ab, o2), y = __shfl_xor_sync(FULL, best_ba, o2);
        best_ab = x < best_ab ? x : best_ab;
        best_ba = y < best_ba ? y : best_ba;
    }
    }
    int D, gap;
    const int min_ab = int(best_ab >> 8), g_ab = int(best_ab & 0xff), min_ba = int(best_ba >> 8), g_ba = int(best_ba & 0xff);
    if (d0 < min_ab && d0 < min_ba) { D = d0; gap = 0; }
    else if (min_ab < min_ba) { D = min_ab - g_ab; gap = g_ab; }
    else { D = min_ba - g_ba; gap = -g_ba; }
    const bool keep = variant && D <= p.max_snvs;
    if (lane == 0) {
        if (compact) {
            CompactEvent* ce = reinterpret_cast<CompactEvent*>(o);
            ce->cluster_start = hdr.cluster_start;
            ce->D = uint8_t(D);
            ce->supp0 = uint8_t(supp[0]);
            ce->supp1 = uint8_t(supp[1]);
            ce->right_len = uint8_t(rl);
            ce->gap = int8_t(gap);
            ce->flags = uint8_t((variant ? 1 : 0) | (keep ? 2 : 0) | int(slot & 3) << 2);
            ce->pad = 0;
            ce->l0[0] = cm[0]; ce->l0[1] = cm[1];
            ce->l1[0] = cm[2]; ce->l1[1] = cm[3];
            ce->r[0] = cm[4]; ce->r[1] = cm[5];
            ce->pad2[0] = ce->pad2[1] = 0;
        } else {
            oh->D = D;
            oh->gap = gap;
            oh->supp0 = supp[0];
            oh->supp1 = supp[1];
            oh->right_len = rl;
            oh->flags = (variant ? 1 : 0) | (keep ? 2 : 0) | int32_t((slot & 3) << 8);  // bits 8-9: allele pair, for the output order
            oh->cluster_start = hdr.cluster_start;
        }
        nv_local += variant ? 1u : 0u;
        ne_local += keep ? 1u : 0u;
    }
    __syncwarp();
    {
        const uint32_t stride = compact ? uint32_t(sizeof(CompactEvent)) : p.stride;
        const uint4* src = reinterpret_cast<const uint4*>(s_rec[w]);
        uint4* dst = reinterpret_cast<uint4*>(p.out + c * stride);
        for (uint32_t i = lane; i < stride / 16; i += 32) dst[i] = src[i];
    }
    }
    if (lane == 0) {
        if (nv_local) atomicAdd(&p.dev->n_variants, nv_local);
        if (ne_local) atomicAdd(&p.dev->n_events, ne_local);
    }
}

// ---------------------------------------------------------------------------------------------
// staged reads: any byte outside ACGTacgt (bit 0), any lower-case base (bit 1)?  One pass when the reads are staged; K4's fast
// path and its compact event records depend on the answer
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reads_check(const uint8_t* __restrict__ bases, uint64_t n, uint32_t* flag) {
    uint32_t bad = 0, lower = 0;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t head = (16 - (reinterpret_cast<uintptr_t>(bases) & 15)) & 15;  // bytes before the first 16-byte boundary
    const uint64_t nv = n > head ? (n - head) / 16 : 0;
    const uint4* v = reinterpret_cast<const uint4*>(bases + head);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 q = v[i];
        const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t u = ws[k] & 0xDFDFDFDFu;
            const uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
            bad |= ~ok ? 1u : 0u;
            lower |= ws[k] & 0x20202020u;
        }
    }
    if (blockIdx.x == 0) {  // the unaligned head and the tail, byte by byte
        auto one = [&](uint64_t x) {
            const uint32_t u = bases[x] & 0xDFu;
            if (!(u == 'A' || u == 'C' || u == 'G' || u == 'T')) bad = 1;
            lower |= bases[x] & 0x20u;
        };
        for (uint64_t x = threadIdx.x; x < (head < n ? head : n); x += blockDim.x) one(x);
        for (uint64_t x = head + nv * 16 + threadIdx.x; x < n; x += blockDim.x) one(x);
    }
    const int any_bad = __syncthreads_or(int(bad != 0)), any_lower = __syncthreads_or(int(lower != 0));
    if (threadIdx.x == 0 && (any_bad || any_lower)) atomicOr(flag, (any_bad ? 1u : 0u) | (any_lower ? 2u : 0u));
}

// flag[1] |= 1 unless every read has the length of the first one (K4 then computes read offsets instead of looking them up: one
// dependent DRAM access less per gathered context)
__global__ void __launch_bounds__(256) k_offsets_check(const uint64_t* __restrict__ off, uint64_t n_reads, uint32_t* flag) {
    const uint64_t L = off[1] - off[0];
    uint32_t bad = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_reads; i += uint64_t(gridDim.x) * blockDim.x)
        bad |= uint32_t(off[i + 1] - off[i] != L);
    if (__syncthreads_or(int(bad)) && threadIdx.x == 0) atomicOr(flag + 1, 1u);
}

cudaError_t launch_reads_check(const uint8_t* d_bases, uint64_t n, const uint64_t* d_off, uint64_t n_reads, uint32_t* flag, cudaStream_t stream,
                               int sm_count) {
    if (!n || !n_reads) return cudaSuccess;
    uint64_t blocks = (n / 16 + 255) / 256 + 1;
    if (blocks > uint64_t(sm_count) * 8) blocks = uint64_t(sm_count) * 8;
    k_reads_check<<<unsigned(blocks), 256, 0, stream>>>(d_bases, n, flag);
    blocks = (n_reads + 255) / 256;
    if (blocks > uint64_t(sm_count) * 8) blocks = uint64_t(sm_count) * 8;
    k_offsets_check<<<unsigned(blocks), 256, 0, stream>>>(d_off, n_reads, flag);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------
struct SnpWork {
    uint8_t* zero_blk = nullptr; size_t zero_cap = 0;  // everything a pass needs zeroed, in one block: one memset
    SurvEntry* survivors = nullptr; size_t survivors_cap = 0;
    SurvEntry* flagged = nullptr; size_t flagged_cap = 0;
    uint64_t cap_surv_used = 0, cap_flag_used = 0;          // capacities of the pass in flight (snp_collect compares the counters with them)
    CandSlot* slots = nullptr; size_t slots_cap = 0;
    uint32_t* slot_text = nullptr; uint32_t* slot_pos = nullptr; size_t slot_list_cap = 0;
    uint64_t* cand = nullptr; size_t cand_cap = 0;
    SnpDev* dev = nullptr;
    SnpDev* h_dev = nullptr;                                // pinned staging of the counters
    uint8_t* h_events = nullptr; size_t h_events_cap = 0;  // pinned + mapped: K4 writes the packed candidates here
    uint8_t* d_events = nullptr;                            // device view of h_events
    uint64_t want_survivors = 0, want_flagged = 0;          // capacity guesses, grown when a pass overflows
    uint64_t n_cand = 0;
    uint32_t stride = 0;
    int k_left = 0, k_right = 0;
};

SnpWork* snp_work_create() { return new SnpWork(); }

void snp_work_destroy(SnpWork* w) {
    if (!w) return;
    cudaFree(w->zero_blk); cudaFree(w->survivors);
    cudaFree(w->flagged);
    cudaFree(w->slots); cudaFree(w->slot_text); cudaFree(w->slot_pos); cudaFree(w->cand);
    cudaFreeHost(w->h_dev);
    cudaFreeHost(w->h_events);
    delete w;
}

template <typename T>
static cudaError_t ensure(T*& ptr, size_t& cap, size_t need) {
    if (need <= cap && ptr) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    size_t n = need + need / 4 + 64;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T));
    cap = e == cudaSuccess ? n : 0;
    return e;
}

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { *err = #x; return _e; } } while (0)

// One pass of K3a / K3x / K3b / K4 with the current capacity guesses, enqueued on the stream together with the copy of the
// device counters to pinned host memory.  No synchronisation.
static cudaError_t snp_enqueue(SnpWork* w, const SnpArrays& a, const e2s_snp_params& p, int max_clust_length,
                               const uint8_t* d_read_bases, const uint64_t* d_read_off, uint64_t n_reads, int sm_count,
                               cudaStream_t stream, uint64_t* launches, const char** err, KernelTimer* timer,
                               const SurvEntry* pre_list, uint64_t pre_count, const unsigned long long* d_pre_count,
                               const int32_t* d_max_len) {
    const uint32_t nr1_big = p.nr_reads1 > 0xffffffffull ? 1u : 0u;
    const uint32_t nr1_lo = nr1_big ? 0xffffffffu : uint32_t(p.nr_reads1);
    const uint32_t num_tiles = uint32_t((a.n_local + PS_T - 1) / PS_T);
    const uint32_t cap = uint32_t(p.consensus_reads < MAX_C_LEN ? p.consensus_reads : MAX_C_LEN);
    w->stride = uint32_t((sizeof(PackedEventHdr) + 2 * size_t(p.k_left) + size_t(p.k_right) + 15) & ~size_t(15));
    // on the fused path (pre_list) the survivor list is the caller's: the flagged list is bounded by ITS capacity, not by the
    // capacity guess of a K3a pass that does not run (the retry could otherwise never grow past that guess)
    const uint64_t cap_surv = pre_list ? (pre_count ? pre_count : 1) : (w->want_survivors < a.m ? w->want_survivors : a.m);
    const uint64_t cap_flag = w->want_flagged < cap_surv ? w->want_flagged : cap_surv;
    w->cap_surv_used = cap_surv;
    w->cap_flag_used = cap_flag;
    const uint64_t n_slots = cap_flag * 4;
    if (!pre_list) CK(ensure(w->survivors, w->survivors_cap, size_t(cap_surv)));
    CK(ensure(w->flagged, w->flagged_cap, size_t(cap_flag)));
    CK(ensure(w->slots, w->slots_cap, size_t(n_slots)));
    if (size_t(n_slots) * 2 * cap > w->slot_list_cap || !w->slot_text) {
        cudaFree(w->slot_text); cudaFree(w->slot_pos);
        w->slot_text = w->slot_pos = nullptr;
        w->slot_list_cap = 0;
        const size_t n = size_t(n_slots) * 2 * cap;
        CK(cudaMalloc(reinterpret_cast<void**>(&w->slot_text), n * 4));
        CK(cudaMalloc(reinterpret_cast<void**>(&w->slot_pos), n * 4));
        w->slot_list_cap = n;
    }
    CK(ensure(w->zero_blk, w->zero_cap, sizeof(SnpDev)));  // the device counters: the one thing a pass needs zeroed
    w->dev = reinterpret_cast<SnpDev*>(w->zero_blk);
    CK(cudaMemsetAsync(w->zero_blk, 0, sizeof(SnpDev), stream));
    CK(ensure(w->cand, w->cand_cap, size_t(n_slots)));
    if (size_t(n_slots) * w->stride > w->h_events_cap) {
        cudaFreeHost(w->h_events);
        w->h_events = nullptr;
        w->h_events_cap = 0;
        const size_t bytes = size_t(n_slots) * w->stride;
        CK(cudaHostAlloc(reinterpret_cast<void**>(&w->h_events), bytes, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&w->d_events), w->h_events, 0));
        w->h_events_cap = bytes;
    }

    if (!pre_list) {   // K3a: base-code prefilter on the resident bit planes
        ScanParams sp;
        sp.a = a;
        sp.limit = uint64_t(num_tiles) * PS_T;
        sp.min_len = uint32_t(2 * p.mcov_out);
        sp.max_len = uint32_t(max_clust_length);
        sp.mcov = uint32_t(p.mcov_out);
        sp.survivors = w->survivors;
        sp.cap_surv = cap_surv;
        sp.dev = w->dev;
        uint64_t grid = (a.m + PS_THREADS * PS_U - 1) / (PS_THREADS * PS_U);
        if (grid > uint64_t(sm_count) * 16) grid = uint64_t(sm_count) * 16;
        if (timer) timer->begin(E2S_KERNEL_SCAN, stream);
        k_code_scan<<<unsigned(grid), PS_THREADS, 0, stream>>>(sp);
        if (timer) timer->end(stream);
        CK(cudaGetLastError());
        ++*launches;
    }
    {   // K3x: exact filters on the survivors
        ExactParams ep;
        ep.a = a;
        ep.list = pre_list ? pre_list : w->survivors;
        ep.n_list = pre_list ? d_pre_count : &w->dev->n_survivors;
        ep.n_list_host = pre_count;
        ep.cap_list = pre_list ? pre_count : cap_surv;
        ep.min_len = uint32_t(2 * p.mcov_out);
        ep.max_len = uint32_t(max_clust_length);
        ep.d_max_len = d_max_len;
        ep.mcov = uint32_t(p.mcov_out);
        ep.k_right = uint32_t(p.k_right);
        ep.nr1_lo = nr1_lo;
        ep.nr1_big = nr1_big;
        ep.flagged = w->flagged;
        ep.cap_flagged = cap_flag;
        ep.dev = w->dev;
        if (timer) timer->begin(E2S_KERNEL_EXACT, stream);
        k_cluster_exact<<<unsigned(sm_count) * 16, EX_THREADS, 0, stream>>>(ep);
        if (timer) timer->end(stream);
        CK(cudaGetLastError());
        ++*launches;
    }
    // Flagged clusters, their candidate slots and the events come out in whatever order the atomics give: the
    // reference's order (eBWT position, then allele pair) is restored from (cluster_start, pair) when the events are
    // fetched (snp_fetch_events), so the step needs no ordered compaction.
    {
        CandParams cp;
        cp.a = a;
        cp.flagged = w->flagged;
        cp.n_flagged = &w->dev->n_flagged;
        cp.cap_flagged = cap_flag;
        cp.mcov = uint32_t(p.mcov_out);
        cp.k_left = uint32_t(p.k_left);
        cp.k_right = uint32_t(p.k_right);
        cp.cap = cap;
        cp.nr1_lo = nr1_lo;
        cp.nr1_big = nr1_big;
        cp.slots = w->slots;
        cp.slot_text = w->slot_text;
        cp.slot_pos = w->slot_pos;
        cp.cand = w->cand;
        cp.dev = w->dev;
        uint64_t blocks = (cap_flag + CA_WARPS - 1) / CA_WARPS;
        if (blocks > uint64_t(sm_count) * 8) blocks = uint64_t(sm_count) * 8;
        if (timer) timer->begin(E2S_KERNEL_CAND, stream);
        k_candidates<<<unsigned(blocks), CA_WARPS * 32, 0, stream>>>(cp);
        if (timer) timer->end(stream);
        CK(cudaGetLastError());
        EventParams ep;
        ep.slots = w->slots;
        ep.slot_text = w->slot_text;
        ep.slot_pos = w->slot_pos;
        ep.cand = w->cand;
        ep.n_cand = &w->dev->n_slots_valid;
        ep.cap = cap;
        ep.k_left = p.k_left;
        ep.k_right = p.k_right;
        ep.max_gap = p.max_gap;
        ep.max_err = p.max_err;
        ep.max_snvs = p.max_snvs;
        ep.bases = d_read_bases;
        ep.off = d_read_off;
        ep.n_reads = n_reads;
        ep.out = w->d_events;
        ep.stride = w->stride;
        ep.dev = w->dev;
        ep.reads_flag = getenv("E2S_K4_GENERIC") ? nullptr : a.reads_flag;  // (test hook: keep K4 on the general path)
        uint64_t eblocks = (n_slots + EV_WARPS - 1) / EV_WARPS;
        if (eblocks > uint64_t(sm_count) * 16) eblocks = uint64_t(sm_count) * 16;
        if (timer) timer->begin(E2S_KERNEL_EVENTS, stream);
        k_events<<<unsigned(eblocks), EV_WARPS * 32, 0, stream>>>(ep);
        if (timer) timer->end(stream);
        CK(cudaGetLastError());
        *launches += 2;
    }
    CK(cudaMemcpyAsync(w->h_dev, w->dev, sizeof(SnpDev), cudaMemcpyDeviceToHost, stream));
    return cudaSuccess;
}

// After the stream has been synchronised: the counters of the pass.  false + *rc == cudaSuccess: a capacity guess was too
// small (the guesses have been raised: run the pass again); false + *rc != cudaSuccess: the pass failed.
bool snp_collect(SnpWork* w, e2s_snp_counts* counts, const char** err, cudaError_t* rc) {
    *rc = cudaSuccess;
    const SnpDev& hd = *w->h_dev;
    const uint64_t cap_surv = w->cap_surv_used, cap_flag = w->cap_flag_used;
    if (hd.n_survivors > cap_surv || hd.n_flagged > cap_flag) {
        if (hd.n_survivors > cap_surv) w->want_survivors = hd.n_survivors + hd.n_survivors / 8 + 1024;
        // flagged clusters are a subset of the survivors; when the survivor list overflowed the flagged count is a lower bound
        uint64_t guess = hd.n_flagged + hd.n_flagged / 8 + 1024;
        if (hd.n_survivors > cap_surv) guess = guess * (hd.n_survivors / cap_surv + 1);
        if (guess > w->want_flagged) w->want_flagged = guess;
        return false;
    }
    counts->n_analysed = hd.n_analysed;
    counts->n_flagged = hd.n_flagged;
    counts->n_candidates = hd.n_slots_valid;
    counts->n_variants = hd.n_variants;
    counts->n_events = hd.n_events;
    counts->saw_n = hd.saw_n;
    if (hd.bad_ref) {
        *err = "a candidate references a read/offset outside the staged reads";
        *rc = cudaErrorInvalidValue;
        return false;
    }
    w->n_cand = hd.n_slots_valid;
    return true;
}

cudaError_t snp_run(SnpWork* w, const SnpArrays& a, const e2s_snp_params& p, int max_clust_length,
                    const uint8_t* d_read_bases, const uint64_t* d_read_off, uint64_t n_reads, int sm_count,
                    cudaStream_t stream, e2s_snp_counts* counts, uint64_t* launches, const char** err,
                    KernelTimer* timer, const SurvEntry* pre_list, uint64_t pre_count, const unsigned long long* d_pre_count,
                    const int32_t* d_max_len, bool sync) {
    memset(counts, 0, sizeof *counts);
    w->n_cand = 0;
    w->k_left = p.k_left;
    w->k_right = p.k_right;
    if (!w->h_dev) CK(cudaHostAlloc(reinterpret_cast<void**>(&w->h_dev), sizeof(SnpDev), cudaHostAllocDefault));
    memset(w->h_dev, 0, sizeof(SnpDev));
    w->cap_surv_used = w->cap_flag_used = 0;
    if (!pre_list && (a.m == 0 || a.n_local == 0)) return cudaSuccess;

    // first guesses: clusters with two frequent base codes are variants and repeats, a small fraction of all
    const uint64_t m_guess = a.m ? a.m : a.n_local / 32 + 1;
    if (!w->want_survivors) w->want_survivors = m_guess / 64 + 4096;
    if (!w->want_flagged) w->want_flagged = m_guess / 256 + 2048;
    if (const char* dbg = getenv("E2S_SNP_FIRST_CAPACITY")) {  // test hook: start from a tiny guess to exercise the retry path
        const uint64_t v = strtoull(dbg, nullptr, 10);
        if (v) w->want_survivors = w->want_flagged = v;
    }
    for (int attempt = 0;; ++attempt) {
        cudaError_t e = snp_enqueue(w, a, p, max_clust_length, d_read_bases, d_read_off, n_reads, sm_count, stream, launches, err, timer,
                                    pre_list, pre_count, d_pre_count, d_max_len);
        if (e != cudaSuccess) return e;
        if (!sync) return cudaSuccess;  // the caller synchronises and calls snp_collect (and snp_run again if a guess was too small)
        CK(cudaStreamSynchronize(stream));  // the only synchronisation of the pass
        cudaError_t rc;
        if (snp_collect(w, counts, err, &rc)) return cudaSuccess;
        if (rc != cudaSuccess) return rc;
        if (attempt >= 3) { *err = "phase 2 capacity did not converge"; return cudaErrorUnknown; }
    }
}

// expands the host copy of the packed candidates, keeping the variants (supp0>0 && supp1>0), in the reference's order:
// clusters by eBWT position, the allele pairs of a cluster in the nested order of ref:clust2snp.cpp:431-435
uint32_t snp_event_stride(const SnpWork* w) { return (w->h_dev && w->h_dev->compact) ? uint32_t(sizeof(CompactEvent)) : w->stride; }

cudaError_t snp_fetch_events(SnpWork* w, e2s_event* host, uint64_t cap, uint64_t* n, cudaStream_t) {
    *n = 0;
    if (w->n_cand == 0) return cudaSuccess;
    const uint8_t* tmp = w->h_events;
    const bool compact = w->h_dev->compact != 0;
    const size_t stride = compact ? sizeof(CompactEvent) : size_t(w->stride);
    struct Key { uint64_t start; uint32_t pair; uint64_t c; };
    std::vector<Key> order;
    order.reserve(size_t(w->n_cand));
    for (uint64_t c = 0; c < w->n_cand; ++c) {
        if (compact) {
            const CompactEvent* h = reinterpret_cast<const CompactEvent*>(tmp + c * stride);
            if (h->flags & 1) order.push_back(Key{h->cluster_start, uint32_t(h->flags >> 2) & 3u, c});
        } else {
            const PackedEventHdr* h = reinterpret_cast<const PackedEventHdr*>(tmp + c * stride);
            if (h->flags & 1) order.push_back(Key{h->cluster_start, uint32_t(h->flags >> 8) & 3u, c});
        }
    }
    std::sort(order.begin(), order.end(), [](const Key& a, const Key& b) { return a.start != b.start ? a.start < b.start : a.pair < b.pair; });
    uint64_t k = 0;
    for (const Key& key : order) {
        const uint8_t* o = tmp + key.c * stride;
        if (host && k < cap) {
            e2s_event* ev = &host[k];
            memset(ev, 0, sizeof *ev);
            if (compact) {
                const CompactEvent* h = reinterpret_cast<const CompactEvent*>(o);
                ev->D = h->D;
                ev->gap = h->gap;
                ev->supp0 = h->supp0;
                ev->supp1 = h->supp1;
                ev->right_len = h->right_len;
                ev->keep = (h->flags & 2) ? 1 : 0;
                ev->cluster_start = h->cluster_start;
                auto expand = [](char* dst, const uint32_t* m, int len) {  // code (char >> 1) & 3: A 0, C 1, T 2, G 3
                    for (int i = 0; i < len; ++i) dst[i] = "ACTG"[((m[0] >> i) & 1u) | (((m[1] >> i) & 1u) << 1)];
                };
                expand(ev->left0, h->l0, w->k_left);
                expand(ev->left1, h->l1, w->k_left);
                expand(ev->right, h->r, h->right_len);
            } else {
                const PackedEventHdr* h = reinterpret_cast<const PackedEventHdr*>(o);
                ev->D = h->D;
                ev->gap = h->gap;
                ev->supp0 = h->supp0;
                ev->supp1 = h->supp1;
                ev->right_len = h->right_len;
                ev->keep = (h->flags & 2) ? 1 : 0;
                ev->cluster_start = h->cluster_start;
                const char* s = reinterpret_cast<const char*>(o + sizeof(PackedEventHdr));
                memcpy(ev->left0, s, size_t(w->k_left));
                memcpy(ev->left1, s + w->k_left, size_t(w->k_left));
                memcpy(ev->right, s + 2 * w->k_left, size_t(h->right_len));
            }
        }
        ++k;
    }
    *n = k;
    return cudaSuccess;
}

}  // namespace e2s
