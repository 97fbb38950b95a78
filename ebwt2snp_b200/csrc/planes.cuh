// Resident bit planes of the 2-bit base code (base_to_int, ref:include.hpp:265-279: A=0 C=1 G=2 T=3, every other
// byte 0) and the BWT-only prefilter of find_variants built on them.  The planes are part of the shard's resident
// layout: e2s_shard_seal derives them once from the BWT bytes (k_bwt_planes), independently of any option of the two
// tools; K3a (k_code_scan) and K2's fused mode (k_cluster_emit) answer every cluster with a few range popcounts on
// them instead of streaming the bytes (0.25 B/position instead of 1 B/position).
#pragma once

#include <stdint.h>

#include "common.cuh"

namespace e2s {

// Layout: one uint4 per 64 positions = { plane0 bits 0..31, plane0 bits 32..63, plane1 bits 0..31, plane1 bits 32..63 }
// (plane0 = bit 0 of the code: C, T; plane1 = bit 1: G, T), so one 16-byte load brings both planes of 64 positions.
// Quad q covers local positions [64 q - PL_PAD, 64 q - PL_PAD + 64).
constexpr int PL_PAD = 192;  // positions before local position 0 (>= PAD_L, multiple of 64)

__host__ __device__ inline uint64_t plane_quads(uint64_t alloc_r) { return (uint64_t(PL_PAD) + alloc_r + 63) / 64; }

// The prefilter: how many base codes (callers only ask: two or more?) occur >= mcov times among the `len` positions from position `lo` on, where
// `planes` holds the quads of the positions from -PL_PAD on (the shard's array: lo = local position; a staged window:
// lo relative to the window's position 0).  GLOBAL: read through the read-only path, else plain (shared memory) loads.
// A cluster with at most one such code cannot pass find_variants (ref:clust2snp.cpp:402-429): counts[s][c] <= total[c].
// first: the quad holding position lo if the caller has already fetched it (it overlaps that load with others), else null.
template <bool GLOBAL>
__device__ __forceinline__ uint32_t frequent_codes(const uint4* __restrict__ planes, int64_t lo, uint32_t len, uint32_t mcov,
                                                   const uint4* first = nullptr) {
    const uint64_t b_lo = uint64_t(lo + PL_PAD), b_last = b_lo + len - 1;
    const uint64_t q_lo = b_lo >> 6, q_last = b_last >> 6;
    const unsigned long long m_first = ~0ull << (b_lo & 63), m_last = ~0ull >> (63 - (b_last & 63));
    // Cheap bound first (one popcount per 64 positions instead of three): if fewer than mcov positions carry a code other
    // than the FIRST position's, no other code can be frequent -- true for all but the clusters at variant sites / repeats
    // and those that happen to begin with a minority symbol, which take the exact count below.
    {
        unsigned long long f0 = 0, f1 = 0;  // the first position's code bits, spread over a whole word
        uint32_t others = 0;
        for (uint64_t q = q_lo; q <= q_last; ++q) {
            const uint4 v = (first && q == q_lo) ? *first : (GLOBAL ? __ldg(planes + q) : planes[q]);
            const unsigned long long x0 = (uint64_t(v.y) << 32) | v.x, x1 = (uint64_t(v.w) << 32) | v.z;
            unsigned long long mask = q == q_lo ? m_first : ~0ull;
            if (q == q_last) mask &= m_last;
            if (q == q_lo) {
                f0 = 0ull - ((x0 >> (b_lo & 63)) & 1ull);
                f1 = 0ull - ((x1 >> (b_lo & 63)) & 1ull);
            }
            others += __popcll(((x0 ^ f0) | (x1 ^ f1)) & mask);
        }
        if (others < mcov) return 1;  // at most one frequent code (the exact number does not matter to the callers)
    }
    uint32_t nC = 0, nG = 0, nT = 0;
    for (uint64_t q = q_lo; q <= q_last; ++q) {
        unsigned long long mask = q == q_lo ? m_first : ~0ull;
        if (q == q_last) mask &= m_last;
        const uint4 v = GLOBAL ? __ldg(planes + q) : planes[q];
        const unsigned long long x0 = ((uint64_t(v.y) << 32) | v.x) & mask, x1 = ((uint64_t(v.w) << 32) | v.z) & mask;
        nT += __popcll(x0 & x1);
        nC += __popcll(x0 & ~x1);
        nG += __popcll(x1 & ~x0);
    }
    const uint32_t nA = len - nC - nG - nT;
    return uint32_t(nA >= mcov) + uint32_t(nC >= mcov) + uint32_t(nG >= mcov) + uint32_t(nT >= mcov);
}

// The one-popcount bound alone, on the shard's plane array in global memory: true = the cluster MAY have two frequent base
// codes (it goes to the exact test), false = fewer than mcov of its positions differ from the first one's code.
__device__ __forceinline__ bool frequent_bound(const uint4* __restrict__ planes, int64_t lo, uint32_t len, uint32_t mcov) {
    const uint64_t b_lo = uint64_t(lo + PL_PAD), b_last = b_lo + len - 1;
    unsigned long long f0 = 0, f1 = 0;
    uint32_t others = 0;
    for (uint64_t q = b_lo >> 6; q <= (b_last >> 6); ++q) {
        const uint4 v = __ldg(planes + q);
        const unsigned long long x0 = (uint64_t(v.y) << 32) | v.x, x1 = (uint64_t(v.w) << 32) | v.z;
        unsigned long long mask = ~0ull;
        if (q == (b_lo >> 6)) {
            mask = ~0ull << (b_lo & 63);
            f0 = 0ull - ((x0 >> (b_lo & 63)) & 1ull);
            f1 = 0ull - ((x1 >> (b_lo & 63)) & 1ull);
        }
        if (q == (b_last >> 6)) mask &= ~0ull >> (63 - (b_last & 63));
        others += __popcll(((x0 ^ f0) | (x1 ^ f1)) & mask);
    }
    return others >= mcov;
}

}  // namespace e2s
