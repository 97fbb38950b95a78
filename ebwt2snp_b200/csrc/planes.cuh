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

// The prefilter: how many base codes occur >= mcov times among the `len` positions from local position `lo` on.
// A cluster with at most one such code cannot pass find_variants (ref:clust2snp.cpp:402-429): counts[s][c] <= total[c].
__device__ __forceinline__ uint32_t frequent_codes(const uint4* __restrict__ planes, int64_t lo, uint32_t len, uint32_t mcov) {
    const uint64_t b_lo = uint64_t(lo + PL_PAD), b_last = b_lo + len - 1;
    uint64_t q = b_lo >> 6;
    const uint64_t q_last = b_last >> 6;
    unsigned long long mask = ~0ull << (b_lo & 63);
    uint32_t nC = 0, nG = 0, nT = 0;
    for (; q <= q_last; ++q) {
        if (q == q_last) mask &= ~0ull >> (63 - (b_last & 63));
        const uint4 v = __ldg(planes + q);
        const unsigned long long x0 = ((uint64_t(v.y) << 32) | v.x) & mask, x1 = ((uint64_t(v.w) << 32) | v.z) & mask;
        nT += __popcll(x0 & x1);
        nC += __popcll(x0 & ~x1);
        nG += __popcll(x1 & ~x0);
        mask = ~0ull;
    }
    const uint32_t nA = len - nC - nG - nT;
    return uint32_t(nA >= mcov) + uint32_t(nC >= mcov) + uint32_t(nG >= mcov) + uint32_t(nT >= mcov);
}

}  // namespace e2s
