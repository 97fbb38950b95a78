// Bit planes of the 2-bit base code (base_to_int, ref:include.hpp:265-279: A=0 C=1 G=2 T=3, every other byte 0)
// and the BWT-only prefilter of find_variants built on them.  Shared by K3a (k_code_scan, cluster-list driven) and
// by K2's fused mode (k_cluster_emit, END driven).
#pragma once

#include <stdint.h>

#include "common.cuh"

namespace e2s {

// 0x80 in every byte of u that equals the corresponding byte of pat (exact, no cross-byte borrow)
__device__ __forceinline__ uint32_t eq_bytes(uint32_t u, uint32_t pat) {
    const uint32_t t = u ^ pat;
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
}
// four byte flags (0x80 each) -> 4 bits
__device__ __forceinline__ uint32_t nibble(uint32_t f) { return (((f >> 7) * 0x01020408u) >> 24) & 0xFu; }

// Bit-sliced base code of four bytes, valid for the bytes k_bwt_alphabet accepts: with u = byte & 0xDF,
//   bit 1 of the code (G, T) = u.bit2 & u.bit6        bit 0 (C, T) = (u.bit1 & ~u.bit2) | u.bit4
// (A = 0x41, C = 0x43, G = 0x47, T = 0x54; '$' = 0x24 and NUL give 0 like every non-ACGT byte must).
__device__ __forceinline__ uint32_t fast_b1(uint32_t w) {  // bit 0 of each byte
    const uint32_t u = w & 0xDFDFDFDFu;
    return (u >> 2) & (u >> 6) & 0x01010101u;
}
__device__ __forceinline__ uint32_t fast_b0(uint32_t w) {
    const uint32_t u = w & 0xDFDFDFDFu;
    return (((u >> 1) & ~(u >> 2)) | (u >> 4)) & 0x01010101u;
}
// both at once: plane 0 at bit 0 and plane 1 at bit 4 of every byte
__device__ __forceinline__ uint32_t fast_b01(uint32_t w) {
    const uint32_t u = w & 0xDFDFDFDFu, b = u >> 2;
    const uint32_t x = ((u >> 1) & ~b) | (u >> 4);
    const uint32_t y = (u << 2) & b;
    return (x & 0x01010101u) | (y & 0x10101010u);
}

// Planes of `n_chunks` 16-byte chunks staged in shared memory at `st` (16-byte aligned): s_b0 / s_b1 get one 16-bit
// word per chunk (bit i = position 16 * chunk + i).  simple = the shard's BWT was proven (k_bwt_alphabet, at seal) to
// hold only bytes on which the bit-sliced code is exact; otherwise per-byte equality tests.
__device__ __forceinline__ void build_planes(const uint8_t* st, int n_chunks, uint16_t* s_b0, uint16_t* s_b1, bool simple,
                                             int tid, int n_threads) {
    if (simple) {
        for (int ch = tid; ch < n_chunks; ch += n_threads) {
            const uint4 q = lds128(st + ch * 16);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            // one multiply per word gathers both planes: with the flags of plane 0 at bit 0 and those of plane 1 at bit 4
            // of every byte, (x * 0x01020408) >> 24 = plane-1 nibble << 4 | plane-0 nibble (no two partial products meet)
            uint32_t B = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) B |= ((fast_b01(w[j]) * 0x01020408u) >> 24) << (8 * j);
            uint32_t lo = B & 0x0F0F0F0Fu, hi = (B >> 4) & 0x0F0F0F0Fu;  // byte j = nibble of word j
            lo = (lo | (lo >> 4)) & 0x00FF00FFu;
            hi = (hi | (hi >> 4)) & 0x00FF00FFu;
            s_b0[ch] = uint16_t(lo | (lo >> 8));
            s_b1[ch] = uint16_t(hi | (hi >> 8));
        }
    } else {
        for (int ch = tid; ch < n_chunks; ch += n_threads) {
            const uint4 q = lds128(st + ch * 16);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t b0 = 0, b1 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t u = w[j] & 0xDFDFDFDFu;  // case-insensitive
                const uint32_t eC = eq_bytes(u, 0x43434343u), eG = eq_bytes(u, 0x47474747u), eT = eq_bytes(u, 0x54545454u);
                b0 |= nibble(eC | eT) << (4 * j);
                b1 |= nibble(eG | eT) << (4 * j);
            }
            s_b0[ch] = uint16_t(b0);
            s_b1[ch] = uint16_t(b1);
        }
    }
}

// The prefilter: how many base codes occur >= mcov times among positions [lo, hi) of the planes (32-bit views w0, w1).
// A cluster with at most one such code cannot pass find_variants (ref:clust2snp.cpp:402-429): counts[s][c] <= total[c].
__device__ __forceinline__ uint32_t frequent_codes(const uint32_t* w0, const uint32_t* w1, uint32_t lo, uint32_t hi, uint32_t mcov) {
    uint32_t nC = 0, nG = 0, nT = 0;
    auto add = [&](uint32_t wi, uint32_t mask) {
        const uint32_t x0 = w0[wi] & mask, x1 = w1[wi] & mask;
        nT += __popc(x0 & x1);
        nC += __popc(x0 & ~x1);
        nG += __popc(x1 & ~x0);
    };
    const uint32_t wlo = lo >> 5, whi = (hi - 1) >> 5;
    const uint32_t m_first = 0xffffffffu << (lo & 31), m_last = 0xffffffffu >> (31 - ((hi - 1) & 31));
    if (wlo == whi) {
        add(wlo, m_first & m_last);
    } else {
        add(wlo, m_first);
        for (uint32_t wi = wlo + 1; wi < whi; ++wi) add(wi, 0xffffffffu);
        add(whi, m_last);
    }
    const uint32_t nA = (hi - lo) - nC - nG - nT;
    return uint32_t(nA >= mcov) + uint32_t(nC >= mcov) + uint32_t(nG >= mcov) + uint32_t(nT >= mcov);
}

}  // namespace e2s
