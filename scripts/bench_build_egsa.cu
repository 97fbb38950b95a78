// Timing tool (not part of the library): the library's index builder (csrc/build_egsa.cu, its own radix sort) next to the
// same word loop driven by cub::DeviceRadixSort (what round 1 shipped), on the same reads, on the same box; the two
// indexes are compared.  Built by scripts/build_bench_tools.sh into ebwt2snp_b200/bin/bench_build_egsa.
//   bench_build_egsa R L [reps]     R reads of L bases sampled at 30x from a random genome
#include <cub/device/device_radix_sort.cuh>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ebwt2snp_b200/csrc/build_egsa.cu"

namespace e2s {
namespace {

__global__ void k_plain_keys(ReadsView v, const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ids, uint64_t n, uint32_t w,
                             uint64_t* __restrict__ keys) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r;
    uint32_t p;
    v.decode(ids[i], r, p);
    keys[i] = suffix_word(packed + v.row(r, v.start(r)), v.L, p, w);
}

// round 1's builder: one cub::DeviceRadixSort::SortPairs per key word on exactly the significant bits
cudaError_t build_with_cub(const uint8_t* d_reads, uint64_t R, uint32_t L, uint32_t* d_lcp, uint32_t* d_text, uint32_t* d_suff, uint8_t* d_bwt,
                           cudaStream_t stream) {
    ReadsView v{d_reads, nullptr, R, L, 0};
    const uint64_t n = R * (uint64_t(L) + 1);
    const uint32_t W = (L + 31) / 32;
    uint64_t *packed, *k0, *k1;
    uint32_t *i0, *i1, *bad;
    void* tmp;
    size_t tmp_bytes = 0;
    cudaMalloc(&packed, ((R * L >> 5) + 2 * R + 2) * 8);
    cudaMalloc(&k0, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&i0, n * 4); cudaMalloc(&i1, n * 4); cudaMalloc(&bad, 4);
    cub::DoubleBuffer<uint64_t> keys(k0, k1);
    cub::DoubleBuffer<uint32_t> ids(i0, i1);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, ids, n, 0, 64, stream);
    cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
    k_pack_reads<<<blocks_for(R * 32, 256), 256, 0, stream>>>(v, packed, bad);
    k_init_ids_equal<uint32_t><<<blocks_for(n, 256), 256, 0, stream>>>(ids.Current(), n);
    for (int w = int(W) - 1; w >= 0; --w) {
        k_plain_keys<<<blocks_for(n, 256), 256, 0, stream>>>(v, packed, ids.Current(), n, uint32_t(w), keys.Current());
        const uint32_t syms = L - 32 * uint32_t(w) < 32 ? L - 32 * uint32_t(w) : 32;
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, ids, n, int(64 - 2 * syms), 64, stream);
    }
    k_egsa_finish<uint32_t><<<blocks_for(n, 256), 256, 0, stream>>>(v, packed, ids.Current(), n, d_lcp, d_text, d_suff, d_bwt, ~uint64_t(0));
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(packed); cudaFree(k0); cudaFree(k1); cudaFree(i0); cudaFree(i1); cudaFree(bad); cudaFree(tmp);
    return e;
}

}  // namespace
}  // namespace e2s

static uint64_t rng_state = 88172645463325252ull;
static inline uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(int argc, char** argv) {
    const uint64_t R = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
    const uint32_t L = argc > 2 ? uint32_t(atoi(argv[2])) : 100;
    const int reps = argc > 3 ? atoi(argv[3]) : 3;
    const uint64_t G = R * L / 30 + L, n = R * (uint64_t(L) + 1);
    std::vector<uint8_t> g(G), reads(R * L);
    for (auto& c : g) c = "ACGT"[rnd() & 3];
    for (uint64_t r = 0; r < R; ++r) {
        const uint64_t st = rnd() % (G - L + 1);
        for (uint32_t j = 0; j < L; ++j) reads[r * L + j] = g[st + j];
    }
    uint8_t *d_reads, *d_bwt[2];
    uint32_t *d_lcp[2], *d_text[2], *d_suff[2];
    cudaMalloc(&d_reads, R * L);
    cudaMemcpy(d_reads, reads.data(), R * L, cudaMemcpyHostToDevice);
    for (int v = 0; v < 2; ++v) { cudaMalloc(&d_bwt[v], n); cudaMalloc(&d_lcp[v], n * 4); cudaMalloc(&d_text[v], n * 4); cudaMalloc(&d_suff[v], n * 4); }
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best[2] = {1e30f, 1e30f};
    uint64_t launches = 0;
    for (int it = 0; it < reps + 1; ++it) {
        for (int v = 0; v < 2; ++v) {
            cudaEventRecord(a, s);
            cudaError_t e = v == 0 ? e2s::build_egsa(d_reads, nullptr, nullptr, R, L, R * L, d_lcp[0], d_text[0], d_suff[0], d_bwt[0], s, &launches)
                                   : e2s::build_with_cub(d_reads, R, L, d_lcp[1], d_text[1], d_suff[1], d_bwt[1], s);
            cudaEventRecord(b, s);
            cudaEventSynchronize(b);
            if (e != cudaSuccess) { printf("{\"error\": \"%s\", \"variant\": %d}\n", cudaGetErrorString(e), v); return 1; }
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it > 0 && ms < best[v]) best[v] = ms;
        }
    }
    std::vector<uint32_t> h0(n), h1(n);
    bool same = true;
    uint32_t* pairs[3][2] = {{d_lcp[0], d_lcp[1]}, {d_text[0], d_text[1]}, {d_suff[0], d_suff[1]}};
    for (auto& p : pairs) {
        cudaMemcpy(h0.data(), p[0], n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h1.data(), p[1], n * 4, cudaMemcpyDeviceToHost);
        same = same && h0 == h1;
    }
    printf("{\"tool\": \"bench_build_egsa\", \"reads\": %llu, \"read_len\": %u, \"suffixes\": %llu, \"own_radix_ms\": %.3f, \"cub_radix_ms\": %.3f, "
           "\"own_suffixes_per_s\": %.4g, \"cub_suffixes_per_s\": %.4g, \"indexes_equal\": %s, \"timing\": \"best of %d, allocation of the scratch included in both\"}\n",
           (unsigned long long)R, L, (unsigned long long)n, best[0], best[1], n / (best[0] * 1e-3), n / (best[1] * 1e-3), same ? "true" : "false", reps);
    return same ? 0 : 2;
}
