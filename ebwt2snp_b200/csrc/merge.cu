// k_merge_stats: the exchange between the phases, on the device (one thread; the logic is merge.cuh's, shared with the host
// entry points).  Input: one exchange row per shard -- the scan's accumulators (ClusterDev: counters, open-cluster state,
// tail values, length histogram of the shard's own records) + the shard's range -- as the NCCL all-gather left them in
// device memory, or this shard's own accumulators when there is one shard.  Output: this shard's merged view (record offset,
// head / tail records, adopted records), the global statistics() with max_clust_length, a status word; the records this
// shard adopts are appended to the survivor list phase 2 starts from.  Phase 2's exact test reads max_clust_length from
// here, so nothing between the scan and the events needs the host.

#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"
#include "merge.cuh"

namespace e2s {

__device__ __forceinline__ void summary_of_row(const unsigned long long* row, uint64_t n_global, uint32_t k, int32_t min_len,
                                               e2s_cluster_summary* sum) {
    const ClusterDev& h = *reinterpret_cast<const ClusterDev*>(row);
    *sum = e2s_cluster_summary{};
    sum->n_local = row[XR_DEV_WORDS + 0];
    sum->global_off = row[XR_DEV_WORDS + 1];
    sum->n_global = n_global;
    sum->n_end = h.n_end;
    sum->n_written = h.n_written;
    sum->head_end = h.head_end;
    sum->any_event = h.any_event;
    sum->open_start = h.any_event ? h.open_start : 0;
    sum->end_nm2_start = h.end_nm2_start;
    sum->k = k;
    sum->min_len = uint64_t(int64_t(min_len));
    sum->lcp_bytes = row[XR_DEV_WORDS + 2];
    sum->tail_lcp_nm2 = h.tail_lcp_nm2;
    sum->tail_lcp_nm1 = h.tail_lcp_nm1;
    sum->tail_bwt_nm1 = h.tail_bwt_nm1;
}

__global__ void k_pack_exchange(const ClusterDev* res, uint64_t n_local, uint64_t global_off, uint64_t lcp_bytes, unsigned long long* row) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(res);
    for (uint32_t i = threadIdx.x; i < XR_DEV_WORDS; i += blockDim.x) row[i] = src[i];
    if (threadIdx.x == 0) {
        row[XR_DEV_WORDS + 0] = n_local;
        row[XR_DEV_WORDS + 1] = global_off;
        row[XR_DEV_WORDS + 2] = lcp_bytes;
        row[XR_DEV_WORDS + 3] = 0;
    }
}

__global__ void __launch_bounds__(128) k_merge_stats(MergeParams p) {
    __shared__ e2s_cluster_summary sums[MERGE_MAX_SHARDS];
    __shared__ e2s_stats tot;
    __shared__ unsigned long long s_cum[E2S_HIST_BINS];
    __shared__ int s_status, s_mcl;
    const int tid = threadIdx.x;
    // the rows -> summaries (one thread per shard) and the sum of the shards' own-record histograms (one thread per bin):
    // everything the serial part below touches is then in shared memory
    for (int g = tid; g < p.world; g += blockDim.x) summary_of_row(p.rows + size_t(g) * XR_WORDS, p.n_global, p.k, p.min_len, &sums[g]);
    for (int i = tid; i < E2S_HIST_BINS; i += blockDim.x) {
        unsigned long long v = 0;
        for (int g = 0; g < p.world; ++g) v += reinterpret_cast<const ClusterDev*>(p.rows + size_t(g) * XR_WORDS)->hist[i];
        tot.hist[i] = v;
    }
    if (tid == 0) {
        tot.n_clust = tot.n_bases = tot.last_len = tot.max_len = 0;
        tot.max_clust_length = tot.reserved = 0;
    }
    __syncthreads();
    MergeOut* out = p.out;
    if (tid == 0) {
        uint64_t last_len = 0;
        bool any = false;
        int status = MERGE_OK;
        for (int g = 0; g < p.world && status == MERGE_OK; ++g) {
            e2s_cluster_merged mg;
            status = merge_core(sums, p.world, g, &mg);
            if (status != MERGE_OK) break;
            if (g == p.my) out->mine = mg;
            // records of shard g in file order: [head record] own records [tail records]
            const ClusterDev& h = *reinterpret_cast<const ClusterDev*>(p.rows + size_t(g) * XR_WORDS);
            uint64_t n_clust = h.n_written, n_bases = h.n_bases, ll = h.last_rec & 0xffff;
            auto add = [&](uint64_t l, bool is_last) {
                if (l <= E2S_MAX_C_LEN) tot.hist[l]++;
                n_bases += l;
                n_clust++;
                if (is_last) ll = l;
            };
            if (mg.n_prepend && mg.prepend_written) add(mg.prepend_len, h.n_written == 0);
            for (uint32_t i = 0; i < mg.n_append; ++i) add(mg.append_len[i], true);
            tot.n_clust += n_clust;
            tot.n_bases += n_bases;
            if (n_clust) {
                last_len = ll;
                any = true;
            }
        }
        tot.last_len = last_len;
        if (status == MERGE_OK) status = any ? stats_finish_head(&tot, last_len, p.mcov) : MERGE_EMPTY;
        s_status = status;
        if (status == MERGE_OK) {  // the pval loop's running sum for every candidate max_clust_length (integers: cheap in one thread)
            const int m0 = 2 * p.mcov;
            unsigned long long cum = 0;
            for (int i = m0; i <= E2S_MAX_C_LEN; ++i) {
                cum += tot.hist[i] * (unsigned long long)i;
                s_cum[i] = cum;
            }
            s_mcl = E2S_MAX_C_LEN;
        }
        out->status = status;
        if (status == MERGE_OK && p.pf_list) {  // the records this shard adopts go to the exact test unconditionally
            for (uint32_t i = 0; i < out->mine.n_adopt; ++i) {
                const unsigned long long at = p.res->n_pf++;
                if (at < p.pf_cap)
                    p.pf_list[at] = SurvEntry{out->mine.adopt_start[i], out->mine.adopt_start[i] - p.own_global_off, uint32_t(out->mine.adopt_len[i]), 0u};
            }
        }
    }
    __syncthreads();
    if (s_status == MERGE_OK) {
        // the pval loop of statistics() (ref:clust2snp.cpp:938-946; stats_finish_core): max_clust_length = the first mcl >= 2 mcov
        // whose cumulative / n_bases is not below pval (or the cap).  One IEEE double division per candidate, the same operation the
        // serial loop performs, here one candidate per thread: a dependent chain of ~60 divisions took 10 us in one thread.
        const int m0 = 2 * p.mcov;
        for (int i = m0 + tid; i < E2S_MAX_C_LEN; i += blockDim.x)
            if (!(double(s_cum[i]) / double(tot.n_bases) < p.pval)) atomicMin(&s_mcl, i);
        __syncthreads();
        if (tid == 0) tot.max_clust_length = s_mcl;
        __syncthreads();
    }
    {   // statistics() out, max_clust_length included (phase 2's exact test reads it from here)
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&tot);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&out->total);
        for (uint32_t i = tid; i < sizeof(e2s_stats) / 8; i += blockDim.x) dst[i] = src[i];
    }
}

cudaError_t launch_pack_exchange(const ClusterDev* res, uint64_t n_local, uint64_t global_off, uint64_t lcp_bytes, uint64_t* row,
                                 cudaStream_t stream) {
    k_pack_exchange<<<1, 256, 0, stream>>>(res, n_local, global_off, lcp_bytes, reinterpret_cast<unsigned long long*>(row));
    return cudaGetLastError();
}

cudaError_t launch_merge_stats(const MergeParams& p, cudaStream_t stream) {
    if (p.world < 1 || p.world > MERGE_MAX_SHARDS) return cudaErrorInvalidValue;
    k_merge_stats<<<1, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace e2s
