"""Multi-GPU plumbing: one process per GPU, contiguous eBWT shards.

The path shards by contiguous eBWT ranges and has **no data-path collective**: the only exchanges are
a few words per rank between the phases,

    1. halo records (2 left / 151 right) once, when the shards are loaded,
    2. the 14-word scan summary of every shard        -> e2s_cluster_merge (host integer logic),
    3. the 151-bin length histogram + 3 counters      -> e2s_statistics_finish (max_clust_length),
    4. the kept-event count                            -> global id_nr offsets of the .snp records,

all done with `torch.distributed.all_gather` on small int64 tensors (NCCL over NVLink on the GPUs,
gloo in the CPU tests).  The functions take ctypes structs of `api` in and out so that the same code
runs in `bench.py` under torchrun and in `tests/test_sharding_gloo.py` without a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import api

HALO_L = 2                      # END(i) needs lcp[i-1], START(i) needs END(i-1)
HALO_R = api.MAX_C_LEN + 1      # clusters analysed in phase 2 are <= 150 long (+1 for the LCP stencil)


def shard_cuts(n, parts):
    """contiguous, nearly equal shards, every one with >= 2 positions (mirrors host_io.hpp:shard_cuts)"""
    parts = max(1, int(parts))
    while parts > 1 and n // parts < 2:
        parts -= 1
    base, rem = divmod(n, parts)
    cuts = [base * g + min(g, rem) for g in range(parts)] + [n]
    return cuts


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_words(words, device, group=None):
    """all-gather of a short list of unsigned 64-bit words; returns one list of python ints per rank"""
    rank, world = _world(group)
    w = np.asarray(words, dtype=np.uint64)
    if world == 1:
        return [[int(x) for x in w]]
    mine = torch.from_numpy(w.view(np.int64).copy()).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [[int(x) for x in t.cpu().numpy().view(np.uint64)] for t in out]


def exchange_summaries(summary: api.ClusterSummary, device, group=None):
    """step 2: every rank gets the scan summaries of all shards, in rank (= eBWT) order"""
    words = np.frombuffer(bytes(summary), dtype=np.uint64)
    gathered = all_gather_words(words, device, group)
    return [api.ClusterSummary.from_buffer_copy(np.asarray(g, dtype=np.uint64).tobytes()) for g in gathered]


def merge_clusters(summary: api.ClusterSummary, device, group=None):
    """-> (merged view of this rank, list of all summaries); ref:ebwt2clust.cpp:90-135 across shards"""
    rank, world = _world(group)
    sums = [summary] if world == 1 else exchange_summaries(summary, device, group)
    return api.cluster_merge(sums, rank), sums


def merge_statistics(st: api.Stats, mcov_out, pval, device, group=None) -> api.Stats:
    """step 3: global statistics() = sum of the shards' histograms; the reference's double count of the
    last record (ref:clust2snp.cpp:889) uses the LAST record of the global file, i.e. of the last
    shard that has any record.  Runs the pval loop (one IEEE double division per step, on the host)."""
    if _world(group)[1] == 1:
        api.statistics_finish(st, st.last_len, mcov_out, pval)
        return st
    words = list(st.hist) + [st.n_clust, st.n_bases, st.last_len]
    gathered = all_gather_words(words, device, group)
    tot = api.Stats()
    last_len = 0
    for g in gathered:
        for i in range(api.HIST_BINS):
            tot.hist[i] += g[i]
        tot.n_clust += g[api.HIST_BINS]
        tot.n_bases += g[api.HIST_BINS + 1]
        if g[api.HIST_BINS] > 0:
            last_len = g[api.HIST_BINS + 2]
    tot.last_len = last_len
    api.statistics_finish(tot, last_len, mcov_out, pval)
    return tot


def event_id_offset(n_events, device, group=None):
    """step 4: -> (first id_nr of this rank's kept events, total kept events); ids start at 1 (ref:clust2snp.cpp:637)"""
    rank, world = _world(group)
    if world == 1:
        return 1, int(n_events)
    counts = [g[0] for g in all_gather_words([int(n_events)], device, group)]
    return 1 + sum(counts[:rank]), sum(counts)


def exchange_halo(lcp, text, suff, bwt, device, group=None):
    """step 1 for shards that were BORN on their GPUs (bench.py): returns
    (left: dict of 2-element tensors from the previous rank or None,
     right: dict of 151-element tensors from the next rank or None).
    Inputs are this rank's torch tensors (int32 lcp/text/suff, uint8 bwt) of >= 151 elements."""
    rank, world = _world(group)
    if world == 1:
        return None, None
    n = lcp.numel()
    if n < HALO_R:
        raise ValueError("shard shorter than the halo")

    def pack(a, b):
        return torch.cat([lcp[a:b].to(torch.int32), text[a:b].to(torch.int32), suff[a:b].to(torch.int32),
                          bwt[a:b].to(torch.int32)])

    mine = torch.cat([pack(n - HALO_L, n), pack(0, HALO_R)]).to(device)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)

    def unpack(v, w):
        v = v.view(4, w)
        return {"lcp": v[0].contiguous(), "text": v[1].contiguous(), "suff": v[2].contiguous(),
                "bwt": v[3].to(torch.uint8).contiguous()}

    left = unpack(allv[rank - 1][: 4 * HALO_L], HALO_L) if rank > 0 else None
    right = unpack(allv[rank + 1][4 * HALO_L:], HALO_R) if rank < world - 1 else None
    return left, right


def hot_path_step(shard: "api.Shard", params: api.SnpParams, k, min_len, device, group=None):
    """One pass of the hot path on this rank's resident shard, collectives included.
    -> (ClusterMerged, Stats (global), SnpCounts (this shard), first event id of this shard)"""
    s = shard.cluster_run(k, min_len)
    mg, _ = merge_clusters(s, device, group)
    shard.cluster_finalize(mg)
    st = shard.statistics(finish=False)
    st = merge_statistics(st, params.mcov_out, params.pval, device, group)
    cnt = shard.find_events(params, st.max_clust_length)
    first_id, _ = event_id_offset(cnt.n_events, device, group)
    return mg, st, cnt, first_id
