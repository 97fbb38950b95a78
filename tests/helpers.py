"""Test-side helpers (numpy): random LCP arrays, a vectorised shard emulation of what the CUDA scan
reports for one shard (used to test the host-side merge logic without a GPU), dataset cache."""
from __future__ import annotations

import os

import numpy as np

from ebwt2snp_b200 import api, synth

BWT_ALPHABET = np.frombuffer(b"ACGT$acgt\x00\xff", dtype=np.uint8)


def random_lcp(rng, n, k, mode):
    if mode == 0:
        lcp = rng.integers(0, 2 * k + 2, size=n)
    elif mode == 1:
        lcp = rng.integers(0, 70000, size=n)
    elif mode == 2:
        lcp = np.maximum(0, np.cumsum(rng.integers(-3, 4, size=n)) + k)
    elif mode == 3:
        lcp = rng.integers(k, k + 3, size=n)  # long runs, few local minima
    else:
        lcp = np.full(n, k + 1)  # one giant cluster
        lcp[rng.integers(0, n, size=max(1, n // 5000))] = 0
    return lcp.astype(np.uint32)


def flags(lcp, k):
    """START/END masks of the stencil form for positions 0..n-1, END(n-1) left out."""
    n = len(lcp)
    l = lcp.astype(np.int64)
    ge = l >= k
    prev = np.concatenate([[0], l[:-1]])
    nxt = np.concatenate([l[1:], [0]])
    ge_n = np.concatenate([ge[1:], [False]])
    end = ge & (((prev > l) & (l <= nxt)) | ~ge_n)
    end[0] = False
    if n > 2 and ge[0] and not ge[1]:
        end[1] = True
    end[n - 1] = False
    ge_p = np.concatenate([[False], ge[:-1]])
    end_p = np.concatenate([[False], end[:-1]])
    start = ge & (~ge_p | end_p)
    return start, end


def emulate_shard(lcp, bwt, lo, hi, k, min_len, lcp_bytes=4):
    """What e2s_cluster_run reports for the shard [lo, hi) of the global arrays: (summary, start[], len[])."""
    n = len(lcp)
    start, end = flags(lcp, k)
    s = api.ClusterSummary()
    s.n_local, s.global_off, s.n_global = hi - lo, lo, n
    s.k, s.min_len = k, min_len & 0xFFFFFFFFFFFFFFFF
    s.lcp_bytes = lcp_bytes
    st_pos = np.flatnonzero(start[lo:hi]) + lo
    en_pos = np.flatnonzero(end[lo:hi]) + lo
    s.n_end = len(en_pos)
    s.any_event = int(len(st_pos) + len(en_pos) > 0)
    recs_s, recs_l = [], []
    ens = list(en_pos)
    if ens and (len(st_pos) == 0 or ens[0] < st_pos[0]):
        s.head_end = int(ens[0]) + 1
        if ens[0] == n - 2:
            s.end_nm2_start = 0xFFFFFFFFFFFFFFFF
        ens = ens[1:]
    for j, e in enumerate(ens):
        st = int(st_pos[j])
        ln = (int(e) - st + 1) & 0xFFFF
        if ln >= min_len:
            recs_s.append(st)
            recs_l.append(ln)
        if e == n - 2:
            s.end_nm2_start = st + 1
    s.n_written = len(recs_s)
    if len(st_pos) > len(ens):
        s.open_start = int(st_pos[-1]) + 1
    if hi == n:
        s.tail_lcp_nm2, s.tail_lcp_nm1, s.tail_bwt_nm1 = int(lcp[n - 2]), int(lcp[n - 1]), int(bwt[n - 1])
    return s, np.array(recs_s, dtype=np.uint64), np.array(recs_l, dtype=np.uint16)


def assemble(summaries, shard_records):
    """Global .clusters (start, len) from per-shard results + e2s_cluster_merge; also n_clust_out."""
    S, L = [], []
    mg = None
    for g, (rs, rl) in enumerate(shard_records):
        mg = api.cluster_merge(summaries, g)
        assert mg.record_offset == len(S)
        if mg.n_prepend and mg.prepend_written:
            S.append(mg.prepend_start)
            L.append(mg.prepend_len)
        S += list(rs)
        L += list(rl)
        for i in range(mg.n_append):
            S.append(mg.append_start[i])
            L.append(mg.append_len[i])
    assert mg.total_written == len(S)
    return np.array(S, dtype=np.uint64), np.array(L, dtype=np.uint16), mg


_DS_CACHE = {}


def dataset(name, seed=1, device="cpu"):
    """(ReadSet, egsa dict of numpy arrays) of a named config, cached per session."""
    key = (name, seed)
    if key not in _DS_CACHE:
        rs = synth.make_config(name, seed=seed)
        e = synth.build_egsa(rs.reads, device=device)
        eg = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in e.items()}
        for f in ("lcp", "text", "suff"):
            eg[f] = eg[f].view(np.uint32)
        _DS_CACHE[key] = (rs, eg)
    return _DS_CACHE[key]
