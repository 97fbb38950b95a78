"""Host-side logic without a GPU: e2s_cluster_merge (shard chaining, head records, tail + phantom rule)
and e2s_statistics_finish, driven by a numpy emulation of the per-shard scan results."""
import numpy as np
import pytest

from ebwt2snp_b200 import api
from oracle import oracle as O
from tests import helpers as H


def random_cuts(rng, n, nsh):
    cuts = sorted(set([0, n] + [int(c) for c in rng.integers(2, max(3, n - 2), size=nsh - 1)]))
    out = [0]
    for c in cuts[1:]:
        if c - out[-1] >= 2:
            out.append(c)
    if out[-1] != n:
        if n - out[-1] < 2 and len(out) > 1:
            out[-1] = n
        else:
            out.append(n)
    return out


@pytest.mark.parametrize("seed", range(6))
def test_merge_matches_oracle(built, seed):
    rng = np.random.default_rng(seed)
    for it in range(120):
        n = int(rng.integers(4, 80)) if it % 2 else int(rng.integers(100, 5000))
        k = int(rng.choice([1, 2, 3, 5, 16, 70]))
        m = int(rng.choice([1, 2, 3, 8]))
        lcp = H.random_lcp(rng, n, k, it % 5)
        bwt = rng.choice(H.BWT_ALPHABET, size=n)
        es, el, enc, eph = O.cluster_lm(lcp, bwt, k, m)
        cuts = random_cuts(rng, n, int(rng.integers(1, 7)))
        sums, recs = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            s, rs, rl = H.emulate_shard(lcp, bwt, lo, hi, k, m)
            sums.append(s)
            recs.append((rs, rl))
        S, L, mg = H.assemble(sums, recs)
        assert (mg.n_clust_out & 0xFFFFFFFF) == enc, (seed, it, cuts)
        assert mg.phantom_lcp == eph
        assert np.array_equal(S, es) and np.array_equal(L, el), (seed, it, cuts)


def test_adversarial_cuts(built):
    """cut at i, i+-1, i+-2 of every END and inside the last 3 positions"""
    rng = np.random.default_rng(42)
    n, k, m = 400, 5, 2
    lcp = H.random_lcp(rng, n, k, 2)
    bwt = rng.choice(H.BWT_ALPHABET, size=n)
    es, el, enc, _ = O.cluster_lm(lcp, bwt, k, m)
    _, end = H.flags(lcp, k)
    cutset = set()
    for e in np.flatnonzero(end):
        for d in (-2, -1, 0, 1, 2):
            if 2 <= e + d <= n - 2:
                cutset.add(int(e + d))
    cutset |= {n - 3, n - 2}
    for c in sorted(cutset):
        sums, recs = [], []
        for lo, hi in ((0, c), (c, n)):
            s, rs, rl = H.emulate_shard(lcp, bwt, lo, hi, k, m)
            sums.append(s)
            recs.append((rs, rl))
        S, L, mg = H.assemble(sums, recs)
        assert mg.n_clust_out == enc and np.array_equal(S, es) and np.array_equal(L, el), c


def test_adoption_covers_every_record(built):
    """phase 2 ownership: every written record is analysed by exactly one shard (the one holding its START)"""
    rng = np.random.default_rng(3)
    for it in range(60):
        n = int(rng.integers(50, 3000))
        k, m = 4, 2
        lcp = H.random_lcp(rng, n, k, it % 5)
        bwt = rng.choice(H.BWT_ALPHABET, size=n)
        es, el, _, _ = O.cluster_lm(lcp, bwt, k, m)
        cuts = random_cuts(rng, n, int(rng.integers(2, 6)))
        sums, recs = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            s, rs, rl = H.emulate_shard(lcp, bwt, lo, hi, k, m)
            sums.append(s)
            recs.append((rs, rl))
        seen = []
        for g, (lo, hi) in enumerate(zip(cuts[:-1], cuts[1:])):
            mg = api.cluster_merge(sums, g)
            mine = list(zip(recs[g][0].tolist(), recs[g][1].tolist()))
            mine += [(mg.adopt_start[i], mg.adopt_len[i]) for i in range(mg.n_adopt)]
            for st, ln in mine:
                assert lo <= st < hi or (g == len(cuts) - 2 and st == n)
            assert [a for a, _ in mine] == sorted(a for a, _ in mine)
            seen += mine
        assert sorted(seen) == sorted(zip(es.tolist(), el.tolist()))


def test_statistics_finish(built):
    rng = np.random.default_rng(0)
    for it in range(50):
        m = int(rng.integers(1, 400))
        ln = rng.integers(1, 200 if it % 2 else 70000, size=m).astype(np.uint16)
        start = np.cumsum(np.concatenate([[0], ln[:-1].astype(np.int64)])).astype(np.uint64)
        mcov = int(rng.choice([1, 3, 5, 10]))
        pval = float(rng.choice([0.99, 0.5, 0.1, 0.9999]))
        ost = O.statistics(start, ln, mcov, pval)
        st = api.Stats()
        for v in ln:
            if v <= 150:
                st.hist[int(v)] += 1
        st.n_clust, st.n_bases = m, int(ln.astype(np.int64).sum())
        api.statistics_finish(st, int(ln[-1]), mcov, pval)
        assert list(st.hist) == list(ost.hist)
        assert (st.n_clust, st.n_bases, st.max_len, st.max_clust_length) == (
            ost.n_clust, ost.n_bases, ost.max_len, ost.max_clust_length)


def test_exchange_rows_equal_exchange_finish(built):
    """the library-side exchange (e2s_pipeline_sharded) hands every rank the raw rows of all shards -- the scan's device
    accumulators + the shard's range; e2s_exchange_rows_finish must turn them into the same merged view and global
    statistics as e2s_exchange_finish gets from summaries + own-record statistics (the torch.distributed path)"""
    import ctypes as C
    W = api.exchange_row_words()
    assert W == C.sizeof(api.ClusterDev) // 8 + 4
    rng = np.random.default_rng(31)
    for it in range(60):
        n = int(rng.integers(200, 4000))
        k = int(rng.choice([2, 5, 16]))
        m = int(rng.choice([1, 2, 3]))
        lcp = H.random_lcp(rng, n, k, it % 5)
        bwt = rng.choice(H.BWT_ALPHABET, size=n)
        cuts = random_cuts(rng, n, int(rng.integers(1, 6)))
        G = len(cuts) - 1
        sums = (api.ClusterSummary * G)()
        stats = (api.Stats * G)()
        rows = np.zeros((G, W), dtype=np.uint64)
        for g, (lo, hi) in enumerate(zip(cuts[:-1], cuts[1:])):
            s, rs, rl = H.emulate_shard(lcp, bwt, lo, hi, k, m)
            sums[g] = s
            st = stats[g]
            d = api.ClusterDev()
            for f in ("n_end", "n_written", "head_end", "any_event", "open_start", "end_nm2_start", "tail_lcp_nm2", "tail_lcp_nm1",
                      "tail_bwt_nm1"):
                setattr(d, f, getattr(s, f))
            for l in rl.tolist():
                if l <= api.MAX_C_LEN:
                    d.hist[l] += 1
                    st.hist[l] += 1
            d.n_bases = st.n_bases = int(rl.astype(np.uint64).sum())
            st.n_clust = len(rl)
            st.last_len = int(rl[-1]) if len(rl) else 0
            d.last_rec = (len(rl) << 16) | (int(rl[-1]) if len(rl) else 0)
            d.ticket, d.n_pf = 12345, 7  # scratch values of the scan: must not matter
            rows[g, :W - 4] = np.frombuffer(bytes(d), dtype=np.uint64)
            rows[g, W - 4:] = (hi - lo, lo, 4, 0)
        for my in range(G):
            try:
                want = api.exchange_finish(sums, stats, my)
            except api.E2SError:
                with pytest.raises(api.E2SError):
                    api.exchange_rows_finish(rows, my, n, k, m)
                continue
            got = api.exchange_rows_finish(rows, my, n, k, m)
            assert bytes(got[0]) == bytes(want[0]) and bytes(got[1]) == bytes(want[1]), (it, my, cuts)
