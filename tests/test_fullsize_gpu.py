"""-m gpu: one shard of more than 2^32 eBWT positions (C3 / C4 territory: 64-bit positions, 59 GB resident).

The oracle cannot hold such an input, so the check is through size-independent properties (task brief, parity (3)):
the eBWT is the C2 read set tiled 8 times with shifted read ids.  Every tile starts with lcp = 0, so no cluster spans
a tile boundary and the records of tile t must be the records of tile 0 shifted by t * n -- including the tiles that
lie beyond position 2^32.  Tile 0's own records are pinned to the oracle on a prefix the oracle can hold."""
import numpy as np
import pytest

from ebwt2snp_b200 import api, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TILES = 8


def test_more_than_2_32_positions(built):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 120 << 30:
        pytest.skip("needs a 180 GB B200")
    dev = torch.device("cuda", 0)
    rs = synth.make_config("C2", seed=1)
    e = synth.build_egsa(rs.reads, device=dev)
    n = int(e["n"])
    R = rs.reads.shape[0]
    N = TILES * n
    assert N > 1 << 32
    ctx = api.Context(0)
    sh = ctx.shard(N)
    for t in range(TILES):
        sh.load_soa(e["lcp"], e["text"] + t * R, e["suff"], e["bwt"], first=t * n, device=True)
    sh.seal()
    k, m = 16, 2
    nw, nc = sh.cluster_lm(k, m)
    start, ln = sh.cluster_fetch()
    assert len(start) == nw and np.all(np.diff(start.astype(np.int64)) > 0)          # position-ordered
    assert np.all(start[:-1] + ln[:-1].astype(np.uint64) <= start[1:])               # disjoint
    assert int(start[-1]) > 1 << 32

    # tile 0 against the oracle on a prefix (the tile ends with lcp = 0 of tile 1, the prefix with its own tail rule: compare away from it)
    npre = 3_000_000
    lcp0 = e["lcp"][:npre].cpu().numpy().view(np.uint32)
    bwt0 = e["bwt"][:npre].cpu().numpy()
    es, el, _, _ = O.cluster_lm(lcp0, bwt0, k, m)
    keep = es < npre - 1000
    sel = start < npre - 1000
    assert np.array_equal(start[sel], es[keep]) and np.array_equal(ln[sel], el[keep])

    # every tile repeats tile 0 (away from the last 1000 positions of a tile, where the tail rule of the LAST tile differs)
    t0 = start < n - 1000
    base_s, base_l = start[t0], ln[t0]
    for t in range(1, TILES):
        lo, hi = np.uint64(t * n), np.uint64((t + 1) * n - 1000)
        sel = (start >= lo) & (start < hi)
        assert np.array_equal(start[sel] - lo, base_s), t
        assert np.array_equal(ln[sel], base_l), t

    # phase 2: with one -n threshold only tile 0 holds reads of sample 0, so only tile 0 can produce candidates;
    # they must be the ones a run over tile 0 alone finds (away from the tile's end)
    reads_dev = torch.from_numpy(rs.reads).to(dev)
    L = rs.reads.shape[1]
    all_reads = reads_dev.repeat(TILES, 1).contiguous().view(-1)
    off = torch.arange(TILES * R + 1, dtype=torch.int64, device=dev) * L
    ctx.stage_reads(all_reads, off, device=True, n_bases=TILES * R * L)
    p = api.default_params(rs.nreads1)
    st = sh.statistics(p.mcov_out, p.pval)
    cnt = sh.find_events(p, st.max_clust_length)
    ev_big = [(ev.cluster_start, ev.D, ev.gap, ev.supp0, ev.supp1, bytes(ev.left0), bytes(ev.left1), bytes(ev.right))
              for ev in sh.events()]
    assert all(cs < n for cs, *_ in ev_big)
    sh.close()

    sh1 = ctx.shard(n)
    sh1.load_soa(e["lcp"], e["text"], e["suff"], e["bwt"], device=True)
    sh1.seal()
    sh1.cluster_lm(k, m)
    cnt1 = sh1.find_events(p, st.max_clust_length)
    ev_one = [(ev.cluster_start, ev.D, ev.gap, ev.supp0, ev.supp1, bytes(ev.left0), bytes(ev.left1), bytes(ev.right))
              for ev in sh1.events()]
    sh1.close()
    cut = n - 1000
    assert [x for x in ev_big if x[0] < cut] == [x for x in ev_one if x[0] < cut]
    assert cnt.n_candidates > 1000 and abs(int(cnt.n_candidates) - int(cnt1.n_candidates)) <= 2
    ctx.close()
