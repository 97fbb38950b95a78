"""-m gpu: chunked shards (streaming, BASELINE config 4's mode): a chunk is a shard in time.  Whatever the chunk size, the
bytes of .clusters and .snp must be the ones of the single resident pass and of the oracle."""
import numpy as np
import pytest

from ebwt2snp_b200 import api, synth
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context(0)
    yield c
    c.close()


def stream_phase1(ctx, lcp, bwt, k, m, chunk, lo=0, hi=None, text=None, suff=None, mcov=0):
    """the range [lo, hi) of the arrays through a chunked shard -> (summary, records of the range, shard)"""
    n = len(lcp)
    hi = n if hi is None else hi
    sh = ctx.shard(hi - lo, lo, n, chunk_positions=chunk)
    S, L = [], []
    for clo, cn in sh.chunks():
        sh.chunk_begin(clo, cn)
        a, b = max(0, clo - 176), min(n, clo + cn + 152)
        sh.load_soa(lcp[a:b], None if text is None else text[a:b], None if suff is None else suff[a:b], bwt[a:b], first=a)
        cnt = sh.chunk_scan(k, m, mcov)
        s, l = sh.cluster_fetch()
        assert len(s) == cnt
        S.append(s)
        L.append(l)
    return sh.chunked_finish(k, m), np.concatenate(S), np.concatenate(L), sh


def test_chunked_cluster_fuzz(ctx):
    """random LCP arrays (dense events, long runs, 16-bit wrap) cut into chunks of one, two and a few scan tiles"""
    rng = np.random.default_rng(77)
    for it, n in enumerate([16384, 16385, 40000, 100003, 262144 + 7, 300000]):
        k = int(rng.choice([1, 2, 3, 16, 70]))
        m = int(rng.choice([1, 2, 3, 8, 33]))
        lcp = np.minimum(H.random_lcp(rng, n, k, it % 5), 127).astype(np.uint32)
        if it == 5:
            lcp[:] = 0
            lcp[1000:71000] = 40  # one 70 000-long cluster across five chunks: length wraps mod 2^16
            lcp[100000:100010] = 20
        bwt = rng.choice(H.BWT_ALPHABET, size=n)
        es, el, enc, _ = O.cluster_lm(lcp, bwt, k, m)
        for chunk in (16384, 32768, 5 * 16384):
            sm, s, l, sh = stream_phase1(ctx, lcp, bwt, k, m, chunk)
            S, L, mg = H.assemble([sm], [(s, l)])
            sh.close()
            assert mg.n_clust_out == enc and np.array_equal(S, es) and np.array_equal(L, el), (n, k, m, chunk)


def test_chunked_shards_merge(ctx):
    """two chunked shards of one eBWT (multi-GPU streaming): summaries merge like those of resident shards"""
    rng = np.random.default_rng(5)
    n, k, m = 150_000, 16, 2
    lcp = np.minimum(H.random_lcp(rng, n, k, 2), 127).astype(np.uint32)
    bwt = rng.choice(H.BWT_ALPHABET, size=n)
    es, el, enc, _ = O.cluster_lm(lcp, bwt, k, m)
    cut = 70_001
    sums, recs = [], []
    for lo, hi in ((0, cut), (cut, n)):
        sm, s, l, sh = stream_phase1(ctx, lcp, bwt, k, m, 32768, lo, hi)
        sums.append(sm)
        recs.append((s, l))
        sh.close()
    S, L, mg = H.assemble(sums, recs)
    assert mg.n_clust_out == enc and np.array_equal(S, es) and np.array_equal(L, el)


@pytest.mark.parametrize("name,seed", [("tiny", 4), ("small", 2)])
def test_chunked_pipeline_equals_oracle(ctx, name, seed, monkeypatch):
    """both tools through chunked shards: chunk = the whole range, a third of it, 100 003 positions (rounded up to scan
    tiles), one scan tile -- through the chunk API and through e2s_pipeline_host (E2S_CHUNK_POSITIONS)"""
    rs, e = H.dataset(name, seed)
    n = e["n"]
    k, m = 16, 2
    es, el, enc, _ = O.cluster_lm(e["lcp"], e["bwt"], k, m)
    off = O.uniform_read_offsets(*rs.reads.shape)
    p, op = api.default_params(rs.nreads1), O.default_params(rs.nreads1)
    ost = O.statistics(es, el, op.mcov_out, op.pval)
    otext, ores = O.find_events(e["lcp"], e["text"], e["suff"], e["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
    assert ores.n_candidates > 0
    want_clusters = O.clusters_to_bytes(es, el)
    ctx.stage_reads(rs.reads, off)
    for chunk in (n, n // 3, 100_003, 16384):
        sm, s, l, sh = stream_phase1(ctx, e["lcp"], e["bwt"], k, m, chunk, text=e["text"], suff=e["suff"], mcov=p.mcov_out)
        mg = api.cluster_merge([sm], 0)
        sh.cluster_finalize(mg)
        S, L, _ = H.assemble([sm], [(s, l)])
        assert np.array_equal(S, es) and np.array_equal(L, el), chunk
        st = sh.statistics(p.mcov_out, p.pval)
        assert st.max_clust_length == ost.max_clust_length and list(st.hist) == list(ost.hist)
        cnt = sh.find_events(p, st.max_clust_length)
        assert (cnt.n_analysed, cnt.n_candidates, cnt.n_events) == (ores.n_analysed, ores.n_candidates, ores.n_events), chunk
        assert api.events_format(sh.events(), p) == otext, chunk
        sh.close()
    # the one-call pipeline from host records
    rec = synth.gesa_records(e).view(np.uint8).reshape(-1)
    for chunk in (n, n // 3, 100_003):
        monkeypatch.setenv("E2S_CHUNK_POSITIONS", str(chunk))
        rec10 = np.empty((len(es) + 16) * 10, dtype=np.uint8)
        evbuf = (api.Event * (ores.n_candidates + 16))()
        res = ctx.pipeline_host(rec, n, rs.reads.reshape(-1), off, p, k, m, rec10=rec10, events=evbuf)
        assert res.n_written == len(es) and rec10[: len(es) * 10].tobytes() == want_clusters, chunk
        assert (res.n_clust_out, res.max_clust_length) == (enc, ost.max_clust_length)
        assert res.snp.n_candidates == ores.n_candidates
        assert api.events_format(list(evbuf)[: res.snp.n_variants], p) == otext, chunk


def test_chunked_lcp_above_127(ctx):
    """LCP values above 127 (saturated in the bit-sliced copy) stream through chunked shards for -k <= 127; -k > 127 on such a
    shard is refused by the chunk API (e2s_pipeline_host falls back to a resident shard)"""
    rs, e = H.dataset("tiny", 4)
    n = e["n"]
    lcp = e["lcp"].copy()
    rng = np.random.default_rng(9)
    at = rng.integers(0, n, size=n // 50)
    lcp[at] = rng.integers(128, 256, size=len(at))
    for k, m in ((16, 2), (127, 2)):
        es, el, enc, _ = O.cluster_lm(lcp, e["bwt"], k, m)
        for chunk in (n // 3, 16384):
            sm, s, l, sh = stream_phase1(ctx, lcp, e["bwt"], k, m, chunk)
            S, L, mg = H.assemble([sm], [(s, l)])
            sh.close()
            assert mg.n_clust_out == enc and np.array_equal(S, es) and np.array_equal(L, el), (k, chunk)
    sh = ctx.shard(n, 0, n, chunk_positions=65536)
    clo, cn = next(iter(sh.chunks()))
    sh.chunk_begin(clo, cn)
    hi = min(n, cn + 152)
    sh.load_soa(lcp[:hi], e["text"][:hi], e["suff"][:hi], e["bwt"][:hi], first=0)
    with pytest.raises(api.E2SError) as ei:
        sh.chunk_scan(128, 2, 5)
    assert ei.value.code == api.ERR_UNSUPPORTED
    sh.close()
    eg = dict(e)
    eg["lcp"] = lcp
    rec = synth.gesa_records(eg).view(np.uint8).reshape(-1)
    off = O.uniform_read_offsets(*rs.reads.shape)
    p, op = api.default_params(rs.nreads1), O.default_params(rs.nreads1)
    for k in (16, 128):
        es, el, enc, _ = O.cluster_lm(lcp, e["bwt"], k, 2)
        rec10 = np.empty((len(es) + 16) * 10, dtype=np.uint8)
        res = ctx.pipeline_host(rec, n, rs.reads.reshape(-1), off, p, k, 2, rec10=rec10)
        assert res.n_written == len(es) and rec10[: len(es) * 10].tobytes() == O.clusters_to_bytes(es, el), k
        ost = O.statistics(es, el, op.mcov_out, op.pval)
        _, ores = O.find_events(lcp, e["text"], e["suff"], e["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
        assert (res.snp.n_candidates, res.snp.n_events) == (ores.n_candidates, ores.n_events), k


@pytest.mark.parametrize("name,seed", [("tiny", 4), ("small", 2)])
@pytest.mark.parametrize("xyz", [(4, 4, 4), (1, 4, 1), (2, 8, 2)])
def test_lean_soa_pipeline_equals_oracle(ctx, name, seed, xyz, monkeypatch):
    """e2s_pipeline_host_soa: the BCR triple with only lcp + bwt sent to the device in full and the survivors' text / suff
    fetched from the host's pairSA -- same .clusters / .snp bytes as the oracle, whatever the chunk size and field widths"""
    x, y, z = xyz
    rs, e = H.dataset(name, seed)
    n = e["n"]
    k, m = 16, 2
    es, el, enc, _ = O.cluster_lm(e["lcp"], e["bwt"], k, m)
    off = O.uniform_read_offsets(*rs.reads.shape)
    p, op = api.default_params(rs.nreads1), O.default_params(rs.nreads1)
    ost = O.statistics(es, el, op.mcov_out, op.pval)
    otext, ores = O.find_events(e["lcp"], e["text"], e["suff"], e["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
    assert ores.n_candidates > 0
    want_clusters = O.clusters_to_bytes(es, el)
    lcp = np.ascontiguousarray(e["lcp"].astype(f"<u{x}")).view(np.uint8)
    pair = np.empty(n, dtype=np.dtype([("suff", f"<u{z}"), ("text", f"<u{y}")]))
    pair["suff"], pair["text"] = e["suff"], e["text"]
    pair = pair.view(np.uint8)
    bwt = np.ascontiguousarray(e["bwt"])
    for chunk in (n, n // 3, 100_003, 16384):
        monkeypatch.setenv("E2S_CHUNK_POSITIONS", str(chunk))
        rec10 = np.empty((len(es) + 16) * 10, dtype=np.uint8)
        evbuf = (api.Event * (ores.n_candidates + 16))()
        res = ctx.pipeline_host_soa(lcp, bwt, pair, n, rs.reads.reshape(-1), off, p, k, m, x, y, z, rec10=rec10, events=evbuf)
        assert res.n_written == len(es) and rec10[: len(es) * 10].tobytes() == want_clusters, chunk
        assert (res.n_clust_out, res.max_clust_length) == (enc, ost.max_clust_length)
        assert res.snp.n_candidates == ores.n_candidates
        assert api.events_format(list(evbuf)[: res.snp.n_variants], p) == otext, chunk
        if x == 1 and chunk == n:  # lcp + bwt in full, the survivors' records only: far from the 13 B/position of EGSA records
            assert res.h2d_bytes - rs.reads.size - 8 * len(off) < 6 * n, (res.h2d_bytes, n)


@pytest.mark.parametrize("xyz", [(4, 4, 4), (1, 4, 1), (2, 8, 2), (8, 2, 1)])
def test_load_bcr_equals_load_soa(ctx, xyz):
    """e2s_shard_load_bcr (the triple's bytes widened on the device) leaves the shard as e2s_shard_load_soa of the widened arrays
    does: both tools' outputs equal the oracle's, resident and with lcp + BWT only for ebwt2clust"""
    x, y, z = xyz
    rs, e = H.dataset("tiny", 4)
    n = e["n"]
    k, m = 16, 2
    text = e["text"] & (0xFFFF if y == 2 else 0xFFFFFFFF)  # (what a 2-byte field keeps)
    es, el, enc, _ = O.cluster_lm(e["lcp"], e["bwt"], k, m)
    off = O.uniform_read_offsets(*rs.reads.shape)
    p, op = api.default_params(rs.nreads1), O.default_params(rs.nreads1)
    ost = O.statistics(es, el, op.mcov_out, op.pval)
    otext, ores = O.find_events(e["lcp"], text, e["suff"], e["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
    lcp = np.ascontiguousarray(e["lcp"].astype(f"<u{x}")).view(np.uint8)
    pair = np.empty(n, dtype=np.dtype([("suff", f"<u{z}"), ("text", f"<u{y}")]))
    pair["suff"], pair["text"] = e["suff"], e["text"]
    pair = pair.view(np.uint8)
    ctx.stage_reads(rs.reads, off)
    for with_gsa in (True, False):
        sh = ctx.shard(n)
        # in two pieces that do not meet at a multiple of anything
        cut = n // 2 + 13
        for a, b in ((cut, n), (0, cut)):
            sh.load_bcr(lcp[a * x:b * x], e["bwt"][a:b], pair[a * (y + z):b * (y + z)] if with_gsa else None, x, y, z, first=a)
        sh.set_layout(x, y, z, True)
        sh.seal()
        nw, nc = sh.cluster_lm(k, m)
        assert nc == enc and sh.cluster_fetch_packed() == O.clusters_to_bytes(es, el)
        if with_gsa:
            st = sh.statistics(p.mcov_out, p.pval)
            cnt = sh.find_events(p, st.max_clust_length)
            assert cnt.n_candidates == ores.n_candidates and api.events_format(sh.events(), p) == otext
        sh.close()


@pytest.mark.parametrize("name,seed", [("tiny", 4), ("small", 2)])
@pytest.mark.parametrize("lean", [False, True])
def test_chunked_clust2snp_equals_oracle(ctx, name, seed, lean):
    """clust2snp alone on a chunked shard (an index larger than device memory): the .clusters records of each chunk are staged,
    the prefilter runs per chunk, phase 2 on the captured records -- same .snp as the oracle; lean = text / suff from the host"""
    rs, e = H.dataset(name, seed)
    n = e["n"]
    k, m = 16, 2
    es, el, enc, _ = O.cluster_lm(e["lcp"], e["bwt"], k, m)
    off = O.uniform_read_offsets(*rs.reads.shape)
    p, op = api.default_params(rs.nreads1), O.default_params(rs.nreads1)
    ost = O.statistics(es, el, op.mcov_out, op.pval)
    otext, ores = O.find_events(e["lcp"], e["text"], e["suff"], e["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
    rec10 = np.frombuffer(O.clusters_to_bytes(es, el), dtype=np.uint8)
    ctx.stage_reads(rs.reads, off)
    pair = np.empty(n, dtype=np.dtype([("suff", "<u1"), ("text", "<u4")]))
    pair["suff"], pair["text"] = e["suff"], e["text"]
    pair = pair.view(np.uint8)
    lcp1 = e["lcp"].astype(np.uint8)
    for chunk in (n, n // 3, 100_003, 16384):
        sh = ctx.shard(n, 0, n, chunk_positions=chunk)
        if lean:
            sh.host_gsa(pair, 4, 1)
        for clo, cn in sh.chunks():
            sh.chunk_begin(clo, cn)
            a, b = max(0, clo - 176), min(n, clo + cn + 152)
            if lean:
                sh.load_bcr(lcp1[a:b], e["bwt"][a:b], None, 1, 4, 1, first=a)
                sh.set_layout(1, 4, 1, True)
            else:
                sh.load_soa(e["lcp"][a:b], e["text"][a:b], e["suff"][a:b], e["bwt"][a:b], first=a)
            r0, r1 = np.searchsorted(es, clo), np.searchsorted(es, clo + cn)
            sh.chunk_stage_clusters(rec10[r0 * 10:r1 * 10], p.mcov_out, ost.max_clust_length)
        sh.chunked_clusters_finish()
        cnt = sh.find_events(p, ost.max_clust_length)
        assert (cnt.n_analysed, cnt.n_candidates, cnt.n_events) == (ores.n_analysed, ores.n_candidates, ores.n_events), chunk
        assert api.events_format(sh.events(), p) == otext, chunk
        sh.close()


def test_reads_150bp_stream_like_any_other(ctx, monkeypatch):
    """150-base reads (LCP values up to 150: saturated in the bit-sliced copy) through the streaming pipelines and the lean
    SoA pipeline: .clusters / .snp equal the oracle's, and the index is the library builder's"""
    rs = synth.make_read_set(G=40_000, reads_per_sample=5_000, L=150, n_snps=80, n_indels=8, rc=True, seed=150)
    e = synth.build_egsa(rs.reads)
    eg = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in e.items()}
    for f in ("lcp", "text", "suff"):
        eg[f] = eg[f].view(np.uint32)
    n = int(eg["n"])
    assert int(eg["lcp"].max()) > 127
    mine = ctx.build_egsa(rs.reads)
    for key in ("lcp", "text", "suff", "bwt"):
        got = mine[key].cpu().numpy()
        assert np.array_equal(got.view(np.uint32) if key != "bwt" else got, eg[key]), key
    k, m = 16, 2
    es, el, enc, _ = O.cluster_lm(eg["lcp"], eg["bwt"], k, m)
    off = O.uniform_read_offsets(*rs.reads.shape)
    p, op = api.default_params(rs.nreads1), O.default_params(rs.nreads1)
    ost = O.statistics(es, el, op.mcov_out, op.pval)
    otext, ores = O.find_events(eg["lcp"], eg["text"], eg["suff"], eg["bwt"], es, el, op, ost.max_clust_length, rs.reads, off)
    assert ores.n_events > 0
    want = O.clusters_to_bytes(es, el)
    rec = synth.gesa_records(eg).view(np.uint8).reshape(-1)
    lcp1 = eg["lcp"].astype(np.uint8)
    pair = np.empty(n, dtype=np.dtype([("suff", "<u1"), ("text", "<u4")]))
    pair["suff"], pair["text"] = eg["suff"], eg["text"]
    pair = pair.view(np.uint8)
    for chunk in (n, 100_003):
        monkeypatch.setenv("E2S_CHUNK_POSITIONS", str(chunk))
        for lean in (False, True):
            rec10 = np.empty((len(es) + 16) * 10, dtype=np.uint8)
            evbuf = (api.Event * (ores.n_candidates + 16))()
            if lean:
                res = ctx.pipeline_host_soa(lcp1, eg["bwt"], pair, n, rs.reads.reshape(-1), off, p, k, m, 1, 4, 1, rec10=rec10, events=evbuf)
            else:
                res = ctx.pipeline_host(rec, n, rs.reads.reshape(-1), off, p, k, m, rec10=rec10, events=evbuf)
            assert res.n_written == len(es) and rec10[: len(es) * 10].tobytes() == want, (chunk, lean)
            assert (res.n_clust_out, res.max_clust_length, res.snp.n_candidates) == (enc, ost.max_clust_length, ores.n_candidates)
            assert api.events_format(list(evbuf)[: res.snp.n_variants], p) == otext, (chunk, lean)
