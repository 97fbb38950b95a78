"""-m gpu: the index builder (csrc/build_egsa.cu: the library's own radix sort, reads of any lengths, 32- and 64-bit
suffix ids) against the oracle's comparison sort (oracle_build_egsa / oracle_build_egsa_ragged) and, through the files,
against the UNMODIFIED reference tools run on the index it wrote."""
import os
import subprocess

import numpy as np
import pytest

from ebwt2snp_b200 import api, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ebwt2snp_b200", "bin")
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def ctx(built):
    c = api.Context(0)
    yield c
    c.close()


def ragged_from_genome(rng, G, R, lo, hi, genome=None):
    """R reads with lengths in [lo, hi] sampled from one random genome (so suffixes share long prefixes and tie)"""
    g = BASES[rng.integers(0, 4, size=G)] if genome is None else genome
    lens = rng.integers(lo, hi + 1, size=R)
    st = rng.integers(0, len(g) - hi + 1, size=R)
    off = np.zeros(R + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    bases = np.empty(int(off[R]), dtype=np.uint8)
    for r in range(R):
        bases[int(off[r]):int(off[r + 1])] = g[st[r]:st[r] + lens[r]]
    return bases, off


def check_ragged(ctx, bases, off, what):
    want = O.build_egsa_ragged(bases, off)
    got = ctx.build_egsa_ragged(bases, off)
    assert got["n"] == want["n"]
    for key in ("text", "suff", "lcp", "bwt"):
        g = got[key].cpu().numpy()
        g = g.view(np.uint32) if key != "bwt" else g
        bad = np.flatnonzero(g != want[key])
        assert bad.size == 0, (what, key, int(bad[0]), g[bad[:4]], want[key][bad[:4]])


CASES = ["tiny", "genome", "two_length_places", "skewed", "many_tiles", "skewed_many_tiles"]


def make_case(name):
    rng = np.random.default_rng(CASES.index(name) + 7)
    if name == "tiny":  # empty reads, equal reads, a read that is a prefix of another
        reads = [b"", b"ACGT", b"ACG", b"", b"ACGT", b"T", b"GATTACA", b"ACGTT", b"A", b""]
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in reads])
        return np.frombuffer(b"".join(reads), dtype=np.uint8).copy(), off
    if name == "genome":  # 30x of a 5 kbp genome, lengths 0..150 (five key words), a few tiles
        return ragged_from_genome(rng, 5_000, 2_000, 0, 150)
    if name == "two_length_places":  # a read of 300 and one of 700 bases: the length key needs two digit places
        b, off = ragged_from_genome(rng, 3_000, 60, 1, 90)
        g = BASES[rng.integers(0, 4, size=1000)]
        extra = [g[:300], g[100:800], g[:300]]
        lens = np.concatenate([np.diff(off.astype(np.int64)), [len(x) for x in extra]])
        off2 = np.zeros(len(lens) + 1, dtype=np.uint64)
        off2[1:] = np.cumsum(lens)
        return np.concatenate([b] + extra), off2
    if name == "skewed":  # one digit takes nearly everything: poly-A reads, copies of one read, copies of a short read
        parts = [np.full(int(l), ord("A"), dtype=np.uint8) for l in rng.integers(0, 120, size=600)]
        one = BASES[rng.integers(0, 4, size=97)]
        parts += [one] * 300 + [np.frombuffer(b"ACGTACGTACGTTTTT", dtype=np.uint8)] * 50
        lens = [len(p) for p in parts]
        off = np.zeros(len(parts) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        return np.concatenate(parts), off
    if name == "many_tiles":  # 1.1 M suffixes = 270 tiles: the chained scan across many tiles
        return ragged_from_genome(rng, 200_000, 11_000, 60, 140)
    if name == "skewed_many_tiles":  # 170 tiles in which one or two digits take everything: poly-A / poly-T reads and copies
        parts = [np.full(int(l), ord("A"), dtype=np.uint8) for l in rng.integers(0, 150, size=5000)]
        parts += [np.full(int(l), ord("T"), dtype=np.uint8) for l in rng.integers(0, 150, size=2000)]
        one = BASES[rng.integers(0, 4, size=120)]
        parts += [one[: int(l)] for l in rng.integers(100, 121, size=1500)]
        off = np.zeros(len(parts) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(p) for p in parts])
        return np.concatenate(parts), off
    raise KeyError(name)


@pytest.mark.parametrize("threads", [None, "256", "512"])
@pytest.mark.parametrize("ids64", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_ragged_builder_vs_oracle(ctx, monkeypatch, name, ids64, threads):
    """both id widths x both shapes of the radix pass (256 threads x 16 pairs, 512 x 8; None = the default of the id width)"""
    if ids64:
        monkeypatch.setenv("E2S_BUILD_IDS64", "1")
    else:
        monkeypatch.delenv("E2S_BUILD_IDS64", raising=False)
    if threads:
        monkeypatch.setenv("E2S_RADIX_THREADS", threads)
    else:
        monkeypatch.delenv("E2S_RADIX_THREADS", raising=False)
    bases, off = make_case(name)
    check_ragged(ctx, bases, off, name)


@pytest.mark.parametrize("ids64", [False, True])
@pytest.mark.parametrize("R,L", [(1, 1), (3, 32), (700, 33), (2_000, 64), (1_500, 100), (900, 150), (40, 257),
                                 (64, 63), (65, 63), (128, 63), (32, 127), (1024, 31), (1023, 31)])  # n = whole tiles of 4096, one more, one less
def test_equal_length_builder_vs_oracle(ctx, monkeypatch, R, L, ids64):
    """equal lengths take the arithmetic ids (no length pass): every word-count boundary, both id widths; the same reads
    through the ragged entry point give the same index"""
    if ids64:
        monkeypatch.setenv("E2S_BUILD_IDS64", "1")
    else:
        monkeypatch.delenv("E2S_BUILD_IDS64", raising=False)
    rng = np.random.default_rng(R * 1000 + L)
    g = BASES[rng.integers(0, 4, size=max(2 * L, R // 4 + L))]
    st = rng.integers(0, len(g) - L + 1, size=R)
    reads = g[st[:, None] + np.arange(L)[None, :]]
    want = O.build_egsa(reads)
    got = ctx.build_egsa(reads)
    off = (np.arange(R + 1, dtype=np.uint64) * np.uint64(L))
    rag = ctx.build_egsa_ragged(reads.reshape(-1), off)
    for key in ("text", "suff", "lcp", "bwt"):
        for res, tag in ((got, "equal"), (rag, "ragged")):
            a = res[key].cpu().numpy()
            a = a.view(np.uint32) if key != "bwt" else a
            assert np.array_equal(a, want[key]), (tag, key)


def test_builder_refuses_what_it_cannot_key(ctx):
    with pytest.raises(Exception) as ei:
        ctx.build_egsa(np.frombuffer(b"ACGTNACG", dtype=np.uint8).reshape(2, 4))
    assert "ACGT" in str(ei.value)
    with pytest.raises(Exception):  # nothing but empty reads
        ctx.build_egsa_ragged(np.zeros(0, dtype=np.uint8), np.array([0, 0, 0], dtype=np.uint64))
    off = np.array([0, 70_000], dtype=np.uint64)
    with pytest.raises(Exception) as ei:
        ctx.build_egsa_ragged(np.full(70_000, ord("A"), dtype=np.uint8), off)
    assert "65536" in str(ei.value)


def run(tool, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([os.path.join(BIN, tool), *[str(a) for a in args]], capture_output=True, text=True, timeout=600, env=e)


def test_ragged_reads_through_the_tool_chain(built, tmp_path):
    """trimmed reads (70..100 bases) of two samples with planted SNPs: FASTA -> build_gesa -> ebwt2clust -> clust2snp; the
    index equals the oracle builder's, and the UNMODIFIED reference tools (oracle/_ref, when built) give the same
    .clusters / .snp from the index build_gesa wrote"""
    rs = synth.make_read_set(G=30_000, reads_per_sample=6_000, L=100, n_snps=60, n_indels=0, rc=True, seed=31)
    rng = np.random.default_rng(5)
    lens = rng.integers(70, 101, size=rs.reads.shape[0])
    fa = str(tmp_path / "ALL.fasta")
    with open(fa, "wb") as f:
        for r in range(rs.reads.shape[0]):
            f.write(b">r%d\n" % r + rs.reads[r, :lens[r]].tobytes() + b"\n")
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    bases = np.concatenate([rs.reads[r, :lens[r]] for r in range(len(lens))])
    want = O.build_egsa_ragged(bases, off)
    w = ["-x", 1, "-y", 4, "-z", 1]
    r = run("build_gesa", "-i", fa, *w)
    assert r.returncode == 0, r.stdout + r.stderr
    rec = np.fromfile(fa + ".gesa", dtype=np.dtype([("text", "<u4"), ("suff", "u1"), ("lcp", "u1"), ("bwt", "u1")]))
    assert len(rec) == want["n"]
    for key in ("text", "suff", "lcp", "bwt"):
        assert np.array_equal(rec[key].astype(np.uint32), want[key].astype(np.uint32)), key
    r = run("ebwt2clust", "-i", fa, *w)
    assert r.returncode == 0, r.stdout + r.stderr
    mine_cl = open(fa + ".clusters", "rb").read()
    r = run("clust2snp", "-i", fa, "-n", rs.nreads1, *w)
    assert r.returncode == 0, r.stdout + r.stderr
    snp = str(tmp_path / "ALL.snp")
    mine_snp = open(snp, "rb").read()
    assert len(mine_cl) > 0 and len(mine_snp) > 0
    if O.ref_available():
        os.remove(fa + ".clusters")
        os.remove(snp)
        ref = os.path.join(ROOT, "oracle", "_ref")
        r1 = subprocess.run([os.path.join(ref, "ebwt2clust"), "-i", fa, "-x", "1", "-y", "4", "-z", "1"], capture_output=True, text=True, timeout=600)
        assert r1.returncode == 0
        assert open(fa + ".clusters", "rb").read() == mine_cl
        r2 = subprocess.run([os.path.join(ref, "clust2snp"), "-i", fa, "-n", str(rs.nreads1), "-x", "1", "-y", "4", "-z", "1"],
                            capture_output=True, text=True, timeout=600)
        assert r2.returncode == 0
        assert open(snp, "rb").read() == mine_snp


def key_of(prefix):
    """the 64-bit first key word of a suffix that starts with `prefix` (bytes of ACGT), zero padded"""
    v = 0
    for ch in prefix:
        v = (v << 2) | b"ACGT".index(ch)
    return v << (64 - 2 * len(prefix))


@pytest.mark.parametrize("ids64", [False, True])
def test_key_ranges_concatenate_to_the_index(ctx, monkeypatch, ids64):
    """e2s_build_egsa_range_dev: the index built range after range (each range told the last record of the one before) equals
    the index built in one call; ranges built independently (predecessor not known) differ only in lcp[0]; first_position is the
    range's place in the index; an empty range and a capacity that is too small answer as documented"""
    if ids64:
        monkeypatch.setenv("E2S_BUILD_IDS64", "1")
    else:
        monkeypatch.delenv("E2S_BUILD_IDS64", raising=False)
    rs = synth.make_read_set(G=20_000, reads_per_sample=3_000, L=100, n_snps=40, n_indels=4, rc=True, seed=77)
    whole = O.build_egsa(rs.reads)
    n = whole["n"]
    cuts = [0, key_of(b"AC"), key_of(b"CGT"), key_of(b"CGTC"), key_of(b"G"), key_of(b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"), 0]
    at, before = 0, None
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = ctx.build_egsa_range(rs.reads, lo, hi, before=before)
        alone = ctx.build_egsa_range(rs.reads, lo, hi)
        m = part["n"]
        assert part["first"] == at and alone["first"] == at and alone["n"] == m, (hex(lo), hex(hi))
        for key in ("text", "suff", "lcp", "bwt"):
            got = part[key].cpu().numpy()
            got = got.view(np.uint32) if key != "bwt" else got
            assert np.array_equal(got, whole[key][at:at + m]), (key, hex(lo), hex(hi))
            solo = alone[key].cpu().numpy()
            solo = solo.view(np.uint32) if key != "bwt" else solo
            if key == "lcp" and m:
                assert solo[0] == 0 and np.array_equal(solo[1:], whole[key][at + 1:at + m])
            else:
                assert np.array_equal(solo, whole[key][at:at + m])
        if m:
            before = (int(whole["text"][at + m - 1]), int(whole["suff"][at + m - 1]))
        at += m
    assert at == n
    empty = ctx.build_egsa_range(rs.reads, key_of(b"CGTC") + 1, key_of(b"CGTC") + 2)
    assert empty["n"] == 0
    with pytest.raises(Exception) as ei:
        ctx.build_egsa_range(rs.reads, 0, 0, capacity=n - 1)
    assert "capacity" in str(ei.value)


def test_build_gesa_by_key_ranges_when_the_index_does_not_fit(built, tmp_path):
    """e2s_build_egsa (host in / out, what build_gesa calls) builds the index key range by key range when one call's device memory
    would not hold it; E2S_BUILD_RANGE_RECORDS forces that path with tiny ranges (some are cut in two on the way): same files"""
    rs = synth.make_config("tiny", seed=21)
    outs = []
    for j, env in enumerate([None, {"E2S_BUILD_RANGE_RECORDS": "30000"}, {"E2S_BUILD_RANGE_RECORDS": "25000", "E2S_BUILD_IDS64": "1"}]):
        d = tmp_path / f"v{j}"
        os.makedirs(d)
        fa = str(d / "ALL.fasta")
        synth.write_fasta(fa, rs.reads)
        r = run("build_gesa", "-i", fa, "-x", 1, "-y", 4, "-z", 1, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(open(fa + ".gesa", "rb").read())
    assert len(outs[0]) == rs.reads.shape[0] * (rs.reads.shape[1] + 1) * 7
    assert outs[1] == outs[0] and outs[2] == outs[0]
