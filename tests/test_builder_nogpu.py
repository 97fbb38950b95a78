"""The checker of the ragged index builder, checked: oracle_build_egsa_ragged against a plain Python sort, against the
equal-length oracle builder, and (with oracle/_ref built) the whole reference tool chain on ragged reads against the oracle
port -- the reference's FASTA parser takes reads of any lengths (ref:clust2snp.cpp:147-212)."""
import os
import subprocess

import numpy as np
import pytest

from ebwt2snp_b200 import synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def naive(reads):
    suf = [(r[p:], i, p) for i, r in enumerate(reads) for p in range(len(r) + 1)]
    suf.sort(key=lambda t: (t[0], t[1]))  # bytes order: a proper prefix first (= `$` smallest), then the read id
    lcp = [0]
    for (a, _, _), (b, _, _) in zip(suf, suf[1:]):
        l = 0
        while l < min(len(a), len(b)) and a[l] == b[l]:
            l += 1
        lcp.append(l)
    text = [t[1] for t in suf]
    suff = [t[2] for t in suf]
    bwt = [reads[t[1]][t[2] - 1] if t[2] else ord("$") for t in suf]
    return lcp, text, suff, bwt


def as_arrays(reads):
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    return np.frombuffer(b"".join(reads), dtype=np.uint8).copy(), off


def test_ragged_oracle_builder_vs_plain_python():
    rng = np.random.default_rng(3)
    reads = [b"", b"ACGT", b"ACG", b"", b"ACGT", b"T", b"GATTACA", b"ACGTT", b"A", b""]
    reads += [bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 2, size=int(l))]) for l in rng.integers(0, 40, size=200)]
    bases, off = as_arrays(reads)
    e = O.build_egsa_ragged(bases, off)
    lcp, text, suff, bwt = naive(reads)
    assert e["n"] == len(lcp)
    assert e["lcp"].tolist() == lcp and e["text"].tolist() == text and e["suff"].tolist() == suff and e["bwt"].tolist() == bwt


def test_ragged_oracle_builder_equals_equal_length_builder():
    rs = synth.make_config("tiny", seed=4)
    reads = rs.reads[:400]
    R, L = reads.shape
    a = O.build_egsa(reads)
    b = O.build_egsa_ragged(reads.reshape(-1), np.arange(R + 1, dtype=np.uint64) * np.uint64(L))
    for k in ("lcp", "text", "suff", "bwt"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
def test_reference_tools_on_ragged_reads_equal_the_oracle_port(tmp_path):
    rs = synth.make_read_set(G=8_000, reads_per_sample=1_600, L=100, n_snps=20, n_indels=0, rc=True, seed=31)
    rng = np.random.default_rng(5)
    lens = rng.integers(70, 101, size=rs.reads.shape[0])
    reads = [rs.reads[r, :lens[r]].tobytes() for r in range(len(lens))]
    fa = str(tmp_path / "ALL.fasta")
    with open(fa, "wb") as f:
        for r, s in enumerate(reads):
            f.write(b">r%d\n" % r + s + b"\n")
    bases, off = as_arrays(reads)
    e = O.build_egsa_ragged(bases, off)
    rec = np.empty(e["n"], dtype=np.dtype([("text", "<u4"), ("suff", "u1"), ("lcp", "u1"), ("bwt", "u1")]))
    for k in ("text", "suff", "lcp", "bwt"):
        rec[k] = e[k]
    rec.tofile(fa + ".gesa")
    ref = os.path.join(ROOT, "oracle", "_ref")
    w = ["-x", "1", "-y", "4", "-z", "1"]
    assert subprocess.run([os.path.join(ref, "ebwt2clust"), "-i", fa, *w], capture_output=True, timeout=300).returncode == 0
    assert subprocess.run([os.path.join(ref, "clust2snp"), "-i", fa, "-n", str(rs.nreads1), *w], capture_output=True, timeout=300).returncode == 0
    es, el, _, _ = O.cluster_lm(e["lcp"], e["bwt"], 16, 2)
    assert O.clusters_to_bytes(es, el) == open(fa + ".clusters", "rb").read()
    op = O.default_params(rs.nreads1)
    ost = O.statistics(es, el, op.mcov_out, op.pval)
    otext, ores = O.find_events(e["lcp"], e["text"], e["suff"], e["bwt"], es, el, op, ost.max_clust_length, bases, off)
    snp = open(str(tmp_path / "ALL.snp"), "rb").read()
    assert ores.n_events > 0 and (otext if isinstance(otext, bytes) else otext.encode()) == snp


def test_key_range_cuts_tile_the_key_space_and_balance_the_records():
    """sharding.key_range_cuts: the ranges e2s_build_egsa_range_dev takes, one per rank.  They tile [0, 2^64); with the reads given
    they are sample quantiles, and the oracle's index splits into nearly equal record counts by them (first_key_words = the key
    the library compares, checked against the index order: keys are non-decreasing along the index)"""
    from ebwt2snp_b200 import sharding
    for parts in (1, 2, 3, 8):
        cuts = sharding.key_range_cuts(parts)
        assert cuts[0][0] == 0 and cuts[-1][1] == 0 and len(cuts) == parts
        assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(cuts, cuts[1:]))
    rs = synth.make_read_set(G=30_000, reads_per_sample=2_000, L=100, n_snps=30, n_indels=2, rc=True, seed=9)
    e = O.build_egsa(rs.reads)
    keys = sharding.first_key_words(rs.reads, e["text"], e["suff"])
    assert np.all(keys[1:] >= keys[:-1])  # the first key word is the most significant part of the index order
    n = e["n"]
    for parts in (2, 4, 8):
        cuts = sharding.key_range_cuts(parts, rs.reads, sample=1 << 14, seed=1)
        assert cuts[0][0] == 0 and cuts[-1][1] == 0
        assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(cuts, cuts[1:]))
        counts = [int(np.count_nonzero((keys >= np.uint64(lo)) & ((keys < np.uint64(hi)) if hi else True))) for lo, hi in cuts]
        assert sum(counts) == n
        assert max(counts) < 1.25 * n / parts and min(counts) > 0.75 * n / parts, counts
