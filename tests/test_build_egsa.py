"""EGSA construction: the oracle's comparison sort against the torch builder the synthetic data
has always come from (CPU), and the CUDA builder (e2s_build_egsa_dev) against both (-m gpu)."""
import numpy as np
import pytest

from ebwt2snp_b200 import synth
from oracle import oracle as O

CASES = [  # G, reads per sample, L, SNPs, indels, RC, seed
    (300, 40, 20, 3, 0, False, 1),
    (500, 60, 33, 4, 1, True, 2),      # L = 33: two key words, the second with one symbol
    (2000, 300, 64, 6, 2, True, 3),    # L = 64: exactly two full key words
    (3000, 250, 100, 8, 2, True, 4),
    (50, 200, 31, 0, 0, False, 5),     # tiny genome: many identical reads (ties by read id)
]


def _reads(case):
    G, rps, L, ns, ni, rc, seed = case
    return synth.make_read_set(G, rps, L, ns, ni, rc, seed).reads


def _same(a, b):
    for k in ("lcp", "text", "suff", "bwt"):
        x = a[k].cpu().numpy() if hasattr(a[k], "cpu") else a[k]
        y = b[k].cpu().numpy() if hasattr(b[k], "cpu") else b[k]
        assert np.array_equal(x.view(np.uint32) if x.dtype != np.uint8 else x, y.view(np.uint32) if y.dtype != np.uint8 else y), k


@pytest.mark.parametrize("case", CASES)
def test_oracle_builder_matches_torch_builder(case):
    reads = _reads(case)
    _same(O.build_egsa(reads), synth.build_egsa(reads, device="cpu"))


def test_oracle_builder_definition():
    """hand-checked: reads AC, AA -> suffixes $0 $1 A$(r1,1) AA$(r1,0) AC$(r0,0) C$(r0,1)"""
    reads = np.frombuffer(b"ACAA", dtype=np.uint8).reshape(2, 2)
    e = O.build_egsa(reads)
    assert e["text"].tolist() == [0, 1, 1, 1, 0, 0]
    assert e["suff"].tolist() == [2, 2, 1, 0, 0, 1]
    assert e["lcp"].tolist() == [0, 0, 0, 1, 1, 0]
    assert bytes(e["bwt"]) == b"CAA$$A"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_builder_matches_oracle(built, case):
    from ebwt2snp_b200 import api
    reads = _reads(case)
    ctx = api.Context(0)
    try:
        _same(ctx.build_egsa(reads), O.build_egsa(reads))
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_builder_rejects_bases_outside_acgt(built):
    """an N (or any byte that is not ACGT / acgt) has no 2-bit code: the builder must refuse it, not index it as A"""
    from ebwt2snp_b200 import api
    reads = np.frombuffer(b"ACGTACGTACGTACGNACGTACGTacgtACGT", dtype=np.uint8).reshape(4, 8)
    ctx = api.Context(0)
    try:
        with pytest.raises(api.E2SError) as ei:
            ctx.build_egsa(reads)
        assert ei.value.code == api.ERR_UNSUPPORTED and "ACGT" in str(ei.value)
        ok = reads.copy()
        ok[1, 7] = ord("a")  # lower case is a base
        assert ctx.build_egsa(ok)["n"] == 4 * 9
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_builder_matches_torch_builder_small(built):
    """the 'small' configuration (8.1 M suffixes): CUDA builder == the torch sorts on the same GPU, and the hot path on
    arrays born on the device gives the oracle's clusters"""
    import torch
    from ebwt2snp_b200 import api
    rs = synth.make_config("small", seed=3)
    ctx = api.Context(0)
    try:
        mine = ctx.build_egsa(rs.reads)
        ref = synth.build_egsa(rs.reads, device="cuda:0")
        _same(mine, ref)
        n = mine["n"]
        sh = ctx.shard(n)
        sh.load_soa(mine["lcp"], mine["text"], mine["suff"], mine["bwt"], device=True)
        sh.seal()
        nw, nc = sh.cluster_lm(16, 2)
        s, l = sh.cluster_fetch()
        es, el, enc, _ = O.cluster_lm(ref["lcp"].cpu().numpy().view(np.uint32), ref["bwt"].cpu().numpy(), 16, 2)
        assert nc == enc and np.array_equal(s, es) and np.array_equal(l, el)
        sh.close()
        torch.cuda.synchronize()
    finally:
        ctx.close()


@pytest.mark.parametrize("case", CASES[:3] + [CASES[4]])
def test_two_bit_zero_padded_keys_with_shortest_first_ties_give_the_terminated_order(case):
    """the ordering argument of csrc/build_egsa.cu, restated in plain Python (no GPU): sort the suffixes by their 2-bit
    packed, ZERO-padded keys with a STABLE sort whose initial order is (offset descending, read id ascending) -> exactly
    the order of the `$`-terminated strings (`$` < A < C < G < T, equal suffixes by read id) the oracle's comparison sort gives"""
    reads = _reads(case)
    R, L = reads.shape
    code = np.zeros(256, dtype=np.int64)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    c = code[reads]
    ids = [(L - p) * R + r for p in range(L, -1, -1) for r in range(R)]  # ascending ids = the initial order
    assert ids == sorted(ids)

    def key(i):
        p, r = L - i // R, i % R
        v = 0
        for s in range(p, L):
            v = (v << 2) | int(c[r, s])
        return v << (2 * p)  # zero padding up to L symbols

    order = sorted(ids, key=key)  # Python's sort is stable
    e = O.build_egsa(reads)
    assert [i % R for i in order] == e["text"].tolist()
    assert [L - i // R for i in order] == e["suff"].tolist()
