"""not gpu: bench.py's sampled property check of an index (what vouches for the C3-size index nobody else can build):
passes on a correct index, fails on a wrong LCP value, on two swapped records and on a wrong BWT byte."""
import importlib.util
import os

import numpy as np

from ebwt2snp_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_index_property_check_detects_damage():
    b = _bench()
    rs = synth.make_config("tiny", seed=9)
    eg = synth.build_egsa(rs.reads)
    ok, n_checked = b.check_egsa_sample(rs, eg, samples=30000)
    assert ok and n_checked > 30000
    n = int(eg["n"])
    every = dict(samples=4 * n)  # (nearly) every record gets sampled

    bad = dict(eg)
    lcp = eg["lcp"].clone()
    lcp[n // 2] += 1
    bad["lcp"] = lcp
    assert not b.check_egsa_sample(rs, bad, **every)[0]

    bad = dict(eg)
    t, s = eg["text"].clone(), eg["suff"].clone()
    i = n // 3
    t[[i, i + 1]] = t[[i + 1, i]]
    s[[i, i + 1]] = s[[i + 1, i]]
    bad["text"], bad["suff"] = t, s
    assert not b.check_egsa_sample(rs, bad, **every)[0]

    bad = dict(eg)
    bw = eg["bwt"].clone()
    j = int(np.flatnonzero(eg["suff"].numpy() > 0)[1234])
    bw[j] = ord("A") if bw[j] != ord("A") else ord("C")
    bad["bwt"] = bw
    assert not b.check_egsa_sample(rs, bad, **every)[0]
