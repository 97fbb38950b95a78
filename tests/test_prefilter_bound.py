"""not gpu: the two exactness arguments of the BWT-only prefilter (csrc/planes.cuh), as properties on random
clusters: (1) fewer than mcov positions differing from the cluster's FIRST base code => at most one base code reaches mcov,
(2) at most one code reaching mcov over both samples => find_variants' filters (ref:clust2snp.cpp:402-429) reject the cluster
whatever the sample assignment."""
import numpy as np


def test_bound_and_prefilter_never_reject_a_passing_cluster():
    rng = np.random.default_rng(8)
    for _ in range(20000):
        n = int(rng.integers(2, 151))
        mcov = int(rng.choice([1, 2, 3, 5, 8]))
        skew = rng.dirichlet(np.ones(4) * rng.choice([0.05, 0.3, 1.0]))
        codes = rng.choice(4, size=n, p=skew)
        sample = rng.integers(0, 2, size=n)
        total = np.bincount(codes, minlength=4)
        others = int((codes != codes[0]).sum())
        n_frequent = int((total >= mcov).sum())
        if others < mcov:  # the one-popcount bound
            assert n_frequent <= 1
        counts = np.array([[((codes == c) & (sample == s)).sum() for c in range(4)] for s in range(2)])
        f = [set(np.flatnonzero(counts[s] >= mcov).tolist()) for s in range(2)]
        passes = bool(f[0]) and bool(f[1]) and len(f[0]) <= 2 and len(f[1]) <= 2 and f[0] != f[1] and len(f[0] | f[1]) <= 3
        if n_frequent <= 1:  # what the prefilter drops
            assert not passes
