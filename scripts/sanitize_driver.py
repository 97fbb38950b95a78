"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_driver.py
Checks results against the oracle as well, so a sanitizer-clean run is also a correct run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ebwt2snp_b200 import api, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

ctx = api.Context(0)
rng = np.random.default_rng(9)
# phase 1: dense / sparse / long-run masks, bit-parallel and EXACT paths, multi-window tiles, shards
for it, n in enumerate([5, 300, 8193, 40000, 70001]):
    for m in (2, 40):
        k = int(rng.choice([2, 16]))
        lcp = H.random_lcp(rng, n, k, it % 5)
        bwt = rng.choice(H.BWT_ALPHABET, size=n)
        es, el, enc, _ = O.cluster_lm(lcp, bwt, k, m)
        sh = ctx.shard(n)
        sh.load_soa(lcp, None, None, bwt)
        sh.seal()
        nw, nc = sh.cluster_lm(k, m)
        s, l = sh.cluster_fetch()
        assert nc == enc and np.array_equal(s, es) and np.array_equal(l, el), (n, k, m)
        sh.close()
n = 140000
lcp = np.zeros(n, dtype=np.uint32)
lcp[100:100 + 70000] = 40
bwt = np.full(n, ord("A"), dtype=np.uint8)
es, el, enc, _ = O.cluster_lm(lcp, bwt, 16, 2)
for lo, hi in ((0, n), (0, 50000), (50000, n)):
    sh = ctx.shard(hi - lo, lo, n)
    a, b = max(0, lo - 2), min(n, hi + 151)
    sh.load_soa(lcp[a:b], None, None, bwt[a:b], first=a)
    sh.seal()
    sh.cluster_run(16, 2)
    sh.close()
# both phases on a tiny read set, .gesa staging path (unpack kernel) included
rs = synth.make_config("tiny", seed=5)
e = synth.build_egsa(rs.reads)
eg = {kk: (v.numpy() if hasattr(v, "numpy") else v) for kk, v in e.items()}
for f in ("lcp", "text", "suff"):
    eg[f] = eg[f].view(np.uint32)
n = int(eg["n"])
sh = ctx.shard(n)
sh.load_gesa(synth.gesa_records(eg).view(np.uint8).reshape(-1), 0, n)
sh.seal()
sh.cluster_lm(16, 2)
es, el, _, _ = O.cluster_lm(eg["lcp"], eg["bwt"], 16, 2)
assert sh.cluster_fetch_packed() == O.clusters_to_bytes(es, el)
off = O.uniform_read_offsets(*rs.reads.shape)
ctx.stage_reads(rs.reads, off)
p = api.default_params(rs.nreads1)
st = sh.statistics(p.mcov_out, p.pval)
cnt = sh.find_events(p, st.max_clust_length)
op = O.default_params(rs.nreads1)
otext, ores = O.find_events(eg["lcp"], eg["text"], eg["suff"], eg["bwt"], es, el, op, O.statistics(es, el, 5, 0.99).max_clust_length, rs.reads, off)
assert api.events_format(sh.events(), p) == otext and cnt.n_candidates == ores.n_candidates
sh.close()
ctx.close()
print("sanitize_driver: all results match the oracle")
