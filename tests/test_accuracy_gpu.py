"""-m gpu: semantic sanity beside bit parity (SURVEY.md 8(f) rank 4 in miniature): the generator knows where it planted
the SNPs, so the calls of the CUDA path can be scored the way snp_vs_vcf scores them -- every SNP event's 61-mer must
sit on a planted position of sample 1's genome (precision) and nearly every planted SNP must be called (sensitivity)."""
import numpy as np
import pytest

from ebwt2snp_b200 import api
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def find_all(g, s):
    i = g.find(s)
    while i >= 0:
        yield i
        i = g.find(s, i + 1)


def test_planted_snps_are_called(built):
    rs, e = H.dataset("small", 1)
    n = e["n"]
    ctx = api.Context(0)
    sh = ctx.shard(n)
    sh.load_soa(e["lcp"], e["text"], e["suff"], e["bwt"])
    sh.seal()
    ctx.stage_reads(rs.reads, O.uniform_read_offsets(*rs.reads.shape))
    p = api.default_params(rs.nreads1)
    sh.pipeline_resident(p, 16, 2)
    lines = api.events_format(sh.events(), p).decode().split("\n")
    sh.close()
    ctx.close()
    g1 = rs.genome1.tobytes()
    planted = set(int(x) for x in rs.snp_pos)
    called, n_snp_events, unmatched = set(), 0, 0
    for i in range(0, len(lines) - 3, 4):
        if not lines[i].startswith(">SNP"):
            continue
        n_snp_events += 1
        dna0 = lines[i + 1].encode()  # 31 bases ending ON the variant + 30 bases of right context
        hits = [pos + 30 for pos in find_all(g1, dna0)] + \
               [pos + len(dna0) - 31 for pos in find_all(g1, dna0.translate(COMP)[::-1])]
        ok = [h for h in hits if h in planted]
        called.update(ok)
        unmatched += not ok
    assert n_snp_events >= len(planted)                       # most SNPs are seen on both strands
    assert unmatched <= 0.05 * n_snp_events                   # precision
    assert len(called) >= 0.9 * len(planted)                  # sensitivity
