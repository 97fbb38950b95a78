"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference binaries
(oracle/_ref, built from /root/reference by oracle/Makefile) on seeded inputs.

    python tests/golden/make_golden.py

Only runs where oracle/_ref exists (the build container).  The fixtures hold the inputs in compact form
(2-bit packed reads + generator parameters; the EGSA is rebuilt deterministically by
ebwt2snp_b200.synth and checked against the recorded sha256 of the .gesa the reference consumed) and the
reference's own outputs (.clusters records, .snp text, printed counts)."""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ebwt2snp_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GESA_DT = np.dtype([("text", "<u4"), ("suff", "<u4"), ("lcp", "<u4"), ("bwt", "u1")])

MICRO = {
    "micro_a": dict(gen=dict(G=6000, reads_per_sample=1200, L=100, n_snps=12, n_indels=3, rc=True, seed=11), k=16, m=2),
    "micro_b": dict(gen=dict(G=4000, reads_per_sample=1500, L=64, n_snps=10, n_indels=2, rc=False, seed=12), k=12, m=3),
}
SNP_VARIANTS = [
    [], ["-m", "3"], ["-c", "3", "-g", "4"], ["-L", "25", "-R", "20", "-e", "1"], ["-p", "0.5"], ["-v", "0"],
]


def pack2(reads):
    code = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), reads).astype(np.uint8)
    flat = code.reshape(-1)
    pad = (-len(flat)) % 4
    flat = np.concatenate([flat, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (flat[:, 0] | (flat[:, 1] << 2) | (flat[:, 2] << 4) | (flat[:, 3] << 6)).astype(np.uint8)


def phase1_fuzz(d):
    rng = np.random.default_rng(2024)
    fa = os.path.join(d, "F.fasta")
    open(fa, "w").write(">a\nA\n")
    alphabet = np.frombuffer(b"ACGT$acgt\x00\xff", dtype=np.uint8)
    lcps, bwts, meta, outs = [], [], [], []
    for it in range(72):
        n = int(rng.integers(2, 40)) if it % 3 == 0 else int(rng.integers(40, 2500))
        k = int(rng.choice([1, 2, 3, 5, 16, 70]))
        m = int(rng.choice([1, 2, 3, 8]))
        mode = it % 4
        if mode == 0:
            lcp = rng.integers(0, 2 * k + 2, size=n)
        elif mode == 1:
            lcp = rng.integers(0, 70000, size=n)
        elif mode == 2:
            lcp = np.maximum(0, np.cumsum(rng.integers(-3, 4, size=n)) + k)
        else:
            lcp = rng.integers(k, k + 3, size=n)
        lcp = lcp.astype(np.uint32)
        bwt = rng.choice(alphabet, size=n)
        rec = np.zeros(n, dtype=GESA_DT)
        rec["lcp"], rec["bwt"] = lcp, bwt
        rec["text"] = rng.integers(0, 1000, size=n)
        rec["suff"] = rng.integers(0, 100, size=n)
        rec.tofile(fa + ".gesa")
        r, ncl = O.ref_ebwt2clust(fa, k=k, m=m)
        assert r.returncode == 0
        out = np.frombuffer(open(fa + ".clusters", "rb").read(), dtype=np.uint8)
        lcps.append(lcp)
        bwts.append(bwt)
        outs.append(out)
        meta.append((n, k, m, ncl, len(out)))
    np.savez_compressed(os.path.join(HERE, "phase1_fuzz.npz"), lcp=np.concatenate(lcps), bwt=np.concatenate(bwts),
                        out=np.concatenate(outs), meta=np.array(meta, dtype=np.int64))


def micro(name, cfg, d):
    rs = synth.make_read_set(**cfg["gen"])
    e = synth.build_egsa(rs.reads)
    fa = synth.write_dataset(d, rs, e, name=name + ".fasta")
    sha = hashlib.sha256(open(fa + ".gesa", "rb").read()).hexdigest()
    r, ncl = O.ref_ebwt2clust(fa, k=cfg["k"], m=cfg["m"])
    assert r.returncode == 0
    clusters = np.frombuffer(open(fa + ".clusters", "rb").read(), dtype=np.uint8)
    snp_path = os.path.join(d, name + ".snp")
    res = {}
    for vi, extra in enumerate(SNP_VARIANTS):
        if os.path.exists(snp_path):
            os.remove(snp_path)
        r, info = O.ref_clust2snp(fa, rs.nreads1, extra=extra)
        res[f"v{vi}_args"] = np.array(" ".join(extra))
        res[f"v{vi}_rc"] = np.int64(info["returncode"])
        res[f"v{vi}_allowed"] = np.array(info.get("allowed", (-1, -1)), dtype=np.int64)
        res[f"v{vi}_ncand"] = np.int64(info.get("n_candidates", -1))
        res[f"v{vi}_snp"] = np.frombuffer(open(snp_path, "rb").read() if os.path.exists(snp_path) else b"", dtype=np.uint8)
    gen = cfg["gen"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), reads2bit=pack2(rs.reads), shape=np.array(rs.reads.shape),
                        nreads1=np.int64(rs.nreads1), gen=np.array(repr(gen)), k=np.int64(cfg["k"]), m=np.int64(cfg["m"]),
                        gesa_sha256=np.array(sha), n_clust_out=np.int64(ncl), clusters=clusters, n_variants=np.int64(len(SNP_VARIANTS)),
                        **res)
    print(name, "n =", e["n"], "clusters =", len(clusters) // 10, "closed =", ncl,
          "events(v0) =", bytes(res["v0_snp"]).count(b">") // 2)


if __name__ == "__main__":
    assert O.ref_available(), "oracle/_ref missing: run `make -C oracle ref` where /root/reference exists"
    d = tempfile.mkdtemp(prefix="golden_")
    try:
        phase1_fuzz(d)
        for name, cfg in MICRO.items():
            micro(name, cfg, d)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print("golden vectors written to", HERE)
