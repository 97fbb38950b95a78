"""not gpu: snp_vs_vcf against the committed stdout of the unmodified reference tool on
planted-truth inputs (tests/golden/make_golden.py:snp_vs_vcf_inputs): the generator's SNPs as VCF, sample 1's genome as
reference FASTA (wrapped, mixed case, one or two contigs), the oracle's .snp as calls, plus damaged / empty / malformed
calls files.  Where oracle/_ref is present the reference is also run live on the same files."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "ebwt2snp_b200", "bin", "snp_vs_vcf")
GOLD = os.path.join(ROOT, "tests", "golden", "snp_vs_vcf.npz")


def _cases(tmp_path):
    z = np.load(GOLD)
    for j in range(int(z["n"])):
        paths = {}
        for k in ("fasta", "vcf", "calls"):
            paths[k] = str(tmp_path / f"c{j}.{k}")
            open(paths[k], "wb").write(z[f"{j}_{k}"].tobytes())
        yield str(z[f"{j}_name"]), paths, [str(a) for a in z[f"{j}_argv"]], z[f"{j}_stdout"].tobytes(), int(z[f"{j}_rc"])


def test_snp_vs_vcf_vs_reference_golden(built, tmp_path):
    seen = 0
    for name, paths, argv, want, rc in _cases(tmp_path):
        cmd = ["-v", paths["vcf"], "-c", paths["calls"], "-f", paths["fasta"], *argv]
        r = subprocess.run([TOOL, *cmd], capture_output=True, timeout=120)
        assert r.returncode == rc, name
        assert r.stdout == want, name
        ref = os.path.join(O.REF_DIR, "snp_vs_vcf")
        if os.access(ref, os.X_OK):
            live = subprocess.run([ref, *cmd], capture_output=True, timeout=120)
            assert live.stdout == want and live.returncode == rc, name
        seen += 1
    assert seen >= 6


def test_snp_vs_vcf_planted_truth_is_found(built, tmp_path):
    """the semantic check beside bit parity: on the planted set the calls of the pipeline find (nearly) all planted SNPs
    and invent none"""
    name, paths, argv, want, rc = next(_cases(tmp_path))
    assert name == "planted"
    out = subprocess.run([TOOL, "-v", paths["vcf"], "-c", paths["calls"], "-f", paths["fasta"]], capture_output=True, text=True,
                         timeout=120).stdout
    val = {l.split("=")[0].strip(): l.split("=")[-1].strip() for l in out.splitlines() if l.startswith(("TP", "FP", "FN"))}
    assert int(val["FP"]) == 0 and int(val["TP"]) >= 38 and int(val["FN"]) <= 2


def test_snp_vs_vcf_help(built):
    for args in ([], ["-h"], ["-v", "x", "-c", "y"], ["-v", "x", "-c", "y", "-Q", "z"]):
        r = subprocess.run([TOOL, *args], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and "snp_vs_vcf [options]" in r.stdout  # help exits 0 (ref:snp_vs_vcf.cpp:37)
