"""not gpu: filter_snp and snp2fastq against the committed outputs of the unmodified reference
tools on the fixtures' .snp files and on damaged copies (tests/golden/make_golden.py:snp_text_tools)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ebwt2snp_b200", "bin")
GOLD = os.path.join(ROOT, "tests", "golden", "snp_text.npz")


def test_filter_snp_and_snp2fastq_vs_reference(built, tmp_path):
    z = np.load(GOLD)
    for j in range(int(z["n"])):
        path = str(tmp_path / f"t{j}.snp")
        open(path, "wb").write(z[f"in{j}"].tobytes())
        for M in (0, 5, 9):
            r = subprocess.run([os.path.join(BIN, "filter_snp"), path, str(M)], capture_output=True, timeout=60)
            assert r.returncode == 0 and r.stdout == z[f"filter{j}_{M}"].tobytes(), (j, M)
        for nflag, flag in ((0, []), (1, ["-i"])):
            if os.path.exists(path + ".fastq"):
                os.remove(path + ".fastq")
            r = subprocess.run([os.path.join(BIN, "snp2fastq"), path, *flag], capture_output=True, timeout=60)
            assert r.returncode == 0
            assert open(path + ".fastq", "rb").read() == z[f"fastq{j}_{nflag}"].tobytes(), (j, flag)


def test_text_tools_help(built):
    for tool, args in (("filter_snp", []), ("filter_snp", ["x"]), ("snp2fastq", []), ("snp2fastq", ["x", "-q"])):
        r = subprocess.run([os.path.join(BIN, tool), *args], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and tool + " calls.snp" in r.stdout  # help exits 0 (ref:filter_snp.cpp:20, ref:snp2fastq.cpp:26)
