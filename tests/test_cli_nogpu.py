"""not gpu: CLI behaviour that is decided before any GPU work (SURVEY.md §8(b) 'Errors' row)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ebwt2snp_b200", "bin")


def run(tool, *args):
    return subprocess.run([os.path.join(BIN, tool), *args], capture_output=True, text=True)


def test_help_exits_zero(built):
    for tool in ("ebwt2clust", "clust2snp"):
        for args in ([], ["-h"], ["-Q"]):
            r = run(tool, *args)
            assert r.returncode == 0 and tool + " [options]" in r.stdout, (tool, args)  # ref:ebwt2clust.cpp:51
    assert run("ebwt2clust", "-k", "3").returncode == 0           # no -i  -> help
    assert run("clust2snp", "-i", "x.fasta", "-n", "0").returncode == 0  # -n 0 -> help (ref:clust2snp.cpp:1045)
    # -M and -b are not in the option string: help, exit 0 (ref:clust2snp.cpp:976,981,996)
    assert "clust2snp [options]" in run("clust2snp", "-i", "x", "-n", "3", "-M", "100").stdout
    assert "clust2snp [options]" in run("clust2snp", "-i", "x", "-n", "3", "-b").stdout


def test_help_text_is_the_reference_s(built):
    """tests/golden/help_*.txt = what the reference binaries print for -h (tests/golden/make_golden.py); live against
    oracle/_ref when it is there"""
    for tool in ("ebwt2clust", "clust2snp"):
        want = open(os.path.join(ROOT, "tests", "golden", "help_" + tool + ".txt")).read()
        assert run(tool, "-h").stdout == want
        ref = os.path.join(ROOT, "oracle", "_ref", tool)
        if os.access(ref, os.X_OK):
            assert subprocess.run([ref, "-h"], capture_output=True, text=True).stdout == want


def test_missing_index_exits_one(built):
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "none.fasta")
        for tool, extra in (("ebwt2clust", []), ("clust2snp", ["-n", "5"])):
            r = run(tool, "-i", fa, *extra)
            assert r.returncode == 1 and "Error: missing index files." in r.stdout  # ref:include.hpp:72-77


def test_missing_clusters_prints_help(built):
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "r.fasta")
        open(fa + ".gesa", "wb").write(b"\0" * 26)
        r = run("clust2snp", "-i", fa, "-n", "5", "-x", "4", "-y", "4", "-z", "4")
        assert r.returncode == 0 and "ERROR: Could not find BWT clusters file" in r.stdout  # ref:clust2snp.cpp:1063-1068
        assert "Output events will be stored" not in r.stdout


def test_fasta_reader_multiline_and_ragged(built, tmp_path):
    """host_io.hpp's FASTA reader (get_reads, ref:clust2snp.cpp:147-212): wrapped sequences, ragged lengths, a first line
    that is a header whatever it holds, no trailing newline"""
    import textwrap
    src = tmp_path / "t.cpp"
    src.write_text(textwrap.dedent('''
        #include <cstdio>
        #include "host_io.hpp"
        int main(int argc, char** argv) {
            host::Reads r;
            if (!r.load(argv[1])) return 1;
            printf("%zu", size_t(r.n_reads()));
            for (size_t i = 0; i < r.n_reads(); ++i)
                printf(" %.*s", int(r.off[i + 1] - r.off[i]), reinterpret_cast<const char*>(r.bases.data() + r.off[i]));
            printf("\\n");
            return 0;
        }'''))
    exe = tmp_path / "t"
    inc, host, lib = os.path.join(ROOT, "include"), os.path.join(ROOT, "ebwt2snp_b200", "host"), os.path.join(ROOT, "ebwt2snp_b200", "lib")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", inc, "-I", host, str(src), "-o", str(exe), "-L", lib, "-lebwt2snp_b200",
                    f"-Wl,-rpath,{lib}", "-ldl", "-lpthread", "-lrt"], check=True)
    fa = tmp_path / "r.fasta"
    fa.write_text("first line is a header\nACGT\nAC\n>r2 desc\nGGGTTT\n>r3\nA\nC\nG\n>r4\nTTTT")
    out = subprocess.run([str(exe), str(fa)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["4", "ACGTAC", "GGGTTT", "ACG", "TTTT"]


def test_build_gesa_decisions_before_gpu_work(built, tmp_path):
    """build_gesa: help exits 0 like the reference's tools; unreadable input -> 1; nothing but empty reads -> 2 (before any
    CUDA call; ragged reads are fine: tests/test_builder_gpu.py); without a GPU the library refuses loudly (exit 3, no index written) -- there is no CPU builder"""
    for args in ([], ["-h"], ["-Q"], ["-x", "3", "-i", "x"]):
        r = run("build_gesa", *args)
        assert r.returncode == 0 and "build_gesa [options]" in r.stdout, args
    assert run("build_gesa", "-i", str(tmp_path / "missing.fasta")).returncode == 1
    fa = tmp_path / "r.fasta"
    fa.write_text(">a\n\n>b\n")
    r = run("build_gesa", "-i", str(fa))
    assert r.returncode == 2 and "empty" in r.stdout
    import torch
    if not torch.cuda.is_available():
        fa.write_text(">a\nACGT\n>b\nACG\n")
        r = run("build_gesa", "-i", str(fa))
        assert r.returncode == 3 and "no CPU fallback" in r.stdout and not os.path.exists(str(fa) + ".gesa")
