"""not gpu: the C-ABI library loads, exports every symbol the header declares, its structs have the layout
the ctypes mirror assumes, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from ebwt2snp_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ebwt2snp_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(e2s_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(built):
    lib = api.load_library()
    syms = header_symbols()
    assert syms == sorted(api.SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.e2s_version() == 100


def test_struct_layouts(built):
    prog = r'''
#include <stdio.h>
#include "ebwt2snp_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(e2s_cluster_summary), sizeof(e2s_cluster_merged), sizeof(e2s_stats),
         sizeof(e2s_snp_params), sizeof(e2s_event), sizeof(e2s_snp_counts), sizeof(e2s_pipeline_result));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    mine = [C.sizeof(t) for t in (api.ClusterSummary, api.ClusterMerged, api.Stats, api.SnpParams, api.Event,
                                  api.SnpCounts, api.PipelineResult)]
    assert [int(x) for x in out] == mine


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.E2SError) as ei:
        api.Context(0)
    assert ei.value.code == api.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/"""
    pkg = os.path.join(ROOT, "ebwt2snp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle.h" not in txt, f


def test_events_format_host_only(built):
    p = api.default_params(10)
    ev = api.Event()
    ev.D, ev.gap, ev.supp0, ev.supp1, ev.right_len, ev.keep = 0, 0, 7, 9, 30, 1
    ev.left0 = b"A" * 30 + b"C"
    ev.left1 = b"A" * 30 + b"G"
    ev.right = b"T" * 30
    txt = api.events_format([ev], p)
    assert txt == (b">SNP_higher_path_1|P_1:30_C/G|7|nb_pol_1\n" + b"A" * 30 + b"C" + b"T" * 30 + b"\n" +
                   b">SNP_lower_path_1|P_1:30_C/G|9|nb_pol_1\n" + b"A" * 30 + b"G" + b"T" * 30 + b"\n")
    ev.gap = 2
    ev.left0 = b"A" * 29 + b"CG"
    txt = api.events_format([ev], p, first_id=5)
    assert txt.startswith(b">INDEL_higher_path_5|P_1:30_CG/|7|nb_pol_1\n" + b"A" * 29 + b"CG" + b"T" * 30 + b"\n" +
                          b">INDEL_lower_path_5|P_1:30_CG/|9|nb_pol_1\n" + b"A" * 28 + b"G" + b"T" * 30)
