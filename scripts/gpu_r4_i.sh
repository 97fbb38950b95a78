#!/bin/bash
# session 4: radix pass with 512 threads x 8 pairs next to 256 x 16 (same tile), both id widths; builder tests with both
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for T in 256 512; do
  E2S_RADIX_THREADS=$T timeout 200 ebwt2snp_b200/bin/bench_build_egsa 5520000 100 3 > gpurun_out/r4i_sort_c2_t$T.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4i_sort_c2_t$T.json
  E2S_RADIX_THREADS=$T E2S_BUILD_IDS64=1 timeout 200 ebwt2snp_b200/bin/bench_build_egsa 5520000 100 3 > gpurun_out/r4i_sort_c2_ids64_t$T.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4i_sort_c2_ids64_t$T.json
done
E2S_RADIX_THREADS=512 timeout 400 python -m pytest tests/test_builder_gpu.py tests/test_build_egsa.py -m gpu -q > gpurun_out/r4i_builder_t512.log 2>&1; echo "builder 512 rc=$?"
tail -3 gpurun_out/r4i_builder_t512.log
