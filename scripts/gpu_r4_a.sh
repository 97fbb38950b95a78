#!/bin/bash
# session 4, first run: the new index builder (own radix sort, ragged reads, 64-bit ids) against the oracle, the tests that use
# the builder, and its time next to the cub-driven loop of round 1
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_builder_gpu.py -m gpu -q -x --durations=5 > gpurun_out/r4a_builder.log 2>&1; echo "builder rc=$?"
tail -25 gpurun_out/r4a_builder.log
timeout 120 ebwt2snp_b200/bin/bench_build_egsa 1000000 100 3 > gpurun_out/r4a_sort_1m.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4a_sort_1m.json
timeout 200 ebwt2snp_b200/bin/bench_build_egsa 5520000 100 3 > gpurun_out/r4a_sort_c2.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4a_sort_c2.json
timeout 300 python -m pytest tests/test_streaming_gpu.py tests/test_named_configs_gpu.py -m gpu -q -k "150bp or builder or c1 or C1" > gpurun_out/r4a_users.log 2>&1; echo "users rc=$?"
tail -5 gpurun_out/r4a_users.log
