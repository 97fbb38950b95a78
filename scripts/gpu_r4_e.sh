#!/bin/bash
# session 4: builder v4 (write-out unrolled, finish with neighbour shuffle)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_builder_gpu.py -m gpu -q --durations=3 > gpurun_out/r4e_builder.log 2>&1; echo "builder rc=$?"
tail -8 gpurun_out/r4e_builder.log
timeout 120 ebwt2snp_b200/bin/bench_build_egsa 1000000 100 3 > gpurun_out/r4e_sort_1m.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4e_sort_1m.json
timeout 200 ebwt2snp_b200/bin/bench_build_egsa 5520000 100 3 > gpurun_out/r4e_sort_c2.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4e_sort_c2.json
timeout 200 ebwt2snp_b200/bin/bench_build_egsa 3700000 150 3 > gpurun_out/r4e_sort_150.json 2>&1; echo "tool rc=$?"; cat gpurun_out/r4e_sort_150.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4e_builder_launches.csv ebwt2snp_b200/bin/bench_build_egsa 1000000 100 1 > gpurun_out/r4e_ncu_list.log 2>&1; echo "ncu list rc=$?"
