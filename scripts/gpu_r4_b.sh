#!/bin/bash
# session 4: builder tests again, per-kernel times of the builder (own radix vs the cub loop) and one full capture of a radix pass
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_builder_gpu.py -m gpu -q --durations=5 > gpurun_out/r4b_builder.log 2>&1; echo "builder rc=$?"
tail -12 gpurun_out/r4b_builder.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4b_builder_launches.csv ebwt2snp_b200/bin/bench_build_egsa 1000000 100 1 > gpurun_out/r4b_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 30 -c 1 -o gpurun_out/r4b_radix_pass ebwt2snp_b200/bin/bench_build_egsa 1000000 100 1 > gpurun_out/r4b_ncu_full.log 2>&1; echo "ncu full rc=$?"
