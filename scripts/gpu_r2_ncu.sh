#!/bin/bash
# ncu capture of the one-pass scan kernel at C2 (after the same command exited 0 without ncu)
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload C2 --no-e2e --no-cpu-baseline --no-egsa-build --steps 2 --warmup 3"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cluster_scan -s 3 -c 1 -f -o gpurun_out/prof_r2_scan $CMD > gpurun_out/ncu_scan.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scan.log
