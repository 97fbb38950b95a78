#!/bin/bash
# ncu capture of the phase-2 tail kernels at C3
set -x
mkdir -p gpurun_out
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_events|k_candidates|k_cluster_exact" -s 3 -c 3 -f -o gpurun_out/prof_r2_tail_c3 $CMD > gpurun_out/ncu_tail.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tail.log
