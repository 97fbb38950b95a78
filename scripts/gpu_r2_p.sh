#!/bin/bash
# launch list of the default bench step (C3, N = 1): per-launch durations of this library's kernels
set -x
mkdir -p gpurun_out
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 14 -c 40 --csv --log-file gpurun_out/r2_v9_launches_c3.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_list.log; grep -c k_ gpurun_out/r2_v9_launches_c3.csv
