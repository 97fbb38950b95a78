#!/bin/bash
# failing-test recheck + launch list of our kernels in the C2 step
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/r2f_pytest.log
CMD="python bench.py --workload C2 --no-e2e --no-cpu-baseline --no-egsa-build --steps 3 --warmup 3"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 14 -c 28 --csv --log-file gpurun_out/r2_launches_c2.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
