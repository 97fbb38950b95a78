#!/bin/bash
# one-pass scan kernel: parity (all -m gpu tests, slow C2 skipped) then the C3 and C2 bench lines
set -x
mkdir -p gpurun_out
E2S_SKIP_SLOW=1 timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -25 gpurun_out/r2c_pytest.log
if [ $rc -eq 0 ]; then
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2c_bench_c3.err; cat gpurun_out/r2c_bench_c3.json
timeout 900 python bench.py --workload C2 --no-e2e --no-cpu-baseline --no-egsa-build > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2c_bench_c2.err; cat gpurun_out/r2c_bench_c2.json
fi
