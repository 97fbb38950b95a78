#!/bin/bash
# pipelined one-pass scan + narrow copies written by the loads: parity on the fast suites, then the bench lines
set -x
mkdir -p gpurun_out
E2S_SKIP_SLOW=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_build_egsa.py tests/test_cli_gpu.py "tests/test_named_configs_gpu.py::test_named_config_file_vs_file" -m gpu -x -q --durations=5 > gpurun_out/r2d_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -15 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --workload C2 --no-e2e --no-cpu-baseline --no-egsa-build > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench_c2.err; cat gpurun_out/r2d_bench_c2.json
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench_c3.err; cat gpurun_out/r2d_bench_c3.json
