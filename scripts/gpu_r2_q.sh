#!/bin/bash
# phase-2 tail changes (K4 fast path, K3x -> K3b hand-over): parity suites, then the C3 step and the C2 step
set -x
mkdir -p gpurun_out
E2S_SKIP_SLOW=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_streaming_gpu.py tests/test_cli_gpu.py tests/test_accuracy_gpu.py "tests/test_named_configs_gpu.py::test_named_config_file_vs_file" -m gpu -x -q --durations=5 > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2q_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c3.json 2> gpurun_out/r2q_bench_c3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2q_bench_c3.err; cat gpurun_out/r2q_bench_c3.json
timeout 900 python bench.py --workload C2 --no-cpu-baseline --no-egsa-build --no-e2e > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench_c2.err; echo "bench rc=$?"
cat gpurun_out/r2q_bench_c2.json
