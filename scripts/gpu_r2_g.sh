#!/bin/bash
# chunked shards (streaming): tests, then the C2 bench line with e2e through the chunked e2s_pipeline_host
set -x
mkdir -p gpurun_out
E2S_SKIP_SLOW=1 timeout 1500 python -m pytest tests/test_streaming_gpu.py tests/test_gpu_parity.py "tests/test_named_configs_gpu.py::test_named_config_file_vs_file" -m gpu -x -q --durations=5 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --workload C2 --no-cpu-baseline --no-egsa-build > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2g_bench_c2.err; cat gpurun_out/r2g_bench_c2.json
