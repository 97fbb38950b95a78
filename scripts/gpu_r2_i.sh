#!/bin/bash
# bit-sliced LCP scan: parity suites, then the C2 bench line (kernel times), then one ncu capture of the scan
set -x
mkdir -p gpurun_out
E2S_SKIP_SLOW=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_streaming_gpu.py "tests/test_named_configs_gpu.py::test_named_config_file_vs_file" -m gpu -x -q --durations=5 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --workload C2 --no-cpu-baseline --no-egsa-build > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2m_bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2m_bench_c2.err; cat gpurun_out/r2m_bench_c2.json
CMD="python bench.py --workload C2 --no-e2e --no-cpu-baseline --no-egsa-build --steps 2 --warmup 3"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cluster_scan -s 3 -c 1 -f -o gpurun_out/prof_r2_scan_v8 $CMD > gpurun_out/ncu_scan.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scan.log
