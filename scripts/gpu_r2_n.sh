#!/bin/bash
# round-2 validation: smoke, full GPU suite, the default bench line (C3, N = 1), one ncu capture of the scan at C3
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2t_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r2t_pytest.log
timeout 1200 python bench.py > gpurun_out/r2t_bench_c3.json 2> gpurun_out/r2t_bench_c3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2t_bench_c3.err; cat gpurun_out/r2t_bench_c3.json
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cluster_scan -s 3 -c 1 -f -o gpurun_out/prof_r2_scan_final_c3 $CMD > gpurun_out/ncu_scan.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scan.log
