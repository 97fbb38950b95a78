#!/bin/bash
# final 1-GPU run of the session: whole GPU suite, smoke, the default bench line (C3: value, e2e, e2e.soa, cpu baseline, CLI leg), reference arm,
# ncu launch list + full capture of the scan
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r3g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3g_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3g_bench_c3.json 2> gpurun_out/r3g_bench_c3.err; echo "bench rc=$?"
tail -4 gpurun_out/r3g_bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3g_bench_reference.json 2> gpurun_out/r3g_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3g_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r3g_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cluster_scan -s 3 -c 1 -o gpurun_out/r3g_scan_c3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r3g_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/r3g_bench_c3.json | cut -c1-3000
