#!/bin/bash
# session 4 validation run: whole GPU suite, smoke, the default bench line (C3), reference arm, ncu capture of the final radix pass
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r4f_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r4f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4f_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r4f_bench_c3.json 2> gpurun_out/r4f_bench_c3.err; echo "bench rc=$?"
tail -4 gpurun_out/r4f_bench_c3.err
grep "\[data\]\|\[egsa\]" gpurun_out/r4f_bench_c3.err | head
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4f_bench_reference.json 2> gpurun_out/r4f_bench_reference.err; echo "reference arm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 40 -c 1 -o gpurun_out/r4f_radix_pass ebwt2snp_b200/bin/bench_build_egsa 1000000 100 1 > gpurun_out/r4f_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/r4f_bench_c3.json | cut -c1-1500
