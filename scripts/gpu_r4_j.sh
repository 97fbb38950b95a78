#!/bin/bash
# session 4 final validation run: whole GPU suite, smoke, the default bench line (C3), reference arm, ncu capture of the final radix pass
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r4j_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r4j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4j_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r4j_bench_c3.json 2> gpurun_out/r4j_bench_c3.err; echo "bench rc=$?"
tail -4 gpurun_out/r4j_bench_c3.err
grep "\[data\]\|\[egsa\]" gpurun_out/r4j_bench_c3.err | head
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4j_bench_reference.json 2> gpurun_out/r4j_bench_reference.err; echo "reference arm rc=$?"
cat gpurun_out/r4j_bench_c3.json | cut -c1-1500
