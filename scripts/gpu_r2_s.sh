#!/bin/bash
# per-kernel times inside the step (library timers): C3 and C2
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2s_bench_c3.json 2> gpurun_out/r2s_bench_c3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2s_bench_c3.err; cat gpurun_out/r2s_bench_c3.json
timeout 900 python bench.py --workload C2 --no-cpu-baseline --no-egsa-build --no-e2e > gpurun_out/r2s_bench_c2.json 2> gpurun_out/r2s_bench_c2.err; echo "bench rc=$?"
cat gpurun_out/r2s_bench_c2.json
