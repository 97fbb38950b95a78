#!/bin/bash
# default bench line (C3: resident value, chunked e2e, cpu baseline + CLI leg) and a scaled-down C4 streamed line
set -x
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err; echo "bench rc=$?"
tail -4 gpurun_out/r2h_bench_c3.err; cat gpurun_out/r2h_bench_c3.json
timeout 900 python bench.py --workload C4 --scale 0.03 --steps 2 --warmup 1 > gpurun_out/r2h_bench_c4_small.json 2> gpurun_out/r2h_bench_c4_small.err; echo "c4 rc=$?"
tail -4 gpurun_out/r2h_bench_c4_small.err; cat gpurun_out/r2h_bench_c4_small.json
