#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite (named-config parity included) + the default bench (C3) on the round-1 kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; df -h /dev/shm | tail -1; nproc
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_bench_c3.err; cat gpurun_out/r2a_bench_c3.json
