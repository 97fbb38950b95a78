#!/bin/bash
# 2 GPUs: the whole -m gpu suite (named configs, C3 windows, multi-GPU NCCL parity) + the strong-scaling C3 bench line at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2_full_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_full_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n2.json 2> gpurun_out/r2_bench_c3_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/r2_bench_c3_n2.err; cat gpurun_out/r2_bench_c3_n2.json
