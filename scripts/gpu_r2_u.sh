#!/bin/bash
# 2 GPUs: NCCL parity tests (torch.distributed exchange, library exchange, chunked sharded pipeline), CLI sharding, strong-scaling bench line
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q --durations=5 > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2u_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2u_bench_c3_n2.json 2> gpurun_out/r2u_bench_c3_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2u_bench_c3_n2.err; cat gpurun_out/r2u_bench_c3_n2.json
