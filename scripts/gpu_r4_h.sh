#!/bin/bash
# session 4, 2 GPUs: the NCCL parity test and the default bench line at N = 2 (strong scaling on C3, e2e through e2s_pipeline_host_sharded)
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r4h_pytest_multi.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r4h_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4h_bench_n2.json 2> gpurun_out/r4h_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r4h_bench_n2.err
cut -c1-600 gpurun_out/r4h_bench_n2.json
