# multi-GPU check (run with gpurun --gpus N): NCCL parity test (both exchanges), bench at N ranks with the native and the torch.distributed exchange
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -15 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n${N}_native.json 2> gpurun_out/bench_n$N.err; echo "bench native rc=$?"
tail -3 gpurun_out/bench_n$N.err
E2S_PY_EXCHANGE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n${N}_py.json 2>> gpurun_out/bench_n$N.err; echo "bench py rc=$?"
python - <<PY
import json
for w in ('native','py'):
    d=json.loads(open('gpurun_out/bench_n${N}_%s.json' % w).read().strip().splitlines()[-1])
    print(w, 'N=%d value %.4g pos/s  ms/step %.3f exchange_us %s launches %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['config']['exchange_us'], d['gpu_launches']))
PY
