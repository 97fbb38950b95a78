# multi-GPU check (run with gpurun --gpus N): NCCL parity test, sharded CLIs, bench at N ranks
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L > gpurun_out/smi_multi.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -x -q -k "ranks or sharded" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=%d value %.4g pos/s  ms/step %.3f e2e %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['value']))
PY
