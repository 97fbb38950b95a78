# scaling check on one box (run with gpurun --gpus 8): NCCL parity test, then bench at N = 1, 2, 4, 8
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_scale.txt
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_scale.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_scale.log
tail -3 gpurun_out/pytest_scale.log
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1])
    print('N=%d value %.4g pos/s  ms/step %.3f e2e %.4g' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'] if d['e2e'] else 0))
except Exception as e:
    print('N=$N failed', e); print(open('gpurun_out/scale_n$N.err').read()[-1500:])
PY
done
