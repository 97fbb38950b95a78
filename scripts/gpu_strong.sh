# C5: strong scaling on the C3 workload (one 3.88e9-position eBWT cut into N contiguous ranges), run with gpurun --gpus N
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
if [ "${SMALL:-1}" = "1" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --workload C2 --scale 0.1 --scaling strong --no-e2e --no-cpu-baseline > gpurun_out/bench_strong_small_n$N.json 2> gpurun_out/bench_strong_small.err; echo "small strong rc=$?"
tail -2 gpurun_out/bench_strong_small.err
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --workload C3 --scaling strong --no-cpu-baseline > gpurun_out/bench_${TAG:-r1}_c3_strong_n$N.json 2> gpurun_out/bench_strong.err; echo "C3 strong rc=$?"
grep -v "^\*\*\|^$" gpurun_out/bench_strong.err | tail -6
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG:-r1}_c3_strong_n$N.json').read().strip().splitlines()[-1])
print('C3 strong N=%d: value %.4g pos/s  ms/step %.3f scaling=%s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['scaling']), d['results'])
PY
