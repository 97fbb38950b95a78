# quick GPU check: parity tests + one bench line (no e2e / cpu baseline unless FULL=1)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
if [ "$FULL" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
else
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
fi
tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
print('value %.4g pos/s  ms/step %.3f  launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']))
r=d['roofline']
for k,v in r['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
print('  kernel share', r['kernel_share_of_step'], 'e2e', d.get('e2e'))
print('  results', d['results'])
PY
