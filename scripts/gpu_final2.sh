# final single-GPU record of a build: tests, smoke, default bench (e2e + cpu baseline + EGSA build), reference arm
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r1_vX}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_c2.json').read().strip().splitlines()[-1])
print('value %.4g pos/s  ms/step %.3f  launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']), d['clocks'])
for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
print('  roofline', {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac','traffic')})
print('  e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['parity_vs_reference_on_sample']); print('  egsa', d['egsa_build'])
PY
cut -c1-300 gpurun_out/bench_${TAG}_reference.json
