cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --tiles 8 > gpurun_out/bench_tiles8.json 2> gpurun_out/bench_tiles8.err; echo "rc=$?"
tail -2 gpurun_out/bench_tiles8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tiles8.json').read().strip().splitlines()[-1])
print('tiles8: value %.4g pos/s  ms/step %.3f  positions %d' % (d['value'], d['ms_per_step'], d['config']['positions_total']))
for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
print(d['roofline']['pipeline'], d['results'])
PY
