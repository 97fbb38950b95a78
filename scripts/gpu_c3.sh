# C3 (64 Mbp genome, 2 x 30x, SNPs + indels: 3.88e9 positions) on ONE GPU, index built by e2s_build_egsa_dev
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG:-r1}_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
tail -6 gpurun_out/bench_c3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG:-r1}_c3.json').read().strip().splitlines()[-1])
print('C3: value %.4g pos/s  ms/step %.3f  n=%d' % (d['value'], d['ms_per_step'], d['config']['positions_total']), d['index_check'], d['clocks'])
for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
print(d['roofline']['pipeline'], d['results'])
PY
