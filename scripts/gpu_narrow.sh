# narrow-LCP K1: parity tests, bench with the one-byte and the four-byte LCP stream, ncu of k_lcp_flags8
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r1_v17}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_narrow.json 2> gpurun_out/bench_c2.err; echo "bench narrow rc=$?"
E2S_LCP_WIDE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_wide.json 2>> gpurun_out/bench_c2.err; echo "bench wide rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lcp_flags8' -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_flags8 python bench.py --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for w in ('narrow','wide'):
    d=json.loads(open('gpurun_out/bench_${TAG}_c2_%s.json' % w).read().strip().splitlines()[-1])
    print(w, 'value %.4g pos/s  ms/step %.3f  launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']), d['clocks'])
    for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
    print('  pipeline', d['roofline']['pipeline'], d['results']['n_written'], d['results']['n_events_rank0'])
PY
