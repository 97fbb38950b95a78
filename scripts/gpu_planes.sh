# resident planes + narrow K1: parity tests, bench (fused and two-phase), launch list
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r1_v18}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
E2S_NO_FUSED_PREFILTER=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_twophase.json 2>> gpurun_out/bench_c2.err; echo "bench two-phase rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
python - <<PY
import json
for w in ('c2','c2_twophase'):
    d=json.loads(open('gpurun_out/bench_${TAG}_%s.json' % w).read().strip().splitlines()[-1])
    print(w, 'value %.4g pos/s  ms/step %.3f  launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']), d['clocks'])
    for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
    print('  pipeline', d['roofline']['pipeline'], d['results']['n_written'], d['results']['n_events_rank0'])
PY
tail -8 gpurun_out/${TAG}_launches_c2.csv | cut -d, -f5,12-; grep egsa gpurun_out/bench_c2.err
