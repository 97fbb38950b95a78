# full single-GPU record: tests, smoke, bench with e2e + cpu baseline + EGSA build, reference arm, launch list, ncu --set full of K1 / K2, tiles8
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r1_vX}
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 80 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-egsa-build > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lcp_flags8|k_cluster_emit' -s 8 -c 2 -f -o gpurun_out/prof_${TAG} python bench.py --steps 2 --no-e2e --no-cpu-baseline --no-egsa-build > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
E2S_NO_FUSED_PREFILTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_code_scan|k_events' -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_phase2 python bench.py --steps 2 --no-e2e --no-cpu-baseline --no-egsa-build > gpurun_out/ncu_full2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --tiles 8 > gpurun_out/bench_${TAG}_tiles8.json 2> gpurun_out/bench_tiles8.err; echo "tiles8 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_c2.json').read().strip().splitlines()[-1])
print('value %.4g pos/s  ms/step %.3f  launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']))
for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
print('  e2e', d['e2e']); print('  cpu', d['cpu_baseline']); print('  egsa', d['egsa_build'])
t=json.loads(open('gpurun_out/bench_${TAG}_tiles8.json').read().strip().splitlines()[-1])
print('tiles8: value %.4g pos/s  ms/step %.3f' % (t['value'], t['ms_per_step']), t['roofline']['pipeline'])
for k,v in t['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
PY
cat gpurun_out/bench_${TAG}_reference.json | cut -c1-400
