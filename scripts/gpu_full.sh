# full single-GPU record: tests, bench with e2e + cpu baseline, launch list, ncu --set full of the three streaming kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r1_vX}
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 80 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lcp_flags|k_cluster_emit|k_code_scan' -s 9 -c 3 -f -o gpurun_out/prof_${TAG} python bench.py --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_c2.json').read().strip().splitlines()[-1])
print('value %.4g pos/s  ms/step %.3f  launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']))
for k,v in d['roofline']['kernels'].items(): print('  %-18s %.1f us  %.0f GB/s' % (k, v['ms']*1e3, v['GBps']))
print('  e2e', d['e2e']); print('  cpu', d['cpu_baseline'])
PY
cat gpurun_out/bench_${TAG}_reference.json | cut -c1-400
