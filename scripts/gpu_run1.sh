set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; free -g >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lcp_flags|k_cluster_emit|k_code_scan' -s 9 -c 3 -o gpurun_out/prof_r1_v3 python bench.py --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_c2.json | cut -c1-600
