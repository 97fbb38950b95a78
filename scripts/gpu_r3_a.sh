set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g > gpurun_out/r3a_mem.txt; nproc >> gpurun_out/r3a_mem.txt
timeout 600 python -m pytest tests/test_streaming_gpu.py -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r3a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3a_bench_c3.json 2> gpurun_out/r3a_bench_c3.err; echo "bench rc=$?"
tail -5 gpurun_out/r3a_bench_c3.err
cat gpurun_out/r3a_bench_c3.json
