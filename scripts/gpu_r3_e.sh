set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_streaming_gpu.py tests/test_named_configs_gpu.py -m gpu -x -q -k "not C2 and not c3_windows" > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3e_pytest.log
timeout 600 python bench.py --workload C2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r3e_bench_c2.json 2> gpurun_out/r3e_bench_c2.err; echo "bench c2 rc=$?"
E2S_SCAN_OCC=3 timeout 600 python bench.py --workload C2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r3e_bench_c2_occ3.json 2> gpurun_out/r3e_bench_c2_occ3.err; echo "bench c2 occ3 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r3e_bench_c3.json 2> gpurun_out/r3e_bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r3e_bench_c2.json", "gpurun_out/r3e_bench_c2_occ3.json", "gpurun_out/r3e_bench_c3.json"):
    try:
        d = json.load(open(f))
        print(f, d["ms_per_step"], d["roofline"]["frac"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
