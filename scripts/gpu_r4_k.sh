#!/bin/bash
# session 4: k_chunk_resolve / k_merge_stats tweaks -- parity tests that go through them, then the resident step
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_streaming_gpu.py tests/test_accuracy_gpu.py -m gpu -q -x > gpurun_out/r4k_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r4k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r4k_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r4k_bench_c3.json 2> gpurun_out/r4k_bench_c3.err; echo "bench rc=$?"
tail -2 gpurun_out/r4k_bench_c3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r4k_bench_c3.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], {k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}, d['results'])
PY
