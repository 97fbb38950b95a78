#!/bin/bash
# fd loader + quick-exit CLIs: tests, CLI phase timings at two sizes, C2 bench line for the scan with the 64-position bound step
set -x
mkdir -p gpurun_out
E2S_SKIP_SLOW=1 timeout 1500 python -m pytest tests/test_streaming_gpu.py tests/test_cli_gpu.py tests/test_gpu_parity.py "tests/test_named_configs_gpu.py::test_named_config_file_vs_file" -m gpu -x -q --durations=5 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2o_pytest.log
python - <<'PY' > gpurun_out/r2o_cli_timing.txt 2>&1
import os, sys, time, subprocess, shutil
sys.path.insert(0, '.')
import bench, torch
from oracle import oracle as O
B = 'ebwt2snp_b200/bin/'
def run(cmd, **env):
    e = dict(os.environ); e.update(env); e['E2S_CLI_TIMING'] = '1'
    t = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True, env=e); dt = time.perf_counter() - t
    print('$', ' '.join(cmd[:1]), env, '-> %.3f s rc=%d' % (dt, r.returncode)); print(r.stderr, flush=True)
    return dt
for target in (3e7, 1.2e8):
    d, fasta, rs, eg, scale = bench.reference_sample('C3', 1001, target, 'cuda')
    n = int(eg['n']); del eg; torch.cuda.empty_cache()
    print('==== n =', n, 'gesa bytes =', os.path.getsize(fasta + '.gesa'), flush=True)
    for rep in range(2):
        run([B + 'ebwt2clust', '-i', fasta, '-x', '4', '-y', '4', '-z', '4'])
        run([B + 'clust2snp', '-i', fasta, '-n', str(rs.nreads1), '-x', '4', '-y', '4', '-z', '4'])
    run([B + 'ebwt2clust', '-i', fasta, '-x', '4', '-y', '4', '-z', '4'], E2S_CLI_MMAP='1')
    run([B + 'clust2snp', '-i', fasta, '-n', str(rs.nreads1), '-x', '4', '-y', '4', '-z', '4'], E2S_CLI_MMAP='1')
    run([B + 'ebwt2clust', '-i', fasta, '-x', '4', '-y', '4', '-z', '4'], E2S_READ_THREADS='8')
    shutil.rmtree(d)
PY
echo "cli rc=$?"; cat gpurun_out/r2o_cli_timing.txt
timeout 900 python bench.py --workload C2 --no-cpu-baseline --no-egsa-build --no-e2e > gpurun_out/r2o_bench_c2.json 2> gpurun_out/r2o_bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2o_bench_c2.err; cat gpurun_out/r2o_bench_c2.json
