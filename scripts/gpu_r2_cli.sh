#!/bin/bash
# where the drop-in CLIs spend their wall time (E2S_CLI_TIMING stamps), C3 scaled to 3e7 positions, files in /dev/shm
set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2_cli_timing.txt 2>&1
import os, sys, time, subprocess, shutil, tempfile
sys.path.insert(0, '.')
import bench
d, fasta, rs, eg, scale = bench.reference_sample('C3', 1001, 30_000_000, 'cuda')
n = int(eg['n']); del eg
import torch; torch.cuda.empty_cache()
print('n =', n, 'gesa bytes =', os.path.getsize(fasta + '.gesa'), flush=True)
B = 'ebwt2snp_b200/bin/'
def run(cmd, **env):
    e = dict(os.environ); e.update(env); e['E2S_CLI_TIMING'] = '1'
    t = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True, env=e); dt = time.perf_counter() - t
    print('$', ' '.join(cmd[:1]), env, '-> %.3f s rc=%d' % (dt, r.returncode)); print(r.stderr, flush=True)
    return dt
for rep in range(2):
    run([B + 'ebwt2clust', '-i', fasta, '-x', '4', '-y', '4', '-z', '4'])
    run([B + 'clust2snp', '-i', fasta, '-n', str(rs.nreads1), '-x', '4', '-y', '4', '-z', '4'])
run([B + 'ebwt2clust', '-i', fasta, '-x', '4', '-y', '4', '-z', '4'], E2S_CHUNK_POSITIONS='33554432')
run([B + 'ebwt2clust', '-i', fasta, '-x', '4', '-y', '4', '-z', '4'], E2S_SCAN_LEGACY='1')
shutil.rmtree(d)
PY
echo "rc=$?"; cat gpurun_out/r2_cli_timing.txt
