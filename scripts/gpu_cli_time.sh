# wall time of the drop-in CLIs against the reference binaries on the same files (C2 scaled x0.25, files in /dev/shm)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/cli_time.txt
import os, sys, time, subprocess, shutil, tempfile
sys.path.insert(0, '.')
import numpy as np, torch
from ebwt2snp_b200 import synth
from oracle import oracle as O
rs = synth.make_config('C2', seed=3, scale=0.25)
eg = synth.build_egsa(rs.reads, device='cuda')
d = tempfile.mkdtemp(prefix='e2s_cli_', dir='/dev/shm')
fa = synth.write_dataset(d, rs, eg)
n = int(eg['n']); del eg; torch.cuda.empty_cache()
print('n =', n, 'gesa bytes =', os.path.getsize(fa + '.gesa'))
def run(cmd, env=None):
    t = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True, env=env); return time.perf_counter() - t, r
B = 'ebwt2snp_b200/bin/'
for rep in range(2):
    t1, r1 = run([B + 'ebwt2clust', '-i', fa, '-x', '4', '-y', '4', '-z', '4'])
    t2, r2 = run([B + 'clust2snp', '-i', fa, '-n', str(rs.nreads1), '-x', '4', '-y', '4', '-z', '4'])
    assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr, r2.stderr)
    print('B200 CLIs  run %d: ebwt2clust %.2f s + clust2snp %.2f s = %.3g positions/s' % (rep, t1, t2, n / (t1 + t2)))
ours_cl = open(fa + '.clusters', 'rb').read(); ours_snp = open(os.path.join(d, 'ALL.snp'), 'rb').read()
t1, (r1, ncl) = (lambda t0, r: (time.perf_counter() - t0, r))(time.perf_counter(), O.ref_ebwt2clust(fa))
t0 = time.perf_counter(); r2, info = O.ref_clust2snp(fa, rs.nreads1); t2 = time.perf_counter() - t0
print('reference : ebwt2clust %.2f s + clust2snp %.2f s = %.3g positions/s' % (t1, t2, n / (t1 + t2)))
print('identical .clusters:', ours_cl == open(fa + '.clusters', 'rb').read(), ' identical .snp:', ours_snp == open(os.path.join(d, 'ALL.snp'), 'rb').read())
shutil.rmtree(d)
PY
