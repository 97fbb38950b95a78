cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
E2S_PIPELINE_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_e2e_dbg.json 2> gpurun_out/bench_e2e_dbg.err; echo rc=$?
grep "e2s_pipeline_host" gpurun_out/bench_e2e_dbg.err | tail -24
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e_dbg.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.3f' % (d['value'], d['ms_per_step'])); print(d['e2e'])
PY
python - <<'PY'
# raw pinned H2D bandwidth of this box (torch), for reference
import torch, time
a=torch.empty(2<<30,dtype=torch.uint8,pin_memory=True); b=torch.empty_like(a,device='cuda')
for _ in range(2): b.copy_(a,non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); 
for _ in range(3): b.copy_(a,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
print('pinned H2D %.1f GB/s'%(a.numel()/dt/1e9))
PY
