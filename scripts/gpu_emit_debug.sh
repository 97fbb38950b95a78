cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E2S_EMIT_DEBUG=gpurun_out/emit_dbg.txt timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo rc=$?
python - <<'PY'
import numpy as np
a=np.loadtxt('gpurun_out/emit_dbg.txt',dtype=np.int64)
t0=a[:,1].min()
st=(a[:,1]-t0)/1e3; p1=(a[:,2]-t0)/1e3; ex=(a[:,3]-t0)/1e3; en=(a[:,4]-t0)/1e3
print('chunks',len(a))
for name,v in (('start',st),('pass1 done',p1),('exchange done',ex),('end',en)):
    print('%-14s min %.1f med %.1f p90 %.1f max %.1f us'%(name,v.min(),np.median(v),np.percentile(v,90),v.max()))
print('pass1 dur med %.1f max %.1f (chunk %d); exchange dur med %.1f max %.1f; pass2 dur med %.1f max %.1f'%(np.median(p1-st),(p1-st).max(),a[np.argmax(p1-st),0],np.median(ex-p1),(ex-p1).max(),np.median(en-ex),(en-ex).max()))
order=np.argsort(a[:,0]); 
for i in list(order[:8])+list(order[-4:]): print(int(a[i,0]), 'start %.1f p1 %.1f ex %.1f end %.1f'%(st[i],p1[i],ex[i],en[i]))
slow=np.argsort(-(p1-st))[:8]
print('slowest pass1:', [(int(a[i,0]), round(float(p1[i]-st[i]),1)) for i in slow])
PY
