# quick GPU check + ncu --set full of the kernels named in $KREGEX (after the plain run exited 0)
cd $GRAFT_REPO_ROOT
bash scripts/gpu_quick.sh || exit 1
grep -q "pytest rc=0" gpurun_out/pytest_gpu.log || exit 1
KREGEX=${KREGEX:-k_cluster_emit}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-1} -f -o gpurun_out/prof_${TAG:-tmp} python bench.py --steps 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG:-tmp}.log 2>&1; echo "ncu rc=$?"
