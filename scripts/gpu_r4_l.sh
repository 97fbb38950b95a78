#!/bin/bash
# session 4: ncu --set full of the phase-2 / exchange kernels of one C3 step (source-level), to look for waste like k_chunk_resolve's
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:"k_merge_stats|k_cluster_exact|k_candidates|k_events|k_chunk_resolve" -s 15 -c 5 -o gpurun_out/r4l_tail python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r4l_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r4l_ncu.log
