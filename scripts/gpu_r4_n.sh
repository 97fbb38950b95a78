#!/bin/bash
# session 4: smoke on the last build (comment-only change since the full suite)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r4n_smoke.log
