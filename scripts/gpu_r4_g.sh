#!/bin/bash
# session 4: key-range builds
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_builder_gpu.py tests/test_build_egsa.py -m gpu -q --durations=3 > gpurun_out/r4g_builder.log 2>&1; echo "builder rc=$?"
tail -30 gpurun_out/r4g_builder.log
