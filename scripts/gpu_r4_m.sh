#!/bin/bash
# session 4: the whole GPU suite on the final tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r4m_pytest.log 2>&1; echo "pytest rc=$?"
tail -10 gpurun_out/r4m_pytest.log
