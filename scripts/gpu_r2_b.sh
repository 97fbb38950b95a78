#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2b_pytest.log
