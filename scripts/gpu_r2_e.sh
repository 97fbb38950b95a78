#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or pipeline" > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/r2e_pytest.log
