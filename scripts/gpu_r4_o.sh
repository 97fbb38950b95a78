#!/bin/bash
# session 4: build_gesa after its record packing went multi-threaded
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_cli_gpu.py tests/test_builder_gpu.py -m gpu -q -k "build_gesa or tool_chain" > gpurun_out/r4o_build_gesa.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4o_build_gesa.log
