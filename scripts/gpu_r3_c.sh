set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_streaming_gpu.py tests/test_cli_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r3c_pytest.log
