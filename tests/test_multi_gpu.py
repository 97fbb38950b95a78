"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise): one process per GPU over NCCL, sharded run == oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_or_more_ranks_nccl(built):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for world in sorted({2, min(n, 4)}):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(29530 + world),
                            os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert r.stdout.count("OK ") == 2, r.stdout
