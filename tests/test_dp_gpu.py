"""Data-parallel parity on >= 2 GPUs: fixed global batch, identical seeds, W ranks with the flat-gradient
all-reduce follow the 1-GPU loss trajectory within 1e-3 (tools/dp_parity.py; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_loss_trajectory_matches_single_gpu(cuda_lib):
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tools", "dp_parity.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "-> OK" in out.stdout
