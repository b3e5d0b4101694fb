"""GPU, needs >= 2 devices (skipped otherwise): fused NVLink result exchange == NCCL all-gather."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_result_exchange_matches_allgather():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 8)       # all GPUs of the box (2, 4 or 8 ranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_fused_exchange.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert f"FUSED_EXCHANGE_OK world {n}" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
