"""Sharded paths on real GPUs: one process per GPU under torchrun (needs >= 2 GPUs)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_gpu_parity(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "scripts" / "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
