"""Multi-GPU parity under pytest: spawns one rank per visible GPU (torch.distributed.run,
NCCL) running scripts/mgpu_check.py -- sharded index build incl. the irregular carry
exchange, gathered in rank order and compared with the oracle's unsharded index.
Skipped on boxes with fewer than two GPUs (the CPU/gloo twin is test_sharded_gloo.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_build_nccl_all_visible_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "scripts", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "mgpu_check PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
