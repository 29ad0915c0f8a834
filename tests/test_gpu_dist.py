"""GPU, >= 2 devices: the sharded step over NCCL reproduces the single-process global-batch oracle
(tests/dist_check.py under torchrun).  Skipped on single-GPU boxes; the CPU/gloo test covers the exchange logic."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_step_matches_global_batch_oracle_over_nccl():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + os.getpid() % 200), os.path.join(ROOT, "tests", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "symmetric=True" in out.stdout
