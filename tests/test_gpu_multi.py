"""Multi-GPU parity on real devices (NCCL): launches tests/dist/gpu_sharded_check.py under torchrun when the
box has >= 2 GPUs (skipped on a 1-GPU box; the host-side protocol is covered on CPU by test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_bag_and_cohort_on_two_gpus():
    port = 29600 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist", "gpu_sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded bag ok: big" in r.stdout and "cohort gather + cox ok" in r.stdout
    assert "peer all-reduce ok" in r.stdout
