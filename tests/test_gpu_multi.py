"""Real NCCL data-parallel check (needs >= 2 GPUs; the driver's 1-GPU box skips it)."""
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_update_matches_oracle_and_replicas_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
