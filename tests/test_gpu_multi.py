"""Real multi-GPU data-parallel checks (scripts/multi_gpu_check.py under torchrun; needs >= 2 GPUs, the 1-GPU box
skips them - bench.py --gpus N carries the same check in its "parity" block for the driver's scaling runs), plus the
single-GPU run of the same checker, which always executes."""
import os
import subprocess
import sys

import pytest

from util import ROOT, pkg

pytestmark = pytest.mark.gpu


def _torchrun(n, cases, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "multi_gpu_check.py"), cases]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=1500)


def test_two_rank_update_matches_oracle_and_replicas_identical():
    """small shape (P2P kernel and NCCL) + BASELINE cfg3 as specified: S=77, global batch 4096 on 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = _torchrun(2, "small,cfg3", 29511)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_eight_rank_wide_mlp_matches_oracle_and_replicas_identical():
    """BASELINE cfg5 as specified: 1024x4 towers, global batch 8192 on 8 GPUs (+ cfg2 at 8 x 1024)."""
    import torch
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs (gpurun --gpus 8)")
    out = _torchrun(8, "cfg2,cfg5", 29512)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_parity_checker_single_rank():
    """The checker itself at world size 1 (no communicator): the path bench.py's parity block takes on one GPU."""
    import torch
    sys.path.insert(0, ROOT)
    from scripts import dp_parity
    res = dp_parity.check(pkg(), None, torch, 0, 1, 0, 58, 256, (256, 128, 64, 64), n_updates=2, frozen=True)
    assert res["ok"], res
    res = dp_parity.check(pkg(), None, torch, 0, 1, 0, 58, 256, (256, 128, 64, 64), n_updates=2, frozen=False)
    assert res["ok"], res


def _cfg4(n, port):
    script = os.path.join(ROOT, "scripts", "cfg4_rollout.py")
    if n == 1:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
               "--master-addr", "127.0.0.1", "--master-port", str(port), script]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900)


def test_cfg4_rollout_single_gpu():
    """BASELINE cfg4 on one GPU: 64 workers, act path beside asynchronous updates, replay appends on the copy stream."""
    out = _cfg4(1, 0)
    assert "CFG4_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_cfg4_rollout_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = _cfg4(2, 29541)
    assert "CFG4_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_cfg4_rollout_eight_ranks():
    """BASELINE cfg4 as specified: 64 workers -> 8-GPU sharded replay, 8 workers per rank."""
    import torch
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs (gpurun --gpus 8)")
    out = _cfg4(8, 29542)
    assert "CFG4_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
