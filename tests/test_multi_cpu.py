"""World-size-2 gloo test (CPU) of the data-parallel formulation used by the multi-GPU path
(DESIGN.md §5): replay rows sharded by rank, each rank contributes the gradient of its local minibatch
scaled by 1/(B_local * world), one all-reduce(sum) per net, and every replica then holds the gradient of
the single-learner update at the global batch.  The GPU counterpart (real NCCL, whole update, bit-identical
replicas) is scripts/multi_gpu_check.py / tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest

from util import ROOT


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, B, hidden = 58, 32, (64, 48, 32, 16)
    rng = np.random.default_rng(0)                      # identical stream on every rank
    gcfg = O.make_config(state_size=S, batch=B * world, hidden=hidden, critic_lr=0.0, actor_lr=0.0)
    lcfg = O.make_config(state_size=S, batch=B, hidden=hidden, critic_lr=0.0, actor_lr=0.0)
    a0, c0 = O.init_params(gcfg, False, rng, "warm"), O.init_params(gcfg, True, rng, "warm")
    shards = [O.synth_batch(lcfg, rng) for _ in range(world)]
    # local contribution: oracle on the local shard (its critic gradient carries 1/B_local) ...
    st = O.OracleState(lcfg, a0, c0, a0, c0)
    lloss, lq = st.update(*shards[rank], taps=True)
    gc = torch.from_numpy(st.last_taps["critic_grad"] / world)      # ... rescaled to 1/(B_local*world)
    ga = torch.from_numpy(st.last_taps["actor_grad"].copy())        # actor gradient is a plain sum over rows
    scal = torch.tensor([lloss / world, lq / world], dtype=torch.float64)
    for t in (gc, ga, scal):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    # single learner at the global batch (weights frozen so that both phases see the same nets)
    ref = O.OracleState(gcfg, a0, c0, a0, c0)
    cat = [np.concatenate([shards[w][k] for w in range(world)]) for k in range(6)]
    gloss, gq = ref.update(*cat, taps=True)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    res = dict(critic=rel(gc.numpy(), ref.last_taps["critic_grad"]), actor=rel(ga.numpy(), ref.last_taps["actor_grad"]),
               loss=abs(scal[0].item() - gloss) / abs(gloss), avg_q=abs(scal[1].item() - gq) / abs(gq))
    # replicas must agree bit for bit after the all-reduce
    gathered = [torch.empty_like(gc) for _ in range(world)]
    dist.all_gather(gathered, gc)
    res["identical"] = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        import json
        json.dump(res, open(out_path, "w"))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_sharding_matches_global_batch(tmp_path):
    import json
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.json")
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res["identical"]
    assert res["critic"] < 1e-5 and res["actor"] < 1e-5, res
    assert res["loss"] < 1e-5 and res["avg_q"] < 1e-5, res
