"""Data-parallel parity check shared by bench.py (--gpus N > 1), scripts/multi_gpu_check.py and the tests.

One learner per rank (replay sharded by rank, B transitions per rank per update) must equal ONE learner at the
global batch B * world: the CPU oracle (test infrastructure, used here only as the checker) is run on rank 0 on
the concatenation of every rank's injected minibatch, and every tap / post-update parameter of the replicas is
compared with it; the replicas themselves must stay bit-identical (the exchange kernel reduces every element on
exactly one rank in a fixed order).

Tolerance: 1e-4 relative (max|a-b| / max|b| per tensor) on every intermediate when the weights are frozen
(lr = 0); with the reference learning rates the post-update quantities carry Adam's lr*sign(g) ambiguity for
|g| ~ eps, so parameters are bounded by a multiple of the step size (tests/util.py compare_state).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RTOL = 1e-4


def connect(P, d, comm, rank, world, dist, torch):
    """Wires the replicas together: 'p2p' = the IPC/NVLink exchange kernel (product path), 'nccl' = ncclAllReduce."""
    if comm == "nccl":
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(P.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        d.comm_init(bytes(idt.cpu().numpy().tobytes()))
    else:
        mine = torch.frombuffer(bytearray(d.comm_p2p_handle()), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        d.comm_p2p_init(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))


def _gather_np(x, world, dist, torch):
    """all-gather of equally shaped numpy arrays -> list (rank order), through device tensors (NCCL group)."""
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def relerr(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def check(P, dist, torch, rank, world, local, S, B, hidden, comm="p2p", n_updates=2, frozen=False, n_replay=None,
          seed=0):
    """Returns a JSON-able dict; dict['ok'] is the verdict (identical on every rank)."""
    from oracle import oracle as O
    t_start = time.perf_counter()
    lr = dict(critic_lr=0.0, actor_lr=0.0) if frozen else {}
    rng = np.random.default_rng(seed)                      # same stream on every rank
    ocfg = O.make_config(state_size=S, batch=B * world, hidden=hidden, **lr)
    a0, c0 = O.init_params(ocfg, False, rng, "warm"), O.init_params(ocfg, True, rng, "warm")
    at = (a0 + rng.normal(0, 1e-3, a0.size)).astype(np.float32)
    ct = (c0 + rng.normal(0, 1e-3, c0.size)).astype(np.float32)
    n = n_replay or max(2 * B, 1024)
    shard_rng = np.random.default_rng(1000 + seed + rank)  # a different shard of the replay per rank
    shard = O.synth_batch(O.make_config(state_size=S, batch=n, hidden=hidden), shard_rng, p_term=0.2)
    idx = shard_rng.integers(0, n, (n_updates, B)).astype(np.int32)
    d = P.DQNB(device=local, state_size=S, batch=B, hidden=hidden, replay_capacity=n + 64, world_size=world, rank=rank,
               max_act_batch=32, **lr)
    if world > 1:
        connect(P, d, comm, rank, world, dist, torch)
    d.set_params(P.ACTOR, a0); d.set_params(P.CRITIC, c0)
    d.set_params(P.ACTOR_TARGET, at); d.set_params(P.CRITIC_TARGET, ct)
    s, a, r, mc, term, sn = shard
    d.add_transitions(s, a, r, mc, sn, term)
    st = None
    if rank == 0:
        O.load_blas()
        st = O.OracleState(ocfg, a0, c0, at, ct)
    worst = {}
    ok = True
    tap_keys = (("y", B), ("q", B), ("q_pi", B), ("a_pi", B * 10), ("d_raw", B * 10), ("d_inv", B * 10))

    def note(key, err, bound):
        nonlocal ok
        worst[key] = max(worst.get(key, 0.0), float(err))
        if not err <= bound:
            ok = False

    for u in range(n_updates):
        if world > 1:
            # rank 0 spends seconds in the oracle between updates (5 s at the wide shapes): the others wait for it HERE,
            # on the host, not inside the exchange kernel, whose timeout is for peers that are gone
            torch.cuda.synchronize()
            dist.barrier()
        loss, avgq = d.update_with_indices(idx[u])
        mine = [np.ascontiguousarray(shard[k][idx[u]]) for k in range(6)]
        if world > 1:
            cat = [np.concatenate(_gather_np(m.astype(np.float32) if m.dtype != np.float32 else m, world, dist, torch))
                   for m in mine]
            cat[4] = cat[4].astype(np.uint8)
            taps = {k: np.concatenate(_gather_np(d.debug_read(k, cnt), world, dist, torch)) for k, cnt in tap_keys}
        else:
            cat = mine
            taps = {k: d.debug_read(k, cnt) for k, cnt in tap_keys}
        if rank == 0:
            oloss, oavgq = st.update(cat[0], cat[1], cat[2], cat[3], cat[4], cat[5], taps=True)
            t = st.last_taps
            strict = frozen or u == 0                 # taps of the first update come from identical weights
            note("critic_loss", abs(loss - oloss) / abs(oloss), RTOL if strict else 20 * RTOL)
            note("avg_q", abs(avgq - oavgq) / (abs(oavgq) + 1e-6), RTOL if frozen else 30 * RTOL)
            for k in ("y", "q"):
                note(k, relerr(taps[k], t[k]), RTOL if strict else 30 * RTOL)
            # everything after the critic's Adam step sees weights that may differ by lr*sign(g) where |g| ~ eps
            for k in ("a_pi",):
                note(k, relerr(taps[k], t[k]), RTOL if strict else 30 * RTOL)
            if frozen:
                # ReLU-kink rows (tests/util.py relerr_robust, tests/test_oracle_autograd.py): a row whose critic has a
                # pre-activation within rounding (~1e-6 relative for 3xTF32) of zero takes slope 1 here and 0.01 there,
                # which moves that row's action gradient by percents.  Expected share of such rows ~ hidden units per
                # row x 1e-6: 0.15 % at 1024-512-256-128, 0.35 % at 1024x4 -> the bound is on the 99th percentile.
                for k in ("q_pi", "d_raw", "d_inv"):
                    e = np.sort(np.abs(np.asarray(taps[k], np.float64).ravel() - np.asarray(t[k], np.float64).ravel())) \
                        / (np.abs(t[k]).max() + 1e-30)
                    note(k + "_p99", e[int(np.ceil(e.size * 0.99)) - 1], RTOL)
                    note(k + "_median", e[e.size // 2], 0.1 * RTOL)
                    note(k + "_max", e[-1], 0.3)
                for k in ("critic_grad", "actor_grad"):
                    got = d.debug_read(k, t[k].size)
                    e = np.sort(np.abs(got.astype(np.float64) - t[k].astype(np.float64))) / (np.abs(t[k]).max() + 1e-30)
                    note(k + "_p998", e[int(np.ceil(e.size * 0.998)) - 1], RTOL)
                    note(k + "_max", e[-1], 0.3)
    # post-update parameters: oracle on rank 0, bit identity across replicas everywhere
    identical = True
    for net, name, lrv in ((P.CRITIC, "critic", ocfg.critic_lr), (P.ACTOR, "actor", ocfg.actor_lr),
                           (P.CRITIC_TARGET, "critic_target", ocfg.critic_lr * ocfg.tau),
                           (P.ACTOR_TARGET, "actor_target", ocfg.actor_lr * ocfg.tau)):
        got = d.get_params(net)
        if world > 1:
            allp = _gather_np(got, world, dist, torch)
            identical &= all(np.array_equal(allp[0], p) for p in allp)
        if rank == 0:
            ref = getattr(st, name)
            note("param_" + name, np.abs(got - ref).max(), 2.5 * lrv * n_updates + 1e-7)
    if not identical:
        ok = False
    if world > 1 and d.comm_status() != 0:
        ok = False
    d.close()
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() == 1)
    return {"ok": ok, "world": world, "comm": comm if world > 1 else None, "state_size": S, "batch_per_gpu": B,
            "global_batch": B * world, "hidden": list(hidden), "updates": n_updates, "frozen_weights": frozen,
            "replicas_bit_identical": bool(identical), "worst_relerr": {k: float(f"{v:.3e}") for k, v in worst.items()},
            "tolerance": RTOL, "seconds": round(time.perf_counter() - t_start, 2),
            "checker": "CPU oracle (oracle/dqn_oracle.c) at the global batch, rank 0"}
