"""BASELINE cfg4: 64 parallel actor workers -> replay sharded over the ranks, batched SelectActions on the act path
beside asynchronous data-parallel learner updates.  One process per GPU:

  python scripts/cfg4_rollout.py                                   (1 GPU: all 64 workers on one rank)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      scripts/cfg4_rollout.py                                      (8 GPUs: 8 workers per rank)

Per environment step every rank (reference loop: dqn_main.cpp:97-153 PlayOneEpisode, :352-366 KeepPlayingGames)
  1. serves its W/world workers with ONE dqnb_select_actions call (act stream + actor snapshot, never waits for the
     learner) while the updates enqueued in the previous step are still running,
  2. steps the workers (synthetic stand-in for HFO, see ToyWorkers), labels finished episodes (LabelTransitions,
     dqn.cpp:783-797) and appends them to ITS shard of the replay (dqnb_add_transitions on the copy stream),
  3. enqueues the step's updates (dqnb_update_async): every rank the same number, because each update exchanges
     gradients with all ranks; results are collected one step late (dqnb_results).
Checks (exit code 0 and the line CFG4_OK): on every 8th greedy step the learner is drained first and the action batch
must equal the oracle's actor forward on the newest actor (tests/test_gpu_act.py covers the torn-snapshot case while
updates are in flight), replicas bit-identical at the end, shard sizes add up, losses finite and
falling, no exchange timeout.  Prints act latency beside running updates and the env-steps / updates rate.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)


class ToyWorkers:
    """W independent 1-D 'move to the ball' problems with an S-dim observation (as tests/test_gpu_rollout.py)."""

    def __init__(self, W, S, rng):
        self.W, self.S, self.rng = W, S, rng
        self.pos = rng.uniform(-1, 1, W).astype(np.float32)
        self.ball = rng.uniform(-1, 1, W).astype(np.float32)
        self.t = np.zeros(W, np.int32)

    def obs(self):
        o = np.zeros((self.W, self.S), np.float32)
        o[:, 0], o[:, 1], o[:, 2] = self.pos, self.ball, self.ball - self.pos
        o[:, 3:] = np.sin(np.outer(self.ball - self.pos, np.arange(3, self.S)))
        return o

    def step(self, act10):
        move = np.clip(act10[:, 4] / 100.0, -1, 1) * 0.2
        before = np.abs(self.ball - self.pos)
        self.pos = np.clip(self.pos + move, -1.5, 1.5).astype(np.float32)
        after = np.abs(self.ball - self.pos)
        self.t += 1
        done = (after < 0.05) | (self.t >= 20)
        reward = (before - after + np.where(after < 0.05, 1.0, 0.0)).astype(np.float32)
        return reward, done

    def reset(self, mask):
        n = int(mask.sum())
        self.pos[mask] = self.rng.uniform(-1, 1, n)
        self.ball[mask] = self.rng.uniform(-1, 1, n)
        self.t[mask] = 0


def random_actions(rng, n):
    a = np.empty((n, 10), np.float32)                      # GetRandomActorOutput ranges, dqn.cpp:664-682
    a[:, :4] = rng.uniform(-1, 1, (n, 4)); a[:, 4] = rng.uniform(-100, 100, n)
    a[:, 5:8] = rng.uniform(-180, 180, (n, 3)); a[:, 8] = rng.uniform(0, 100, n); a[:, 9] = rng.uniform(-180, 180, n)
    return a


def main(total_workers=64, env_steps=160, S=58, B=256, hidden=(256, 128, 64, 64), updates_per_step=2, warm_rows=512):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = load_package()
    from scripts import dp_parity
    assert total_workers % world == 0
    W = total_workers // world
    rng = np.random.default_rng(100 + rank)
    d = P.DQNB(device=local, state_size=S, batch=B, hidden=hidden, replay_capacity=50000, seed=7 + rank,
               world_size=world, rank=rank, max_act_batch=max(W, 8), critic_lr=1e-3, actor_lr=1e-4)
    if world > 1:
        dp_parity.connect(P, d, "p2p", rank, world, dist, torch)
    d.init_params(seed=5, std=0.05)                          # identical replicas
    env = ToyWorkers(W, S, rng)
    # every shard starts with some random-policy experience so that updates can run from the first step
    warm = O.synth_batch(O.make_config(state_size=S, batch=warm_rows, hidden=hidden), rng, p_term=0.1)
    d.add_transitions(warm[0], warm[1], warm[2], warm[3], warm[5], warm[4])
    appended = warm_rows
    if world > 1:                                            # start together: the first exchange should not have to wait for
        torch.cuda.synchronize(); dist.barrier()             # a rank that is still initialising
    ocfg = O.make_config(state_size=S, batch=B, hidden=hidden)
    episodes = [[] for _ in range(W)]
    pending = 0                                              # sequence number of the last enqueued update
    collected = 0
    losses, act_us, torn = [], [], 0
    checked = 0
    t_loop = time.perf_counter()
    for step in range(env_steps):
        obs = env.obs()
        greedy = rng.uniform() >= 0.3                        # one coin flip per batch (dqn.cpp:700)
        check = greedy and step % 8 == 0
        if check and pending > collected:
            # checked steps: the learner is drained first, so the batch must come from exactly the newest actor
            l, _ = d.results(collected + 1, pending - collected)
            losses += list(l)
            collected = pending
            d.sync()
        t0 = time.perf_counter()
        act = d.select_actions(obs) if greedy else random_actions(rng, W)
        if greedy and not check:
            act_us.append((time.perf_counter() - t0) * 1e6)   # served while this step's predecessors are still updating
        # collect the previous step's updates (they ran under the act call and the env step)
        if pending > collected:
            l, _ = d.results(collected + 1, pending - collected)
            losses += list(l)
            collected = pending
        if check:
            cur = d.get_params(P.ACTOR)
            st = O.OracleState(ocfg, cur, d.get_params(P.CRITIC), cur, d.get_params(P.CRITIC))
            ref = st.actor_forward(obs)
            checked += 1
            if np.abs(act - ref).max() / (np.abs(ref).max() + 1e-9) > 1e-4:
                torn += 1
        reward, done = env.step(act)
        nxt = env.obs()
        for w in range(W):
            episodes[w].append((obs[w], act[w], reward[w], nxt[w], bool(done[w])))
            if done[w]:
                ep = episodes[w]
                r = np.array([e[2] for e in ep], np.float32)
                mc = O.label_transitions(r, 0.99)
                s = np.stack([e[0] for e in ep]); a = np.stack([e[1] for e in ep]); sn = np.stack([e[3] for e in ep])
                term = np.array([e[4] for e in ep], np.uint8)
                d.add_transitions(s, a, r, mc, sn, term)
                appended += len(ep)
                episodes[w] = []
        env.reset(done)
        pending = d.update_async(updates_per_step)           # same count on every rank: each update exchanges gradients
    l, _ = d.results(collected + 1, pending - collected)
    losses += list(l)
    d.sync()
    wall = time.perf_counter() - t_loop
    ok = True
    notes = {}
    if d.memory_size() != appended:
        ok = False; notes["shard_size"] = (d.memory_size(), appended)
    if not np.isfinite(losses).all() or len(losses) != env_steps * updates_per_step:
        ok = False; notes["losses"] = "non-finite or missing"
    if torn:
        ok = False; notes["action_batches_wrong"] = torn
    if world > 1 and d.comm_status() != 0:
        ok = False; notes["exchange"] = "timeout"
    identical = True
    total_rows = appended
    if world > 1:
        for net in (P.ACTOR, P.CRITIC, P.ACTOR_TARGET, P.CRITIC_TARGET):
            allp = dp_parity._gather_np(d.get_params(net), world, dist, torch)
            identical &= all(np.array_equal(allp[0], p) for p in allp)
        t = torch.tensor([appended], device="cuda"); dist.all_reduce(t); total_rows = int(t.item())
        if not identical:
            ok = False; notes["replicas"] = "differ"
    ai, ci = d.iters()
    if (ai, ci) != (len(losses), len(losses)):
        ok = False; notes["iters"] = (ai, ci, len(losses))
    res = {"ok": ok, "world": world, "workers_total": total_workers, "workers_per_rank": W, "env_steps": env_steps,
           "updates": len(losses), "global_batch": B * world, "replay_rows_total": total_rows, "replicas_bit_identical": bool(identical),
           "action_batches_checked": checked, "action_batches_wrong": torn,
           "act_us_median_beside_updates": round(float(np.median(act_us)), 1) if act_us else None,
           "act_us_p90": round(float(np.percentile(act_us, 90)), 1) if act_us else None,
           "loss_first_last": [float(np.mean(losses[:20])), float(np.mean(losses[-20:]))],
           "env_steps_per_s_all_workers": round(env_steps * total_workers / wall, 1), "notes": notes}
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() == 1)
        res["ok"] = ok
    d.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
        print("CFG4_OK" if ok else "CFG4_FAIL", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
