"""cfg4-style integration on one GPU: 64 parallel synthetic actor workers drive batched SelectActions,
their episodes are labelled (LabelTransitions rule) and appended to the HBM ring, and the learner runs
update_ratio updates per environment step (dqn_main.cpp:352-363).  Checks the pieces against the oracle
where they meet: batched greedy actions == oracle actor forward on the current weights after learning,
ring contents == what was appended, iteration bookkeeping, and that the critic actually fits the
(deterministic) returns of this toy task."""
import numpy as np
import pytest

from util import RTOL, pkg, relerr
from oracle import oracle as O

pytestmark = pytest.mark.gpu


class ToyWorkers:
    """W independent 1-D 'move to the ball' problems with a 58-dim observation."""

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
        move = np.clip(act10[:, 4] / 100.0, -1, 1) * 0.2          # dash power parameter drives the agent
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


def test_64_workers_rollout_and_learning_loop():
    P = pkg()
    W, S, B, hidden = 64, 58, 256, (128, 64, 64, 32)
    rng = np.random.default_rng(0)
    d = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=20000, max_act_batch=W, critic_lr=1e-3, actor_lr=1e-4)
    d.init_params(seed=5, std=0.05)
    env = ToyWorkers(W, S, rng)
    episodes = [[] for _ in range(W)]
    appended = []
    losses = []
    eps = 0.5
    for step in range(120):
        obs = env.obs()
        # SelectActions semantics: one coin flip per batch (dqn.cpp:700)
        if rng.uniform() < eps:
            act = np.empty((W, 10), np.float32)
            act[:, :4] = rng.uniform(-1, 1, (W, 4)); act[:, 4] = rng.uniform(-100, 100, W)
            act[:, 5:8] = rng.uniform(-180, 180, (W, 3)); act[:, 8] = rng.uniform(0, 100, W); act[:, 9] = rng.uniform(-180, 180, W)
        else:
            act = d.select_actions(obs)
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
                appended.append((s, a, r, mc, term))
                episodes[w] = []
        env.reset(done)
        if d.memory_size() >= 1000:
            l, _ = d.update(6)                      # update_ratio 0.1 x 64 env steps
            losses += list(l)
    assert d.memory_size() == sum(len(x[2]) for x in appended)
    # ring contents == everything appended, in order
    s_all = np.concatenate([x[0] for x in appended]); r_all = np.concatenate([x[2] for x in appended])
    mc_all = np.concatenate([x[3] for x in appended]); t_all = np.concatenate([x[4] for x in appended])
    gs, ga, gr, gmc, gsn, gt = d.get_transitions(0, d.memory_size())
    assert np.array_equal(gs, s_all) and np.array_equal(gr, r_all) and np.array_equal(gmc, mc_all) and np.array_equal(gt, t_all)
    assert d.iters() == (len(losses), len(losses)) and len(losses) > 100
    assert np.isfinite(losses).all()
    assert np.mean(losses[-50:]) < 0.7 * np.mean(losses[:50]), (np.mean(losses[:50]), np.mean(losses[-50:]))
    # the act path serves the weights the learner just produced
    ocfg = O.make_config(state_size=S, batch=B, hidden=hidden)
    st = O.OracleState(ocfg, d.get_params(P.ACTOR), d.get_params(P.CRITIC), d.get_params(P.ACTOR_TARGET), d.get_params(P.CRITIC_TARGET))
    obs = env.obs()
    assert relerr(d.select_actions(obs), st.actor_forward(obs)) < RTOL
    assert relerr(d.evaluate(obs, act), st.critic_forward(obs, act)) < RTOL
    d.close()
