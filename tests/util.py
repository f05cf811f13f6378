"""Shared helpers for the parity tests (CUDA path through the C-ABI vs the CPU oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402
from oracle import oracle as O  # noqa: E402

RTOL = 1e-4  # BASELINE.json north_star: "within 1e-4 relative fp32"


def pkg():
    return load_package()


def relerr(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def relerr_robust(a, b, frac=2e-3):
    """relerr after discarding the `frac` largest element errors.

    The nets are piecewise linear: a pre-activation within rounding error of zero takes slope 1 on one
    side and 0.01 on the other (Caffe ReLU backward, `bottom_data > 0`), so two fp32 evaluations that
    differ only in summation order can legitimately disagree on a handful of (row, unit) masks, which
    moves that row's action gradient by percents.  Even the fp32 oracle differs from float64 autograd
    this way for unlucky seeds.  The bulk of each tensor must still agree to the parity tolerance and
    the number of outliers is bounded by `frac`."""
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    e = np.sort(np.abs(a - b)) / (np.abs(b).max() + 1e-30)
    keep = max(1, int(np.ceil(e.size * (1.0 - frac))))
    return float(e[keep - 1])


def make_pair(S, B, hidden, mode="warm", gemm_mode=0, n_replay=None, seed=0, capacity=None, p_term=0.2,
              use_graph=1, **cfg_kw):
    """An oracle state and a device learner holding identical weights and replay contents."""
    P = pkg()
    rng = np.random.default_rng(seed)
    use_blas = cfg_kw.pop("use_blas", 0)
    ocfg = O.make_config(state_size=S, batch=B, hidden=hidden, use_blas=use_blas, **cfg_kw)
    a0, c0 = O.init_params(ocfg, False, rng, mode), O.init_params(ocfg, True, rng, mode)
    if mode == "caffe":
        at, ct = a0.copy(), c0.copy()
    else:  # targets slightly off the online nets, as after some training
        at = (a0 + rng.normal(0, 1e-3, a0.size)).astype(np.float32)
        ct = (c0 + rng.normal(0, 1e-3, c0.size)).astype(np.float32)
    st = O.OracleState(ocfg, a0, c0, at, ct)
    n_replay = n_replay or max(4 * B, 256)
    capacity = capacity or (n_replay + 64)
    d = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=capacity, gemm_mode=gemm_mode,
               max_act_batch=max(32, min(B, 128)), use_graph=use_graph, **cfg_kw)
    d.set_params(P.ACTOR, a0); d.set_params(P.CRITIC, c0)
    d.set_params(P.ACTOR_TARGET, at); d.set_params(P.CRITIC_TARGET, ct)
    rcfg = O.make_config(state_size=S, batch=n_replay, hidden=hidden)
    replay = O.synth_batch(rcfg, rng, p_term=p_term)
    s, a, r, mc, term, sn = replay
    d.add_transitions(s, a, r, mc, sn, term)
    return st, d, replay, rng


def oracle_step(st, replay, idx, taps=False):
    s, a, r, mc, term, sn = replay
    return st.update(s[idx], a[idx], r[idx], mc[idx], term[idx], sn[idx], taps=taps)


def compare_state(st, d, lr_tol=0.02):
    """Post-update learner state: weights, Adam moments, target nets, iteration counters."""
    P = pkg()
    cfg = st.cfg
    out = {}
    out["critic"] = np.abs(d.get_params(P.CRITIC) - st.critic).max()
    out["actor"] = np.abs(d.get_params(P.ACTOR) - st.actor).max()
    out["critic_target"] = np.abs(d.get_params(P.CRITIC_TARGET) - st.critic_target).max()
    out["actor_target"] = np.abs(d.get_params(P.ACTOR_TARGET) - st.actor_target).max()
    am, av, ai = d.get_opt_state(P.ACTOR)
    cm, cv, ci = d.get_opt_state(P.CRITIC)
    out["actor_m"], out["actor_v"] = relerr(am, st.actor_m), relerr(av, st.actor_v)
    out["critic_m"], out["critic_v"] = relerr(cm, st.critic_m), relerr(cv, st.critic_v)
    out["iters"] = (ai, ci)
    return out
