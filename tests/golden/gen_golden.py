"""Generates the known-answer fixtures under tests/golden/ from the CPU oracle.

The reference (mhauskn/dqn-hfo) has no tests or recorded outputs and its Caffe dependency cannot
be built here, so these vectors pin the ORACLE (validated against float64 autograd in
tests/test_oracle_autograd.py), not the reference binary: PARITY UNPINNED.
Inputs are regenerated from seeds; only the expected outputs (a few KB) are stored.

    python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

CASES = {  # name: (S, B, hidden, weight mode, n_updates, p_term, extra cfg)
    "cfg1_s59_b32_caffe": (59, 32, (1024, 512, 256, 128), "caffe", 3, 0.2, {}),
    "cfg1_s59_b32_warm": (59, 32, (1024, 512, 256, 128), "warm", 3, 0.2, {}),
    "s77_b48_small_softfreq2": (77, 48, (96, 64, 48, 32), "warm", 4, 0.3, {"soft_update_freq": 2}),
    "s58_b64_all_terminal": (58, 64, (128, 64, 64, 32), "warm", 2, 1.0, {}),
}


def make_case(name):
    S, B, hidden, mode, n_up, p_term, extra = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    cfg = O.make_config(state_size=S, batch=B, hidden=hidden, **extra)
    a0, c0 = O.init_params(cfg, False, rng, mode), O.init_params(cfg, True, rng, mode)
    at = (a0 + rng.normal(0, 1e-3, a0.size)).astype(np.float32)
    ct = (c0 + rng.normal(0, 1e-3, c0.size)).astype(np.float32)
    n_replay = 4 * B
    replay = O.synth_batch(O.make_config(state_size=S, batch=n_replay, hidden=hidden), rng, p_term=p_term)
    idx = rng.integers(0, n_replay, (n_up, B)).astype(np.int32)
    return cfg, (a0, c0, at, ct), replay, idx


def digest(x):
    """Position-weighted float64 checksums: sensitive to any element moving or changing."""
    x = np.asarray(x, np.float64).ravel()
    w = np.cos(np.arange(x.size) * 0.37) + 1.5
    return np.array([x.sum(), (x * w).sum(), np.abs(x).sum(), (x * x).sum()])


def run_case(name):
    cfg, (a0, c0, at, ct), replay, idx = make_case(name)
    st = O.OracleState(cfg, a0, c0, at, ct)
    s, a, r, mc, term, sn = replay
    out = {"loss": [], "avg_q": [], "y0": None}
    for u in range(idx.shape[0]):
        i = idx[u]
        loss, avgq = st.update(s[i], a[i], r[i], mc[i], term[i], sn[i], taps=(u == 0))
        out["loss"].append(loss); out["avg_q"].append(avgq)
        if u == 0:
            out["y0"] = st.last_taps["y"].copy()
            out["q0"] = st.last_taps["q"].copy()
            out["a_pi0"] = st.last_taps["a_pi"].copy()
            out["d_inv0"] = st.last_taps["d_inv"].copy()
            out["critic_grad0_digest"] = digest(st.last_taps["critic_grad"])
            out["actor_grad0_digest"] = digest(st.last_taps["actor_grad"])
    res = dict(loss=np.array(out["loss"], np.float32), avg_q=np.array(out["avg_q"], np.float32),
               y0=out["y0"], q0=out["q0"], a_pi0=out["a_pi0"], d_inv0=out["d_inv0"],
               critic_grad0_digest=out["critic_grad0_digest"], actor_grad0_digest=out["actor_grad0_digest"],
               actor_digest=digest(st.actor), critic_digest=digest(st.critic),
               actor_target_digest=digest(st.actor_target), critic_target_digest=digest(st.critic_target),
               critic_m_digest=digest(st.critic_m), critic_v_digest=digest(st.critic_v),
               actor_head=st.actor[-70:].copy(), critic_head=st.critic[-129:].copy(),
               iters=np.array([st.actor_iter, st.critic_iter], np.int32))
    return res, st


if __name__ == "__main__":
    for name in CASES:
        res, _ = run_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, "loss", res["loss"], "avg_q", res["avg_q"])
