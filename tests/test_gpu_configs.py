"""The other BASELINE.json configurations as parity cases against the oracle:
cfg3 shape (2v1: S=77, batch 4096) and the cfg5 wide-MLP variant (1024 hidden x4).

Two runs per shape.
 * critic_lr = actor_lr = 0: the weights do not move, so every intermediate of the update (both
   forward/backward passes, the dX chain through the critic, inverting gradients, both gradients)
   is a smooth function of the same weights in both implementations -> 1e-4 bound on everything
   (robust to the odd ReLU-kink flip, see util.relerr_robust).
 * reference learning rates: Adam's first step is ~lr*sign(g), so every weight whose gradient is below
   the implementations' rounding difference (|g| <~ 1e-5 max|g|) moves by +lr in one and -lr in the
   other.  That is a property of the algorithm, not of either implementation (the fp32 oracle shows the
   same against float64 autograd); it makes the post-update forward (q_pi, avg_q) agree only to ~1e-3 at
   these widths.  Checked with a correspondingly looser bound plus exact bookkeeping."""
import numpy as np
import pytest

from util import RTOL, compare_state, make_pair, oracle_step, relerr, relerr_robust
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SHAPES = {
    "cfg3_2v1_batch4096": (77, 4096, (1024, 512, 256, 128), 8192),
    "cfg5_wide_1024x4": (58, 1024, (1024, 1024, 1024, 1024), 4096),
}


@pytest.mark.parametrize("name", list(SHAPES))
def test_frozen_weights_every_intermediate_matches(name):
    S, B, hidden, n_replay = SHAPES[name]
    O.load_blas()
    st, d, replay, rng = make_pair(S, B, hidden, "warm", 0, n_replay=n_replay, critic_lr=0.0, actor_lr=0.0)
    idx = rng.integers(0, d.memory_size(), B).astype(np.int32)
    oloss, oavgq = oracle_step(st, replay, idx, taps=True)
    loss, avgq = d.update_with_indices(idx)
    t = st.last_taps
    assert abs(loss - oloss) <= RTOL * abs(oloss), (loss, oloss)
    assert abs(avgq - oavgq) <= RTOL * abs(oavgq) + 1e-6, (avgq, oavgq)
    for key, n in (("y", B), ("q", B), ("a_pi", B * 10), ("q_pi", B), ("critic_grad", t["critic_grad"].size)):
        assert relerr(d.debug_read(key, n), t[key]) < RTOL, key
    # action gradients: rows with a ReLU-kink flip differ by percents (expected share of rows ~ hidden units per row x
    # 1e-6, i.e. 0.15-0.35 % here; tests/test_oracle_autograd.py shows the fp32 oracle doing the same against float64)
    for key, n in (("d_raw", B * 10), ("d_inv", B * 10), ("actor_grad", t["actor_grad"].size)):
        got = d.debug_read(key, n)
        assert relerr_robust(got, t[key], frac=1e-2) < RTOL, (key, relerr_robust(got, t[key], frac=1e-2))
        assert relerr_robust(got, t[key], frac=0.5) < 0.1 * RTOL, (key, relerr_robust(got, t[key], frac=0.5))   # median
        assert relerr(got, t[key]) < 0.3, (key, relerr(got, t[key]))
    assert compare_state(st, d)["iters"] == (1, 1)
    d.close()


@pytest.mark.parametrize("name", list(SHAPES))
def test_reference_learning_rates(name):
    S, B, hidden, n_replay = SHAPES[name]
    O.load_blas()
    st, d, replay, rng = make_pair(S, B, hidden, "warm", 0, n_replay=n_replay)
    for u in range(2):
        idx = rng.integers(0, d.memory_size(), B).astype(np.int32)
        oloss, oavgq = oracle_step(st, replay, idx, taps=True)
        loss, avgq = d.update_with_indices(idx)
        assert abs(loss - oloss) <= (RTOL if u == 0 else 20 * RTOL) * abs(oloss), (u, loss, oloss)
        assert abs(avgq - oavgq) <= 30 * RTOL * abs(oavgq) + 1e-5, (u, avgq, oavgq)
        rows = np.abs(d.debug_read("q_pi", B) - st.last_taps["q_pi"]) / np.abs(st.last_taps["q_pi"]).max()
        assert np.median(rows) < 2 * RTOL and rows.max() < 30 * RTOL, (u, np.median(rows), rows.max())
    cs = compare_state(st, d)
    assert cs["iters"] == (2, 2)
    assert cs["critic"] <= 4.5 * st.cfg.critic_lr and cs["actor"] <= 4.5 * st.cfg.actor_lr, cs   # <= 2 sign flips
    assert cs["critic_target"] < 1e-5 and cs["actor_target"] < 1e-7, cs
    ms = d.benchmark(20)
    assert ms > 0
    d.close()
