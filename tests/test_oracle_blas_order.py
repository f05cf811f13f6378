"""The C oracle against a second, independent CPU restatement written as Caffe's BLAS call sequence
(tests/caffe_blas_model.py: numpy float32, OpenBLAS sgemm / sgemv, level-1 ops with Caffe's rounding points).
Two fp32 implementations that differ only in GEMM summation order must agree to a few ulps on everything computed
from the same weights, and to Adam's step ambiguity (|g| ~ eps -> +-lr) on the weights themselves."""
import numpy as np
import pytest

from oracle import oracle as O
import caffe_blas_model as M


def relerr(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("S,B,hidden,mode,n_updates", [
    (59, 32, (1024, 512, 256, 128), "caffe", 1),    # BASELINE cfg1: reference-native shape and init
    (59, 32, (1024, 512, 256, 128), "warm", 3),
    (77, 48, (96, 64, 48, 32), "warm", 3),          # 2v1 state size, ragged widths
    (58, 256, (256, 128, 64, 32), "warm", 2),
])
def test_oracle_matches_the_caffe_blas_call_sequence(S, B, hidden, mode, n_updates):
    cfg = O.make_config(state_size=S, batch=B, hidden=hidden)
    rng = np.random.default_rng(21)
    a0, c0 = O.init_params(cfg, False, rng, mode), O.init_params(cfg, True, rng, mode)
    at = (a0 + rng.normal(0, 1e-3, a0.size)).astype(np.float32) if mode != "caffe" else a0.copy()
    ct = (c0 + rng.normal(0, 1e-3, c0.size)).astype(np.float32) if mode != "caffe" else c0.copy()
    st = O.OracleState(cfg, a0, c0, at, ct)
    ms = dict(actor=a0.copy(), critic=c0.copy(), actor_target=at.copy(), critic_target=ct.copy(),
              actor_m=np.zeros_like(a0), actor_v=np.zeros_like(a0), critic_m=np.zeros_like(c0), critic_v=np.zeros_like(c0),
              actor_iter=0, critic_iter=0)
    blobs_a, blobs_c = O.net_blobs(cfg, False), O.net_blobs(cfg, True)
    for u in range(n_updates):
        batch = O.synth_batch(cfg, rng, p_term=0.25)
        loss, avgq = st.update(*batch, taps=True)
        mloss, mavgq, t = M.update(cfg, blobs_a, blobs_c, ms, *batch)
        o = st.last_taps
        tol = 1e-5 if u == 0 else 3e-4        # later updates start from weights that may differ by lr*sign(g) where |g| ~ eps
        assert abs(loss - mloss) <= tol * abs(mloss) + 1e-9, (u, loss, mloss)
        assert abs(avgq - mavgq) <= 10 * tol * abs(mavgq) + 1e-7, (u, avgq, mavgq)
        assert relerr(o["y"], t["y"]) < tol and relerr(o["q"], t["q"]) < tol
        assert relerr(o["a_pi"], t["a_pi"]) < tol
        if u == 0:
            assert relerr(o["critic_grad"], t["critic_grad"]) < 2e-5
            assert abs(o["critic_gnorm"][0] - t["critic_gnorm"]) <= 1e-5 * t["critic_gnorm"]
            # downstream of the critic's Adam step: bulk within 1e-4, a few rows may sit on a ReLU kink
            for k in ("q_pi", "d_raw", "d_inv", "actor_grad"):
                e = np.sort(np.abs(o[k].astype(np.float64).ravel() - np.asarray(t[k], np.float64).ravel())) / (np.abs(t[k]).max() + 1e-30)
                assert e[int(np.ceil(e.size * 0.998)) - 1] < 1e-4, (k, e[-1])
    assert (st.actor_iter, st.critic_iter) == (ms["actor_iter"], ms["critic_iter"]) == (n_updates, n_updates)
    assert np.abs(st.critic - ms["critic"]).max() <= 2.5 * cfg.critic_lr * n_updates
    assert np.abs(st.actor - ms["actor"]).max() <= 2.5 * cfg.actor_lr * n_updates
    assert relerr(st.critic_v, ms["critic_v"]) < 1e-3 and relerr(st.actor_v, ms["actor_v"]) < 1e-3
    assert np.abs(st.critic_target - ms["critic_target"]).max() <= 2.5 * cfg.critic_lr * cfg.tau * n_updates + 1e-7
