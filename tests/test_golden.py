"""Known-answer tests against the committed fixtures in tests/golden/ (made by gen_golden.py).

CPU: the oracle must reproduce its own recorded outputs bit-for-bit (guards the oracle against
accidental edits and compiler/flag drift).  GPU: the CUDA path must match the same fixtures within
the 1e-4 parity tolerance."""
import importlib.util
import os

import numpy as np
import pytest

from util import ROOT, RTOL, pkg, relerr

spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(ROOT, "tests", "golden", "gen_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_golden_vectors(name):
    exp = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    res, _ = G.run_case(name)
    for k in exp.files:
        if k.endswith("digest"):
            np.testing.assert_allclose(res[k], exp[k], rtol=1e-12, err_msg=k)
        else:
            assert np.array_equal(res[k], exp[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("gemm_mode", [1, 0], ids=["simt_fp32", "tcgen05_3xtf32"])
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_path_matches_golden_vectors(name, gemm_mode):
    P = pkg()
    exp = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    cfg, (a0, c0, at, ct), replay, idx = G.make_case(name)
    S, B, hidden = cfg.state_size, cfg.batch, tuple(cfg.hidden[i] for i in range(cfg.n_hidden))
    d = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=replay[0].shape[0] + 16, gemm_mode=gemm_mode,
               soft_update_freq=cfg.soft_update_freq)
    d.set_params(P.ACTOR, a0); d.set_params(P.CRITIC, c0)
    d.set_params(P.ACTOR_TARGET, at); d.set_params(P.CRITIC_TARGET, ct)
    s, a, r, mc, term, sn = replay
    d.add_transitions(s, a, r, mc, sn, term)
    for u in range(idx.shape[0]):
        loss, avgq = d.update_with_indices(idx[u])
        assert abs(loss - exp["loss"][u]) <= 3 * RTOL * abs(exp["loss"][u]) + 1e-8, (u, loss, exp["loss"][u])
        assert abs(avgq - exp["avg_q"][u]) <= 3 * RTOL * abs(exp["avg_q"][u]) + 1e-6, (u, avgq, exp["avg_q"][u])
        if u == 0:
            assert relerr(d.debug_read("y", B), exp["y0"]) < 1e-6
            assert relerr(d.debug_read("q", B), exp["q0"]) < RTOL
            assert relerr(d.debug_read("a_pi", B * 10), exp["a_pi0"]) < RTOL
            assert relerr(d.debug_read("d_inv", B * 10), exp["d_inv0"]) < RTOL
            n_c, n_a = d.param_count(P.CRITIC), d.param_count(P.ACTOR)
            dg = G.digest(d.debug_read("critic_grad", n_c))
            assert np.abs(dg - exp["critic_grad0_digest"]).max() <= 2 * RTOL * np.abs(exp["critic_grad0_digest"]).max()
            dg = G.digest(d.debug_read("actor_grad", n_a))
            assert np.abs(dg - exp["actor_grad0_digest"]).max() <= 2 * RTOL * np.abs(exp["actor_grad0_digest"]).max()
    assert tuple(d.iters()) == tuple(exp["iters"])
    assert relerr(d.get_params(P.ACTOR)[-70:], exp["actor_head"]) < RTOL
    assert relerr(d.get_params(P.CRITIC)[-129:], exp["critic_head"]) < 5 * RTOL
    for net, key in ((P.ACTOR, "actor_digest"), (P.CRITIC, "critic_digest"), (P.ACTOR_TARGET, "actor_target_digest"),
                     (P.CRITIC_TARGET, "critic_target_digest")):
        dg = G.digest(d.get_params(net))
        assert np.abs(dg - exp[key])[2:].max() <= 2 * RTOL * np.abs(exp[key])[2:].max(), key
    d.close()
