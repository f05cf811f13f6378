"""Multi-agent parameter sharing (reference dqn.cpp:1037-1083, dqn_main.cpp:305-323).

Upstream `ShareParameters(other, na, nc)` makes the first na / nc layers-with-parameters of the actor / critic (and of
their target nets) ONE set of blobs that both agents' solvers update, each with its own Adam history.  Here the group
writes the updating member's shared layers through to the others (dqnb_copy_shared_layers).  The oracle side of the test
models Blob::ShareData directly: two oracle states whose shared blob ranges (Caffe learnable_params order, computed
here independently of the library's padded layout) are copied after each update."""
import numpy as np
import pytest

from util import O, RTOL, make_pair, oracle_step, pkg, relerr

pytestmark = pytest.mark.gpu


def caffe_ranges(S, hidden, critic, n_layers):
    """[begin, end) ranges of the first n layers-with-parameters in Caffe blob order (W then b per layer)."""
    dims = [S + (10 if critic else 0)] + list(hidden)
    off, out = 0, []
    layers = [(dims[i + 1], dims[i]) for i in range(len(hidden))]
    layers += [(1, hidden[-1])] if critic else [(4, hidden[-1]), (6, hidden[-1])]
    for i, (n, k) in enumerate(layers):
        size = n * k + n
        if i < n_layers:
            out.append((off, off + size))
        off += size
    return out, off


@pytest.mark.parametrize("na,nc", [(2, 3), (5, 1), (6, 5), (0, 2)])
def test_shared_layers_follow_the_updating_agent(na, nc):
    P = pkg()
    S, B, hidden = 58, 64, (128, 64, 64, 64)
    stA, dA, replayA, rng = make_pair(S, B, hidden, seed=5)
    stB, dB, replayB, _ = make_pair(S, B, hidden, seed=6)
    ra, total_a = caffe_ranges(S, hidden, False, na)
    rc, total_c = caffe_ranges(S, hidden, True, nc)
    assert total_a == stA.actor.size and total_c == stA.critic.size

    def share(src, dst):
        for name, ranges in (("actor", ra), ("actor_target", ra), ("critic", rc), ("critic_target", rc)):
            for b, e in ranges:
                getattr(dst, name)[b:e] = getattr(src, name)[b:e]

    share(stA, stB)                                   # ShareParameters: the slave's blobs become the owner's
    dB.copy_shared_layers_from(dA, na, nc)
    for step in range(6):
        upd_o, upd_d, other_o, other_d, replay = (stA, dA, stB, dB, replayA) if step % 2 == 0 else (stB, dB, stA, dA, replayB)
        idx = rng.integers(0, replay[0].shape[0], B).astype(np.int32)
        lo, qo = oracle_step(upd_o, replay, idx)
        ld, qd = upd_d.update_with_indices(idx)
        assert abs(ld - lo) <= RTOL * abs(lo) + 1e-7 and abs(qd - qo) <= RTOL * abs(qo) + 1e-6
        share(upd_o, other_o)
        other_d.copy_shared_layers_from(upd_d, na, nc)
    for st, d in ((stA, dA), (stB, dB)):
        for net, ref in ((P.ACTOR, st.actor), (P.CRITIC, st.critic), (P.ACTOR_TARGET, st.actor_target), (P.CRITIC_TARGET, st.critic_target)):
            got = d.get_params(net)
            assert np.abs(got - ref).max() <= 2e-2 * 1e-3 + RTOL * np.abs(ref).max(), (net, np.abs(got - ref).max())
    # shared ranges are bit-identical between the two handles, everything else differs
    for net, ranges, total in ((P.ACTOR, ra, total_a), (P.CRITIC, rc, total_c), (P.ACTOR_TARGET, ra, total_a), (P.CRITIC_TARGET, rc, total_c)):
        a, b = dA.get_params(net), dB.get_params(net)
        mask = np.zeros(total, bool)
        for s0, e0 in ranges:
            mask[s0:e0] = True
        assert np.array_equal(a[mask], b[mask])
        if (~mask).any():
            assert not np.array_equal(a[~mask], b[~mask])
    # the act path of the slave reads the shared actor layers too (its snapshot was refreshed)
    x = replayA[0][:8]
    ref = stB.actor_forward(x)
    got = dB.select_actions(x)
    assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max() + 1e-5
    dA.close(); dB.close()


def test_share_rejects_bad_arguments():
    P = pkg()
    d1 = P.DQNB(state_size=58, batch=32, hidden=(64, 64), replay_capacity=256)
    d2 = P.DQNB(state_size=58, batch=32, hidden=(64, 64, 64), replay_capacity=256)
    d3 = P.DQNB(state_size=58, batch=32, hidden=(64, 64), replay_capacity=256)
    with pytest.raises(RuntimeError, match="identical net shapes"):
        d1.copy_shared_layers_from(d2, 1, 1)
    with pytest.raises(RuntimeError, match="more layers"):
        d1.copy_shared_layers_from(d3, 5, 1)
    with pytest.raises(RuntimeError, match="bad argument"):
        d1.copy_shared_layers_from(d1, 1, 1)
    d1.close(); d2.close(); d3.close()
