"""The act path (SelectActionGreedily, dqn.cpp:734-766): skinny-M kernels on their own stream and graph, reading the
fp32 actor snapshot the optimiser publishes at the end of every update; rows and completion travel through
host-mapped memory.  Checked against the oracle's actor forward, across updates, and while updates are in flight
(an action batch must come from ONE completed update's actor, never from a mix of two)."""
import numpy as np
import pytest

from util import RTOL, make_pair, oracle_step, pkg, relerr
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 8, 9, 33, 64, 100])
def test_select_actions_matches_oracle_for_every_row_count(n):
    """1..64 rows take the skinny path (row groups of 8, ragged tails), 100 rows the tcgen05 layer kernels."""
    S, B, hidden = 58, 128, (1024, 512, 256, 128)
    st, d, replay, rng = make_pair(S, B, hidden, "warm", 0, n_replay=512)
    s = replay[0]
    got = d.select_actions(s[:n])
    ref = st.actor_forward(s[:n])
    assert got.shape == (n, 10)
    assert relerr(got, ref) < RTOL, relerr(got, ref)
    # per row, against that row's own scale (a row of small outputs must not hide behind a large one)
    rows = np.abs(got - ref).max(axis=1) / (np.abs(ref).max(axis=1) + 1e-6)
    assert rows.max() < 5 * RTOL, rows.max()
    d.close()


def test_act_path_follows_the_updates():
    """After every blocking update the next SelectActions uses the updated actor (the reference's order of events:
    Update() returns, then the next env step calls SelectAction, dqn_main.cpp:360-366)."""
    S, B, hidden = 59, 64, (256, 128, 64, 32)
    st, d, replay, rng = make_pair(S, B, hidden, "warm", 0, n_replay=512, actor_lr=1e-3)   # a visible actor step
    s = replay[0]
    prev = None
    for u in range(4):
        got = d.select_actions(s[:8])
        ref = st.actor_forward(s[:8])
        assert relerr(got, ref) < 3 * RTOL, (u, relerr(got, ref))
        if prev is not None:
            assert np.abs(got - prev).max() > 1e-4 * np.abs(prev).max(), "the act path did not see the update"
        prev = got
        idx = rng.integers(0, d.memory_size(), B).astype(np.int32)
        oracle_step(st, replay, idx)
        d.update_with_indices(idx)
    # set_params replaces the snapshot as well
    P = pkg()
    d.set_params(P.ACTOR, st.actor * 0.5)
    st2 = O.OracleState(st.cfg, st.actor * 0.5, st.critic, st.actor_target, st.critic_target)
    assert relerr(d.select_actions(s[:8]), st2.actor_forward(s[:8])) < RTOL
    d.close()


def test_acting_while_updates_are_in_flight_never_sees_a_torn_actor():
    """Enqueue a train of asynchronous updates and act continuously beside them.  Every action batch must be bit-equal
    to what ONE of the learner's successive actors produces (replayed afterwards, update by update, on a second
    learner with the same seed): the snapshot flip is atomic with respect to an act call."""
    P = pkg()
    S, B, hidden, n_up = 58, 256, (512, 256, 128, 64), 60
    rng = np.random.default_rng(5)
    cfg = O.make_config(state_size=S, batch=2048, hidden=hidden)
    s, a, r, mc, term, sn = O.synth_batch(cfg, rng, p_term=0.1)
    probe = np.ascontiguousarray(s[:8])

    def learner():
        d = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=4096, seed=11, actor_lr=1e-3, max_act_batch=64)
        d.init_params(3, 0.05)
        d.add_transitions(s, a, r, mc, sn, term)
        return d

    d = learner()
    seen = []
    last = d.update_async(n_up)
    for _ in range(1500):                      # ~10 ms of updates in flight, ~25 us per act call
        seen.append(d.select_actions(probe).copy())
    d.results(last, 1)
    seen.append(d.select_actions(probe).copy())
    d.close()
    ref = learner()
    actors = [ref.select_actions(probe).copy()]
    for _ in range(n_up):
        ref.update(1)
        actors.append(ref.select_actions(probe).copy())
    ref.close()
    distinct = {x.tobytes() for x in actors}
    assert len(distinct) > n_up // 2, "the probe does not distinguish successive actors"
    bad = [i for i, x in enumerate(seen) if x.tobytes() not in distinct]
    assert not bad, f"{len(bad)} of {len(seen)} action batches match no completed update's actor (first at {bad[0]})"
    assert seen[-1].tobytes() == actors[-1].tobytes()
    assert len({x.tobytes() for x in seen}) > 3, "acting never overlapped the updates"
