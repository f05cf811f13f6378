"""Asynchronous learner API (dqnb_update_async / dqnb_results) with replay appends on the copy stream.

The reference is strictly sequential: AddTransitions, then Update() samples the memory it finds
(dqn_main.cpp:145-150, :358-363; dqn.cpp:799-826).  The asynchronous path must give exactly that:
appends enqueued while updates are in flight may only become visible between the gather of one update
and the gather of the next.  Two learners with identical weights and memory run the same add/update
sequence, one blocking, one pipelined; a small ring forces evictions (rows being overwritten while
earlier updates are still running).  Everything must agree bit for bit."""
import numpy as np
import pytest

from util import pkg
from bench import synth_replay

pytestmark = pytest.mark.gpu


def _make(P, S, B, hidden, cap):
    d = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=cap, seed=11)
    d.init_params(seed=4, std=0.05)
    return d


@pytest.mark.parametrize("use_graph", [1, 0])
def test_pipelined_adds_and_updates_equal_sequential(use_graph):
    P = pkg()
    S, B, hidden, cap = 58, 128, (256, 128, 64, 64), 700
    s, a, r, mc, term, sn = synth_replay(600 + 40 * 96, S, 7)
    seq = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=cap, seed=11, use_graph=use_graph)
    pip = P.DQNB(state_size=S, batch=B, hidden=hidden, replay_capacity=cap, seed=11, use_graph=use_graph)
    for d in (seq, pip):
        d.init_params(seed=4, std=0.05)
        d.add_transitions(s[:600], a[:600], r[:600], mc[:600], sn[:600], term[:600])
    n_steps = 40
    loss_seq, q_seq = [], []
    for k in range(n_steps):
        o = 600 + 96 * k
        seq.add_transitions(s[o:o + 96], a[o:o + 96], r[o:o + 96], mc[o:o + 96], sn[o:o + 96], term[o:o + 96])
        l, q = seq.update(1)
        loss_seq.append(l[0]); q_seq.append(q[0])
    loss_pip, q_pip = [], []
    first = None
    for k in range(n_steps):
        o = 600 + 96 * k
        pip.add_transitions(s[o:o + 96], a[o:o + 96], r[o:o + 96], mc[o:o + 96], sn[o:o + 96], term[o:o + 96])
        step = pip.update_async(1)
        if first is None:
            first = step
        if k > 0:                                  # read the previous step's result while this one runs
            l, q = pip.results(step - 1, 1)
            loss_pip.append(l[0]); q_pip.append(q[0])
    l, q = pip.results(first + n_steps - 1, 1)
    loss_pip.append(l[0]); q_pip.append(q[0])
    assert np.array_equal(np.array(loss_seq), np.array(loss_pip))
    assert np.array_equal(np.array(q_seq), np.array(q_pip))
    for net in range(4):
        assert np.array_equal(seq.get_params(net), pip.get_params(net))
    assert seq.memory_size() == pip.memory_size()
    got_seq = seq.get_transitions(0, seq.memory_size())
    got_pip = pip.get_transitions(0, pip.memory_size())
    for x, y in zip(got_seq, got_pip):
        assert np.array_equal(x, y)
    # a batch of results in one call, and the bounds of the result window
    l_all, q_all = pip.results(first, n_steps)
    assert np.array_equal(l_all, np.array(loss_seq)) and np.array_equal(q_all, np.array(q_seq))
    with pytest.raises(Exception):
        pip.results(first + n_steps, 1)            # never enqueued
    seq.close(); pip.close()


def test_update_async_many_then_results():
    P = pkg()
    S, B, hidden = 58, 64, (128, 64, 64, 32)
    s, a, r, mc, term, sn = synth_replay(2000, S, 3)
    d1 = _make(P, S, B, hidden, 4000); d2 = _make(P, S, B, hidden, 4000)
    for d in (d1, d2):
        d.add_transitions(s, a, r, mc, sn, term)
    l1, q1 = d1.update(25)
    last = d2.update_async(25)
    l2, q2 = d2.results(last - 24, 25)
    assert np.array_equal(l1, l2) and np.array_equal(q1, q2)
    d1.close(); d2.close()
