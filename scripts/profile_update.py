"""Driver for ncu: a few UpdateActorCritic steps at the BASELINE cfg2 shape (S=58, B=1024)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import synth_replay
P = load_package()
n_updates = int(sys.argv[1]) if len(sys.argv) > 1 else 3
use_graph = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = P.DQNB(state_size=58, batch=1024, hidden=(1024, 512, 256, 128), replay_capacity=70000, use_graph=use_graph)
d.init_params(2, 0.01)
s, a, r, mc, term, sn = synth_replay(65536, 58, 1)
d.add_transitions(s, a, r, mc, sn, term)
loss, q = d.update(n_updates)
print("losses", loss, "avg_q", q, "launches", d.kernel_launches())
d.close()
