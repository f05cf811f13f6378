"""GPU: latency of the act path (dqnb_select_actions) idle / beside asynchronous updates."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import synth_replay
P = load_package()
d = P.DQNB(state_size=58, batch=1024, hidden=(1024, 512, 256, 128), replay_capacity=70000, max_act_batch=128)
d.init_params(2, 0.01)
s, a, r, mc, term, sn = synth_replay(65536, 58, 1)
d.add_transitions(s, a, r, mc, sn, term)
d.update(5)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for busy in (0, 1):
    for n in (1, 8, 64, 128):
        x = np.ascontiguousarray(s[:n])
        for _ in range(20): d.select_actions(x)
        if busy: last = d.update_async(400)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); d.select_actions(x); ts.append(time.perf_counter() - t0)
        if busy: d.results(last, 1)
        ts = np.array(ts) * 1e6
        print(f"busy={busy} n={n:4d}: median {np.median(ts):7.1f} us  p10 {np.percentile(ts,10):7.1f}  p90 {np.percentile(ts,90):7.1f}", flush=True)
d.close()
