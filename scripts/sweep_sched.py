"""GPU: sweep the schedule / tile-shape knobs of engine.cu (Tuning) and time the captured update graph.

usage: python scripts/sweep_sched.py [batch] > gpurun_out/sweep.txt
"""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import synth_replay
P = load_package()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
s, a, r, mc, term, sn = synth_replay(65536, 58, 1)
KEYS = ["DQNB_TS_MIN_KB", "DQNB_BN32_MAX_TILES", "DQNB_BN_SIDE_L1", "DQNB_SIDE_DELAY", "DQNB_FUSE_COLSUM", "DQNB_FUSE_TL", "DQNB_ACTOR_LATE", "DQNB_GATHER_AHEAD", "DQNB_PDL_EARLY", "DQNB_CLUSTER_B", "DQNB_ST_FWD", "DQNB_ST_FWD_SIDE", "DQNB_ST_DX", "DQNB_ST_DW", "DQNB_SCHED", "DQNB_BN_FWD", "DQNB_BN_FWD_SIDE", "DQNB_BN_DX", "DQNB_BN_DW"]

def run(cfg):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ[k] = str(v)
    d = P.DQNB(state_size=58, batch=B, hidden=(1024, 512, 256, 128), replay_capacity=70000, use_graph=1)
    d.init_params(2, 0.01)
    d.add_transitions(s, a, r, mc, sn, term)
    loss, q = d.update(30)
    best = min(d.benchmark(300) / 300 * 1e3 for _ in range(3))
    d.close()
    tag = " ".join(f"{k[5:]}={v}" for k, v in cfg.items()) or "(defaults)"
    print(f"{best:7.1f} us/update  {B / best:6.3f}e6 tr/s   loss={loss[-1]:.6f} q={q[-1]:.6f}   {tag}", flush=True)

cfgs = [{}]
if len(sys.argv) > 2 and sys.argv[2] == "full":
    for sched, dw0, cs in itertools.product([0, 1], [0, 1], [0, 1]):
        if dw0 == 0 and cs == 1:
            continue
        cfgs.append({"DQNB_SCHED": sched, "DQNB_DW0_MAIN": dw0, "DQNB_COLSUM_SIDE": cs})
    for key in ["DQNB_BN_FWD", "DQNB_BN_FWD_SIDE", "DQNB_BN_DX", "DQNB_BN_DW"]:
        cfgs.append({"DQNB_SCHED": 1, "DQNB_DW0_MAIN": 1, key: 128})
    cfgs.append({"DQNB_SCHED": 1, "DQNB_DW0_MAIN": 1, "DQNB_BN_DX": 128, "DQNB_BN_DW": 128})
    cfgs.append({"DQNB_SCHED": 1, "DQNB_DW0_MAIN": 1, "DQNB_BN_DX": 128, "DQNB_BN_DW": 128, "DQNB_BN_FWD_SIDE": 128})
    cfgs.append({"DQNB_SCHED": 1, "DQNB_DW0_MAIN": 1, "DQNB_BN_DX": 128, "DQNB_BN_DW": 128, "DQNB_BN_FWD_SIDE": 128, "DQNB_BN_FWD": 128})
else:
    import json
    for arg in sys.argv[2:]:
        cfgs.append(json.loads(arg))
for c in cfgs:
    try:
        run(c)
    except Exception as e:   # keep sweeping
        print("FAILED", c, repr(e), flush=True)
