"""GPU experiment: what paces the A-in-TMEM mainloop - the movers or the MMAs?  (dbg bits: 1 no MMA, 2 no TMA, 4 TS)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
P = load_package()
rng = np.random.default_rng(0)
K = 8192; kb = K // 32
A = rng.normal(0, 1, (128, K)).astype(np.float32)
for bn in (64, 128):
    B = rng.normal(0, 1, (bn, K)).astype(np.float32)
    for tag, dbg in [("SS", 0), ("SS noMMA", 1), ("TS", 4), ("TS noMMA (movers+TMA)", 5), ("TS noTMA", 6), ("TS noMMA noTMA (movers)", 7)]:
        _, ms = P.gemm_test(0 | (dbg << 8) | (bn << 16), 0, 0, 128, bn, K, 1, A, B)
        print(f"1 CTA bn={bn:3d} {tag:26s} {ms*1e3:8.2f} us  {ms*1e6/kb*1.965:7.1f} cyc/k-block", flush=True)
