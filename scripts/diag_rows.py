"""GPU diagnostic: is the d_raw mismatch a few rows (ReLU kink flips) or broad?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import make_pair, oracle_step, relerr
from oracle import oracle as O
O.load_blas()
for name, (S, B, hidden, n_replay) in {"cfg3": (77, 4096, (1024, 512, 256, 128), 8192), "cfg5": (58, 1024, (1024, 1024, 1024, 1024), 4096)}.items():
    for gm in (1, 0):
        st, d, replay, rng = make_pair(S, B, hidden, "warm", gm, n_replay=n_replay, use_blas=0)
        idx = rng.integers(0, d.memory_size(), B).astype(np.int32)
        oracle_step(st, replay, idx, taps=True)
        d.update_with_indices(idx)
        t = st.last_taps
        for key, w in (("q_pi", 1), ("d_raw", 10)):
            got = d.debug_read(key, B * w).reshape(B, w).astype(np.float64); ref = t[key].reshape(B, w).astype(np.float64)
            e = np.abs(got - ref).max(1) / np.abs(ref).max()
            qs = np.quantile(e, [0.5, 0.9, 0.99, 0.999, 1.0])
            print(f"{name} m{gm} {key:6s} row-err quantiles 50/90/99/99.9/100%: " + " ".join(f"{x:.1e}" for x in qs), f" rows>1e-4: {(e > 1e-4).sum()} of {B}", flush=True)
        d.close()
