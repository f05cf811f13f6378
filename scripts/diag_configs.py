"""GPU diagnostic: per-tensor errors of oracle / SIMT / tcgen05 against a float64 autograd reference."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import make_pair, relerr
from oracle import oracle as O
import refmodel as R
O.load_blas()
cases = {"cfg3": (77, 4096, (1024, 512, 256, 128), 8192), "cfg5": (58, 1024, (1024, 1024, 1024, 1024), 4096), "cfg5b": (58, 1024, (1024, 1024, 1024, 1024), 2048),
         "mid": (58, 1024, (1024, 1024, 256, 128), 4096), "std": (58, 1024, (1024, 512, 256, 128), 4096)}
keys = ("y", "q", "a_pi", "q_pi", "d_raw", "d_inv", "critic_grad", "actor_grad")
for name in sys.argv[1:] or list(cases):
    S, B, hidden, n_replay = cases[name]
    ref = None
    for gm in (1, 0):
        st, d, replay, rng = make_pair(S, B, hidden, "warm", gm, n_replay=n_replay, use_blas=0)
        idx = rng.integers(0, d.memory_size(), B).astype(np.int32)
        s, a, r, mc, term, sn = replay
        if ref is None:
            st64 = {k: np.asarray(getattr(st, k), np.float64) for k in ("actor", "critic", "actor_target", "critic_target", "actor_m", "actor_v", "critic_m", "critic_v")}
            st64["actor_iter"] = st64["critic_iter"] = 0
            _, ref = R.update(st.cfg, st64, s[idx], a[idx], r[idx], mc[idx], term[idx], sn[idx])
            st.update(s[idx], a[idx], r[idx], mc[idx], term[idx], sn[idx], taps=True)
            print(name, "oracle ", " ".join(f"{k}={relerr(st.last_taps[k], np.asarray(ref[k]).ravel()):.1e}" for k in keys), flush=True)
        d.update_with_indices(idx)
        print(name, "gpu m%d" % gm, " ".join(f"{k}={relerr(d.debug_read(k, np.asarray(ref[k]).size), np.asarray(ref[k]).ravel()):.1e}" for k in keys), flush=True)
        d.close()
