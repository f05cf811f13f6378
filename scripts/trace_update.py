"""GPU: timeline of the GEMM launches of one update (DQNB_TRACE=1), from in-kernel globaltimer stamps."""
import os, sys, ctypes as C
os.environ["DQNB_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from bench import synth_replay
P = load_package()
L = P.lib()
L.dqnb_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int64]
L.dqnb_debug_trace.restype = C.c_int64
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = P.DQNB(state_size=58, batch=B, hidden=(1024, 512, 256, 128), replay_capacity=70000, use_graph=1)
d.init_params(2, 0.01)
s, a, r, mc, term, sn = synth_replay(65536, 58, 1)
d.add_transitions(s, a, r, mc, sn, term)
d.update(20)
ms = d.benchmark(200)
print(f"graph replay: {ms/200*1e3:.1f} us per update")
d.update(1)
buf = (C.c_longlong * 4096)()
n = L.dqnb_debug_trace(d._h, buf, 4096)
t = np.array(list(buf[:n]), dtype=np.int64).reshape(-1, 8)
kinds = ["GEMM","GATHER","SAMPLE","HEAD_FWD","CRITIC_HEAD","ACTOR_HEAD_BWD","HEAD_BWD_W","COLSUM","REDUCE","ALLREDUCE","P2P_ALLREDUCE","ADAM","PREP","FINALIZE","FORK","JOIN"]
g = [i for i in range(len(t)) if t[i,7]//1000 == 0]
t0 = min(t[i,0] for i in g)
print(" op kind        br grid(x,y,z) kb |  entry  pdlwait  operands acc_done epi_done | dur(after wait)")
for i in range(len(t)):
    k, br = int(t[i,7]//1000), int(t[i,7]%1000)
    if k != 0:
        print(f"{i:3d} {kinds[k]:14s} {br}")
        continue
    gx, gy, gz, kb = (t[i,6]>>40)&0xfff, (t[i,6]>>20)&0xfffff, t[i,6]&0xfffff, (t[i,6]>>52)
    e = [(t[i,j]-t0)/1e3 for j in range(6)]
    print(f"{i:3d} GEMM           {br} ({gx:2d},{gy:2d},{gz:2d}) {kb:3d} | {e[0]:7.1f} {e[2]:7.1f} {e[3]:8.1f} {e[4]:8.1f} {e[5]:8.1f} | {e[5]-e[2]:6.1f}")
d.close()
