"""GPU experiment 2: MMA pipe pacing (cycles from clock64 inside the MMA-issuing warp)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
P = load_package()
L = P.lib()
L.dqnb_gemm_test_clocks.argtypes = [C.POINTER(C.c_longlong)]
rng = np.random.default_rng(0)
def run(M,N,K,sp,dbg,a_mn=0,b_mn=0):
    A = rng.normal(0,1,(M,K)).astype(np.float32); B = rng.normal(0,1,(N,K)).astype(np.float32)
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    Cm, ms = P.gemm_test(0 | (dbg<<8), a_mn, b_mn, M, N, K, sp, Ain, Bin)
    c = (C.c_longlong*3)(); L.dqnb_gemm_test_clocks(c)
    err = None
    if (dbg & 3) == 0:
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        err = float(np.abs(Cm - ref).max() / np.abs(ref).max())
    return ms*1e3, c[1]-c[0], c[2]-c[0], err
K = 4096; kb = K//32
print("single CTA, K=4096 (128 k-blocks). dbg: 1 noMMA 2 noTMA; bn<<8")
for bn in (64, 128):
    for tag, dbg in [("full", 0), ("noTMA", 2), ("noMMA", 1), ("neither", 3)]:
        us, issued, done, err = run(128, bn, K, 1, dbg | (bn << 8))
        print(f"bn={bn:3d} {tag:8s} {us:8.2f} us  issue {issued:8d} cyc  complete {done:8d} cyc -> {done/kb:7.1f} cyc/k-block  err={err}", flush=True)
print("whole-GEMM timings (us), event-timed back-to-back launches")
for (M,N,K,sp,tag,a_mn,b_mn) in [(1024,1024,64,1,"L1fwd",0,0),(1024,512,1024,1,"L2fwd",0,0),(1024,256,512,1,"L3fwd",0,0),(1024,128,256,1,"L4fwd",0,0),
                       (1024,1024,512,1,"L2dx",0,1),(512,1024,1024,2,"L2dW sp2",1,1),(512,1024,1024,4,"L2dW sp4",1,1),(1024,512,1024,2,"L2fwd sp2",0,0),(1024,512,1024,4,"L2fwd sp4",0,0)]:
    for bn in (64, 128):
        us, _, _, err = run(M,N,K,sp, bn<<8, a_mn, b_mn)
        fl = 2*M*N*K
        print(f"{tag:10s} bn={bn:3d} {us:8.2f} us  {fl/us*1e-6:7.1f} TFLOP/s(alg)  err={err:.2e}", flush=True)
