"""GPU experiment 3: where do the microseconds of one small GEMM launch go? (globaltimer stamps)"""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
P = load_package()
L = P.lib()
L.dqnb_gemm_test_clocks.argtypes = [C.POINTER(C.c_longlong), C.c_int]
rng = np.random.default_rng(0)
names = ["entry", "prologue", "pdl_wait", "operands", "acc_done", "epi_done", "exit"]
def run(M,N,K,sp,bn,a_mn=0,b_mn=0):
    A = rng.normal(0,1,(M,K)).astype(np.float32); B = rng.normal(0,1,(N,K)).astype(np.float32)
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    _, ms = P.gemm_test(0 | ((bn<<8)<<8), a_mn, b_mn, M, N, K, sp, Ain, Bin)
    c = (C.c_longlong*(16*21))(); L.dqnb_gemm_test_clocks(c, 16*21)
    t = np.array(list(c), dtype=np.int64).reshape(21, 16)[:, :7]
    return ms*1e3, t
for (M,N,K,sp,bn,tag,a,b) in [(1024,1024,64,1,64,"L1fwd",0,0),(1024,128,256,1,64,"L4fwd",0,0),(1024,512,1024,1,64,"L2fwd",0,0),(1024,512,1024,4,128,"L2fwd sp4 bn128",0,0),
                              (1024,256,512,1,64,"L3fwd",0,0),(128,256,1024,8,64,"L4dW",1,1)]:
    us, t = run(M,N,K,sp,bn,a,b)
    t = t[5:]   # steady-state launches
    rel = (t - t[:, :1])
    gaps = t[1:, 0] - t[:-1, 6]          # previous exit -> next entry (CTA 0 of each)
    period = t[1:, 0] - t[:-1, 0]
    print(f"{tag:18s} event-timed {us:7.2f} us/launch | period {period.mean()/1e3:6.2f} us | exit->next entry {gaps.mean()/1e3:6.2f} us")
    print("    " + "  ".join(f"{n}+{rel[:, i].mean()/1e3:5.2f}" for i, n in enumerate(names)))
