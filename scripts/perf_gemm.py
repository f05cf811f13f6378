"""GPU experiment: where does the tcgen05 GEMM spend its time? (TMA-only / MMA-only / both)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
P = load_package()
rng = np.random.default_rng(0)
def run(M,N,K,sp,dbg,a_mn=0,b_mn=0):
    A = rng.normal(0,1,(M,K)).astype(np.float32); B = rng.normal(0,1,(N,K)).astype(np.float32)
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    _, ms = P.gemm_test(0 | (dbg<<8), a_mn, b_mn, M, N, K, sp, Ain, Bin)
    return ms*1e3
names = {0:"full", 1:"tma-only", 2:"mma-only", 4:"1mma+tma", 6:"1mma-only", 3:"neither"}
for (M,N,K,sp,tag) in [(128,64,32,1,"1cta-1kb"),(128,64,8192,1,"1cta-256kb"),(1024,512,2048,1,"64cta-64kb"),(1024,1024,2048,1,"128cta-64kb"),
                       (1024,512,1024,1,"L2fwd"),(1024,512,1024,2,"L2fwd-sp2"),(2048,1024,1024,1,"256cta-32kb")]:
    for dbg in (0,1,2,4,6,3):
        us = run(M,N,K,sp,dbg)
        kb = K//32//sp
        print(f"{tag:14s} {names[dbg]:10s} {us:9.2f} us   per-kblock {us/kb*1e3:8.1f} ns", flush=True)
for (a_mn,b_mn) in ((0,1),(1,1)):
    for dbg in (0,1,2):
        us = run(1024,512,1024,1,dbg,a_mn,b_mn)
        print(f"L2 a_mn={a_mn} b_mn={b_mn} {names[dbg]:10s} {us:9.2f} us")
