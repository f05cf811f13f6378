"""GPU experiment: tcgen05.mma pacing with the A operand in tensor memory (TS) vs shared memory (SS)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
P = load_package()
rng = np.random.default_rng(0)
def run(M,N,K,sp,dbg,bn,a_mn=0,b_mn=0):
    A = rng.normal(0,1,(M,K)).astype(np.float32); B = rng.normal(0,1,(N,K)).astype(np.float32)
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    _, ms = P.gemm_test(0 | (dbg << 8) | (bn << 16), a_mn, b_mn, M, N, K, sp, Ain, Bin)
    return ms * 1e3
K = 8192; kb = K // 32
for bn in (32, 64, 128):
    for tag, dbg in [("SS full", 0), ("SS noTMA", 2), ("TS full", 4), ("TS noTMA", 6)]:
        us = run(128, max(bn, 64), K, 1, dbg, bn)
        print(f"1 CTA  bn={bn:3d} {tag:9s} {us:8.2f} us  {us/kb*1e3*1.965:7.1f} cyc/k-block  {us/kb*1e3*1.965/12:6.1f} cyc/MMA", flush=True)
for bn in (64, 128):
    for tag, dbg in [("SS full", 0), ("TS full", 4)]:
        us = run(1024, 1024, 2048, 1, dbg, bn)
        print(f"128/64 CTAs bn={bn:3d} {tag:9s} {us:8.2f} us  {us/64*1e3*1.965:7.1f} cyc/k-block", flush=True)
