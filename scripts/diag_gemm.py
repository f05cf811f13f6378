"""GPU diagnostic: per-case error of the tcgen05 GEMM and the structure of any mismatch."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
P = load_package()

def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())

cases = [(0,0,128,64,32,1),(0,0,128,64,128,1),(0,0,128,64,160,1),(0,0,128,64,256,1),(0,0,1024,512,1024,1),
         (0,1,128,64,32,1),(0,1,128,64,64,1),(1,0,128,64,32,1),(1,1,128,64,32,1),(1,1,128,64,128,1),
         (0,0,128,64,4096,8),(1,1,512,1024,1024,2)]
for mode in (0, 1):
    for (a_mn,b_mn,M,N,K,sp) in cases:
        rng = np.random.default_rng(1)
        A = rng.normal(0,1,(M,K)).astype(np.float32); B = rng.normal(0,1,(N,K)).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        Ain = np.ascontiguousarray(A.T) if a_mn else A
        Bin = np.ascontiguousarray(B.T) if b_mn else B
        Cm, ms = P.gemm_test(mode, a_mn, b_mn, M, N, K, sp, Ain, Bin)
        print(f"mode={mode} a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K} sp={sp} relerr={rel(Cm,ref):.3e} ms={ms:.4f}", flush=True)

# structure: one-hot probes.  A = e_(m0,k0), B = all-ones rows scaled by index -> see which (n) get which k
print("--- integer probes (tcgen05)")
for (a_mn, b_mn) in ((0,0),(0,1),(1,0),(1,1)):
    M,N,K = 128,64,32
    rng = np.random.default_rng(3)
    A = rng.integers(-4,5,(M,K)).astype(np.float32); B = rng.integers(-4,5,(N,K)).astype(np.float32)
    ref = A @ B.T
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    Cm,_ = P.gemm_test(0, a_mn, b_mn, M, N, K, 1, Ain, Bin)
    bad = (Cm != ref)
    print(f"a_mn={a_mn} b_mn={b_mn}: mismatched {bad.mean()*100:.1f}% rows_bad={np.where(bad.any(1))[0][:20]} cols_bad={np.where(bad.any(0))[0][:40]}")
    if bad.any():
        # try to explain: does Cm equal A @ B[perm].T for a permutation of k within B?
        # probe with B = one-hot in k to find what k the hardware pairs with
        for k0 in (0, 1, 4, 8, 9, 31):
            A1 = np.zeros((M,K),np.float32); A1[:, k0] = 1.0          # picks column k0 of B
            Bp = np.arange(N*K, dtype=np.float32).reshape(N,K) % 1024   # B[n,k] = n*K+k (exact small ints)
            Ain = np.ascontiguousarray(A1.T) if a_mn else A1
            Bin = np.ascontiguousarray(Bp.T) if b_mn else Bp
            C1,_ = P.gemm_test(0, a_mn, b_mn, M, N, K, 1, Ain, Bin)
            print(f"   probe A=e_k{k0}: C[0,:8]={C1[0,:8]} expect {Bp[:8,k0]}  C[1,:4]={C1[1,:4]} C[33,:4]={C1[33,:4]}")
        for n0 in (0, 1, 33):
            B1 = np.zeros((N,K),np.float32); B1[n0,:] = 1.0             # C[m,n0] = sum_k A[m,k]
            Ap = (np.arange(M*K, dtype=np.float32).reshape(M,K) % 64)
            Ain = np.ascontiguousarray(Ap.T) if a_mn else Ap
            Bin = np.ascontiguousarray(B1.T) if b_mn else B1
            C1,_ = P.gemm_test(0, a_mn, b_mn, M, N, K, 1, Ain, Bin)
            nz = np.where(np.abs(C1).sum(0) > 0)[0]
            print(f"   probe B=row{n0} ones: nonzero cols={nz[:10]} C[:4,n0]={C1[:4,n0]} expect {Ap[:4].sum(1)}")
