"""Kernel unit tests: the dense contraction kernels (csrc/gemm.cuh) against float64 numpy, for every
operand-major combination the layers use (forward: K/K, dX: K/MN, dW: MN/MN) and split-K."""
import numpy as np
import pytest

from util import pkg, relerr

pytestmark = pytest.mark.gpu

CASES = [  # (a_mn, b_mn, M, N, K, splits)
    (0, 0, 128, 64, 32, 1),
    (0, 0, 256, 128, 96, 1),
    (0, 0, 1024, 512, 1024, 1),     # layer-2 forward at batch 1024
    (0, 1, 128, 64, 64, 1),
    (0, 1, 1024, 1024, 512, 1),     # layer-2 dX
    (1, 1, 128, 64, 128, 1),
    (1, 1, 512, 1024, 1024, 2),     # layer-2 dW, split-K over the minibatch
    (1, 1, 64, 128, 256, 4),        # M smaller than the 128-row tile (TMA zero fill)
    (1, 0, 128, 64, 64, 1),
    (0, 0, 128, 64, 4096, 8),
]


@pytest.mark.parametrize("mode", [1, 0], ids=["simt_fp32", "tcgen05_3xtf32"])
@pytest.mark.parametrize("a_mn,b_mn,M,N,K,splits", CASES)
def test_gemm_matches_float64(mode, a_mn, b_mn, M, N, K, splits):
    P = pkg()
    rng = np.random.default_rng(M + N + K + a_mn * 2 + b_mn)
    A = rng.normal(0, 1, (M, K)).astype(np.float32)
    B = rng.normal(0, 1, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    Cm, ms = P.gemm_test(mode, a_mn, b_mn, M, N, K, splits, Ain, Bin)
    # 3xTF32 drops only lo*lo terms (~2^-22 per product).  What remains is the accumulator: the
    # tensor core truncates (does not round) when adding into fp32, a bias that grows ~linearly with
    # K (measured 7e-6 at K=1024 with one accumulator); the fp32 FFMA path stays near 2e-6.
    tol = 2e-5 if mode == 0 else 3e-6
    assert relerr(Cm, ref) < tol, (relerr(Cm, ref), ms)


@pytest.mark.parametrize("bn,stages", [(128, 0), (128, 2), (64, 2), (64, 3), (32, 0), (32, 2)])
@pytest.mark.parametrize("a_mn,b_mn,M,N,K,splits", [(0, 0, 1024, 512, 1024, 1), (0, 1, 1024, 1024, 512, 1), (1, 1, 512, 1024, 1024, 2),
                                                    (1, 1, 64, 128, 256, 4), (0, 0, 256, 192, 96, 1)])
def test_gemm_tile_and_ring_variants(bn, stages, a_mn, b_mn, M, N, K, splits):
    """128x128 tiles and shallower TMA->MMA rings (two CTAs per SM) of the tcgen05 kernel; N=192 leaves the last
    128-wide tile half outside the matrix (TMA clips loads and stores)."""
    P = pkg()
    rng = np.random.default_rng(7 * M + N + K + bn + stages)
    A = rng.normal(0, 1, (M, K)).astype(np.float32)
    B = rng.normal(0, 1, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    Ain = np.ascontiguousarray(A.T) if a_mn else A
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    Cm, ms = P.gemm_test(0 | (stages << 12) | (bn << 16), a_mn, b_mn, M, N, K, splits, Ain, Bin)
    assert relerr(Cm, ref) < 2e-5, (relerr(Cm, ref), ms)


def test_gemm_exact_on_small_integers():
    """Integer-valued operands are exact in TF32, so any layout/descriptor bug shows up as a
    bit-level mismatch rather than a tolerance question."""
    P = pkg()
    rng = np.random.default_rng(3)
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 1), (1, 0)):
        M, N, K = 256, 128, 64
        A = rng.integers(-4, 5, (M, K)).astype(np.float32)
        B = rng.integers(-4, 5, (N, K)).astype(np.float32)
        ref = (A.astype(np.int64) @ B.astype(np.int64).T).astype(np.float32)
        Ain = np.ascontiguousarray(A.T) if a_mn else A
        Bin = np.ascontiguousarray(B.T) if b_mn else B
        Cm, _ = P.gemm_test(0, a_mn, b_mn, M, N, K, 1, Ain, Bin)
        assert np.array_equal(Cm, ref), (a_mn, b_mn, np.abs(Cm - ref).max())


@pytest.mark.parametrize("bn", [32, 64, 128])
@pytest.mark.parametrize("b_mn,M,N,K,splits", [(0, 1024, 512, 1024, 1), (1, 1024, 1024, 512, 1), (0, 256, 192, 96, 1), (0, 128, 64, 4096, 8),
                                               (1, 128, 128, 32, 1), (0, 1024, 256, 544, 2)])
def test_gemm_a_in_tensor_memory(bn, b_mn, M, N, K, splits):
    """A-in-TMEM mainloop (K-major A copied shared -> tensor memory by the mover warps, tcgen05.mma reads it there):
    same products as the shared-memory mode, for K-major and MN-major B, ragged K splits and a single k-block."""
    P = pkg()
    rng = np.random.default_rng(11 * M + N + K + bn + b_mn)
    A = rng.normal(0, 1, (M, K)).astype(np.float32)
    B = rng.normal(0, 1, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    Bin = np.ascontiguousarray(B.T) if b_mn else B
    Cm, ms = P.gemm_test(0 | (4 << 8) | (bn << 16), 0, b_mn, M, N, K, splits, A, Bin)
    assert relerr(Cm, ref) < 2e-5, (relerr(Cm, ref), ms)
    Ai = rng.integers(-4, 5, (M, K)).astype(np.float32)
    Bi = rng.integers(-4, 5, (N, K)).astype(np.float32)
    Ci, _ = P.gemm_test(0 | (4 << 8) | (bn << 16), 0, b_mn, M, N, K, splits, Ai, np.ascontiguousarray(Bi.T) if b_mn else Bi)
    assert np.array_equal(Ci, (Ai.astype(np.int64) @ Bi.astype(np.int64).T).astype(np.float32))
