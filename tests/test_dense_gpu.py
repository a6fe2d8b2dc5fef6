"""Building blocks of the replicated m x m stage (K6) against NumPy/LAPACK: DMMA GEMM, Cholesky, inverse."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(ctx, ta, tb, M, N, K, alpha, A, B, beta, Cm, lower_only=False):
    from sparsergps_b200 import _lib as L
    A, B, Cm = np.asfortranarray(A), np.asfortranarray(B), np.asfortranarray(Cm.copy())
    L.check(ctx._lib.srgp_test_gemm(ctx.handle, int(ta), int(tb), M, N, K, alpha, L.ptr(A), A.shape[0], L.ptr(B),
                                    B.shape[0], beta, L.ptr(Cm), Cm.shape[0], int(lower_only), 0, None))
    return Cm


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 128, 16), (256, 384, 208), (384, 128, 1024)])
def test_dmma_gemm_all_layouts(ctx, ta, tb, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K + 2 * ta + tb)
    A = rng.normal(size=(K, M) if ta else (M, K))
    B = rng.normal(size=(N, K) if tb else (K, N))
    C0 = rng.normal(size=(M, N))
    opA, opB = (A.T if ta else A), (B.T if tb else B)
    ref = 0.75 * opA @ opB - 1.25 * C0
    out = _gemm(ctx, ta, tb, M, N, K, 0.75, A, B, -1.25, C0)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12 * np.abs(ref).max() * np.sqrt(K))
    out0 = _gemm(ctx, ta, tb, M, N, K, 1.0, A, B, 0.0, np.full((M, N), np.nan))     # beta = 0 never reads C
    np.testing.assert_allclose(out0, opA @ opB, rtol=0, atol=1e-12 * np.abs(ref).max() * np.sqrt(K))


def test_dmma_gemm_lower_only(ctx):
    rng = np.random.default_rng(5)
    M = 384
    A = rng.normal(size=(M, 128))
    C0 = rng.normal(size=(M, M))
    out = _gemm(ctx, 0, 1, M, M, 128, -1.0, A, A, 1.0, C0, lower_only=True)
    ref = C0 - A @ A.T
    for tm in range(3):
        for tn in range(3):
            blk = np.s_[tm * 128:(tm + 1) * 128, tn * 128:(tn + 1) * 128]
            if tn <= tm:
                np.testing.assert_allclose(out[blk], ref[blk], rtol=0, atol=1e-11)
            else:
                np.testing.assert_array_equal(out[blk], C0[blk])


@pytest.mark.parametrize("m", [5, 64, 128, 129, 257, 384, 640, 896, 1024, 1025])
def test_cholesky_inverse_logdet(ctx, m):
    from sparsergps_b200 import _lib as L
    rng = np.random.default_rng(m)
    X = rng.normal(size=(m, 8))
    D = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    A = np.asfortranarray(np.exp(-0.5 * D) + 1e-2 * np.eye(m) + 0.1 * (X @ X.T) / 8)
    Lo, Ai = np.empty((m, m), order="F"), np.empty((m, m), order="F")
    logdet, info = L.cd(), L.ci()
    L.check(ctx._lib.srgp_test_chol_inverse(ctx.handle, m, L.ptr(A), L.ptr(Lo), L.ptr(Ai), C.byref(logdet),
                                            C.byref(info), 0, None))
    assert info.value == 0
    Lref = np.linalg.cholesky(A)
    np.testing.assert_allclose(np.tril(Lo), Lref, rtol=0, atol=1e-11 * np.abs(Lref).max() * np.linalg.cond(A) ** 0.5)
    assert logdet.value == pytest.approx(2 * np.log(np.diag(Lref)).sum(), rel=1e-12, abs=1e-10)
    resid = np.abs(Ai @ A - np.eye(m)).max()
    assert resid < 1e-13 * np.linalg.cond(A) * m
    np.testing.assert_allclose(Ai, Ai.T, rtol=0, atol=1e-12 * np.abs(Ai).max())


def test_cholesky_reports_not_pd(ctx):
    from sparsergps_b200 import _lib as L
    m = 200
    A = np.asfortranarray(np.eye(m))
    A[150, 150] = -1.0
    logdet, info = L.cd(), L.ci()
    L.check(ctx._lib.srgp_test_chol_inverse(ctx.handle, m, L.ptr(A), None, None, C.byref(logdet), C.byref(info), 0, None))
    assert info.value == 151


def test_dense_timing(ctx):
    """Not an assertion on speed: prints the m = 1024 dense-stage costs for profiles/ (Amdahl)."""
    from sparsergps_b200 import _lib as L
    m = 1024
    rng = np.random.default_rng(0)
    A = rng.normal(size=(m, m))
    Cm = np.zeros((m, m), order="F")
    ms = L.cd()
    A = np.asfortranarray(A)
    L.check(ctx._lib.srgp_test_gemm(ctx.handle, 0, 1, m, m, m, 1.0, L.ptr(A), m, L.ptr(A), m, 0.0, L.ptr(Cm), m, 0, 20, C.byref(ms)))
    print("\ngemm NT 1024^3: %.3f ms  (%.1f TFLOP/s)" % (ms.value, 2 * m ** 3 / ms.value / 1e9))
    L.check(ctx._lib.srgp_test_gemm(ctx.handle, 1, 0, m, m, m, 1.0, L.ptr(A), m, L.ptr(A), m, 0.0, L.ptr(Cm), m, 0, 20, C.byref(ms)))
    print("gemm TN 1024^3: %.3f ms  (%.1f TFLOP/s)" % (ms.value, 2 * m ** 3 / ms.value / 1e9))
    S = np.asfortranarray(A @ A.T + m * np.eye(m))
    logdet, info = L.cd(), L.ci()
    L.check(ctx._lib.srgp_test_chol_inverse(ctx.handle, m, L.ptr(S), None, None, C.byref(logdet), C.byref(info), 10, C.byref(ms)))
    print("chol + inverse m=1024: %.3f ms" % ms.value)
