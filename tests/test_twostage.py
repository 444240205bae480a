"""Two-stage tridiagonalisation of large AE matrices (saamge_b200/csrc/twostage.cu).

CPU: the numpy statement of the algorithm (tests/twostage_ref.py) against numpy.linalg.
GPU: the CUDA kernels, through the C ABI's diagnostic entry points, against that statement
stage by stage (band, tridiagonal matrix, back-transformed eigenvectors)."""
import ctypes

import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

import twostage_ref as ts


def _sym(n, seed, scale_rows=False):
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(n, n))
    A = M + M.T
    if scale_rows:  # strongly graded entries (like a 1e6-contrast AE matrix after scaling)
        s = 10.0 ** rng.uniform(-3, 0, size=n)
        A = A * s[:, None] * s[None, :]
    return A / np.abs(A).max()


@pytest.mark.parametrize("n,b", [(5, 2), (9, 4), (20, 4), (37, 8), (70, 32), (33, 32), (34, 32), (35, 32),
                                 (65, 32), (66, 32), (131, 32), (97, 16)])
def test_reference_statement_is_a_tridiagonalisation(n, b):
    A = _sym(n, n)
    d, e, T, tau1, V2 = ts.tridiagonalise(A, b)
    w0 = np.linalg.eigvalsh(A)
    w, Y = eigh_tridiagonal(d, e[:n - 1])
    assert np.abs(w - w0).max() <= 1e-13 * n
    for j in range(min(3, n)):
        z = ts.back_transform(Y[:, j], T, tau1, V2, b)
        assert np.linalg.norm(A @ z - w[j] * z) <= 1e-13 * n
        assert abs(np.linalg.norm(z) - 1.0) <= 1e-13 * n


def _gpu_reduce(ctx, A):
    n = A.shape[0]
    lib = ctx.lib
    dp = ctypes.POINTER(ctypes.c_double)
    Af = np.asfortranarray(A)
    T = np.zeros((n, n), order="F")
    tau1, d, e = np.zeros(n), np.zeros(n), np.zeros(n)
    lib.sa_gpu_debug_twostage.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp, dp, dp, dp]
    rc = lib.sa_gpu_debug_twostage(ctx.h, n, Af.ctypes.data_as(dp), T.ctypes.data_as(dp), tau1.ctypes.data_as(dp),
                                   d.ctypes.data_as(dp), e.ctypes.data_as(dp))
    assert rc == 0, lib.sa_gpu_last_error()
    return T, tau1, d, e


@pytest.mark.gpu
@pytest.mark.parametrize("n,cluster", [(130, 2), (300, 4), (700, 8)])
def test_cuda_two_stage_cluster_mode(n, cluster, monkeypatch):
    """Thread-block clusters sharing one matrix (what a handful of very large AEs get)."""
    monkeypatch.setenv("SA_GPU_TS_CLUSTER", str(cluster))
    test_cuda_two_stage_matches_the_statement(n, False)


@pytest.mark.gpu
@pytest.mark.parametrize("n,graded", [(3, False), (20, False), (33, False), (34, False), (35, False), (64, False),
                                      (65, False), (66, False), (97, True), (130, False), (257, True),
                                      (300, False), (700, True), (1203, False)])
def test_cuda_two_stage_matches_the_statement(n, graded):
    from saamge_b200 import cabi

    ctx = cabi.Context(0)
    try:
        A = _sym(n, 7 * n + 1, graded)
        T, tau1, d, e = _gpu_reduce(ctx, A)
        b = 32
        w0 = np.linalg.eigvalsh(A)
        # stage 1: same band (the CUDA kernel follows the statement operation by operation, so
        # entries agree to roundoff, not only the spectrum)
        Tr, tau1r = ts.stage1(A, b)
        band_g = np.zeros((n, n))
        band_r = np.zeros((n, n))
        for j in range(n):
            hi = min(n, j + b + 1)
            band_g[j:hi, j] = T[j:hi, j]
            band_r[j:hi, j] = Tr[j:hi, j]
        Bg = np.tril(band_g) + np.tril(band_g, -1).T
        assert np.abs(np.linalg.eigvalsh(Bg) - w0).max() <= 1e-12 * max(1, n) * np.abs(w0).max(), "band spectrum"
        if n <= 300:
            assert np.abs(band_g - band_r).max() <= 1e-10, "band entries"
            assert np.abs(tau1 - tau1r).max() <= 1e-10, "tau1"
        # stage 2: tridiagonal matrix with the spectrum of A
        w = eigh_tridiagonal(d, e[:n - 1], eigvals_only=True) if n > 1 else d
        assert np.abs(w - w0).max() <= 1e-12 * max(1, n) * np.abs(w0).max(), "tridiagonal spectrum"
        # back-transformation of the lowest eigenvectors
        k = min(4, n)
        wv, Y = eigh_tridiagonal(d, e[:n - 1], select="i", select_range=(0, k - 1))
        Yf = np.asfortranarray(Y)
        dp = ctypes.POINTER(ctypes.c_double)
        ctx.lib.sa_gpu_debug_twostage_back.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp]
        rc = ctx.lib.sa_gpu_debug_twostage_back(ctx.h, n, k, Yf.ctypes.data_as(dp))
        assert rc == 0, ctx.lib.sa_gpu_last_error()
        for j in range(k):
            z = Yf[:, j]
            assert np.linalg.norm(A @ z - wv[j] * z) <= 1e-12 * n * np.abs(w0).max(), "residual"
            assert abs(np.linalg.norm(z) - 1.0) <= 1e-12 * n
    finally:
        ctx.close()
