"""Cholesky + shift-invert subspace iteration (saamge_b200/csrc/cholsi.cu) against
numpy.linalg.eigh: the eigenpairs with lambda <= theta of dense symmetric matrices with spectrum
in [0, 1] -- the large-AE form of xpacks_calc_lower_eigens_dense (amg/src/xpacks.cpp:222-314).
Tolerances are those of the hierarchy parity tests: eigenvalues 1e-10, eigenspaces 1e-8."""
import ctypes

import numpy as np
import pytest

import saamge_b200 as sab

pytestmark = pytest.mark.gpu
K = 8
THETA = 0.003


def _matrices(nmats, n, seed, small):
    """Q diag(lam) Q^T; `small(rng)` gives the eigenvalues near / below theta, the rest lies in
    [0.02, 1]."""
    rng = np.random.default_rng(seed)
    out, lams = [], []
    for _ in range(nmats):
        lo = np.sort(np.asarray(small(rng), dtype=float))
        rest = np.sort(rng.uniform(0.02, 1.0, n - len(lo)))
        lam = np.concatenate([lo, rest])
        Q, _r = np.linalg.qr(rng.standard_normal((n, n)))
        A = (Q * lam) @ Q.T
        out.append(0.5 * (A + A.T))
        lams.append(lam)
    return out, lams


def _run(mats, theta=THETA):
    g = sab.gpu_lib()
    h = sab.host_lib()
    h.sa_drv_ctx.restype = ctypes.c_void_p
    ctx = ctypes.c_void_p(h.sa_drv_ctx())
    n = mats[0].shape[0]
    A = np.ascontiguousarray(np.stack([m.T for m in mats]))  # column-major blocks
    info = np.zeros(2 * len(mats), dtype=np.int32)
    lam = np.zeros(K * len(mats))
    X = np.zeros(len(mats) * n * K)
    dp = ctypes.POINTER(ctypes.c_double)
    g.sa_gpu_debug_cholsi.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp, ctypes.c_double,
                                      ctypes.POINTER(ctypes.c_int), dp, dp]
    rc = g.sa_gpu_debug_cholsi(ctx, len(mats), n, A.ctypes.data_as(dp), theta,
                               info.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), lam.ctypes.data_as(dp),
                               X.ctypes.data_as(dp))
    assert rc == 0, g.sa_gpu_last_error()
    return info.reshape(-1, 2), lam.reshape(-1, K), X.reshape(len(mats), K, n)


@pytest.mark.parametrize("nmats,n", [(3, 250), (2, 729), (5, 1203), (160, 300)])
def test_lower_eigenpairs_match_eigh(nmats, n):
    mats, _ = _matrices(nmats, n, 7 + n, lambda rng: [0.0, 1e-4, 2.0e-3, 3.1e-3, 6e-3][: 1 + rng.integers(1, 5)])
    info, lam, X = _run(mats)
    worst_l, worst_v = 0.0, 0.0
    for b, A in enumerate(mats):
        w, V = np.linalg.eigh(A)
        m = int((w <= THETA).sum())
        assert info[b, 0] == m, (b, info[b], w[:6])
        assert np.max(np.abs(lam[b, :m] - w[:m])) <= 1e-10
        worst_l = max(worst_l, float(np.max(np.abs(lam[b, :m] - w[:m]))))
        Y = X[b, :m].T
        assert np.max(np.abs(Y.T @ Y - np.eye(m))) <= 1e-10
        # sine of the largest principal angle between the two m-dimensional spaces
        s = np.linalg.svd(Y - V[:, :m] @ (V[:, :m].T @ Y), compute_uv=False)
        worst_v = max(worst_v, float(s[0]))
        assert s[0] <= 1e-8, (b, s[0])
    print("n %d x %d: max |dlambda| %.1e, max sin(angle) %.1e, iterations %s" % (
        n, nmats, worst_l, worst_v, sorted(set(info[:, 1].tolist()))))


def test_failures_are_reported():
    """All K Ritz values below theta (-2: Ritz values bound the K lowest eigenvalues from above)
    and an indefinite matrix (-1): the caller falls back to the two-stage tridiagonalisation."""
    mats, _ = _matrices(1, 200, 3, lambda rng: np.linspace(0, 2.9e-3, 9))
    info, _lam, _X = _run(mats)
    assert info[0, 0] == -2, info
    mats, _ = _matrices(1, 200, 4, lambda rng: [0.0])
    info, _lam, _X = _run([mats[0] - 0.5 * np.eye(200)])
    assert info[0, 0] == -1, info
