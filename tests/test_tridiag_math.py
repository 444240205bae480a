"""CPU unit tests of the per-thread tridiagonal routines used by the CUDA eigen-stage
(Sturm count = dstebz's bisection kernel, pivoted LU = dstein's dlagtf/dlagts)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
from scipy.linalg import eigvalsh_tridiagonal

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def th():
    out = os.path.join(tempfile.mkdtemp(), "libtridiag_host.so")
    subprocess.check_call(
        ["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "tridiag_host.cpp")]
    )
    lib = ctypes.CDLL(out)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.th_sturm_count.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double]
    lib.th_gershgorin.argtypes = [ctypes.c_int, dp, dp, dp]
    lib.th_shifted_solve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_double, dp, ctypes.c_int, ctypes.c_int]
    lib.th_hash_uniform.restype = ctypes.c_double
    lib.th_hash_uniform.argtypes = [ctypes.c_ulonglong]
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (5, 2), (64, 3), (257, 4)])
def test_sturm_count_matches_eigenvalues(th, n, seed):
    rng = np.random.default_rng(seed)
    d = rng.uniform(0, 1, n)
    e = rng.uniform(-0.5, 0.5, max(n - 1, 1))
    lam = eigvalsh_tridiagonal(d, e[: n - 1]) if n > 1 else d.copy()
    for x in np.concatenate([rng.uniform(-1, 2, 20), 0.5 * (lam[:-1] + lam[1:])]):
        assert th.th_sturm_count(n, _p(d), _p(e), float(x)) == int(np.sum(lam <= x))


def test_sturm_count_decoupled_and_zero_pivot(th):
    # e = 0 decouples the matrix; shifts equal to diagonal entries hit zero pivots
    d = np.array([0.25, 0.5, 0.5, 1.0])
    e = np.array([0.0, 0.0, 0.0])
    assert th.th_sturm_count(4, _p(d), _p(e), 0.5) == 3
    assert th.th_sturm_count(4, _p(d), _p(e), 0.25) == 1
    assert th.th_sturm_count(4, _p(d), _p(e), 0.1) == 0


def test_gershgorin_encloses_spectrum(th):
    rng = np.random.default_rng(7)
    n = 50
    d = rng.normal(size=n)
    e = rng.normal(size=n - 1)
    out = np.zeros(3)
    th.th_gershgorin(n, _p(d), _p(e), _p(out))
    lam = eigvalsh_tridiagonal(d, e)
    assert out[0] <= lam.min() and lam.max() <= out[1]
    assert np.isclose(out[2], np.max(e * e))


@pytest.mark.parametrize("n,seed,stride,lane", [(2, 0, 1, 0), (7, 1, 32, 5), (100, 2, 32, 31), (300, 3, 4, 1)])
def test_shifted_solve(th, n, seed, stride, lane):
    rng = np.random.default_rng(seed)
    d = rng.uniform(0, 1, n)
    e = rng.uniform(-0.5, 0.5, n - 1)
    shift = 0.3
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1) - shift * np.eye(n)
    b = rng.normal(size=n)
    x = b.copy()
    th.th_shifted_solve(n, _p(d), _p(e), shift, 1e-16, _p(x), stride, lane)
    assert np.linalg.norm(T @ x - b) <= 1e-10 * np.linalg.norm(b) * np.linalg.cond(T)


def test_inverse_iteration_converges_to_eigenvector(th):
    # one solve with a shift accurate to eps must give the eigenvector (dstein's premise)
    rng = np.random.default_rng(11)
    n = 80
    d = rng.uniform(0, 1, n)
    e = rng.uniform(-0.3, 0.3, n - 1)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, V = np.linalg.eigh(T)
    for j in (0, 3, n - 1):
        x = np.array([th.th_hash_uniform(i) for i in range(n)])
        for _ in range(3):
            th.th_shifted_solve(n, _p(d), _p(e), float(lam[j]), 2.2e-16 * np.abs(lam).max(), _p(x), 32, 0)
            x /= np.linalg.norm(x)
        assert min(np.linalg.norm(x - V[:, j]), np.linalg.norm(x + V[:, j])) < 1e-8
        assert np.linalg.norm(T @ x - lam[j] * x) < 1e-13


def test_hash_uniform_range(th):
    v = np.array([th.th_hash_uniform(k) for k in range(2000)])
    assert v.min() > -1 and v.max() < 1 and abs(v.mean()) < 0.1
