"""The drop-in boundary beyond the batched fast path (SURVEY 8b plug points): a user provider that
implements only ElementMatrixProvider::GetMatrix (amg/inc/elmat.hpp:53-77), a user smoother with
the smpr_ft signature (amg/inc/smpr.hpp:59-60), and ElementMatrixParallelCoarse::GetMatrix /
BuildAEStiff on a coarse level (amg/src/elmat.cpp:105-195) -- all through the C++ host mirror."""
import ctypes

import numpy as np
import pytest

import saamge_b200 as sab

pytestmark = pytest.mark.gpu


def _build(pr, p, flags):
    h = sab.host_lib()
    h.sa_drv_ml_build_user.restype = ctypes.c_void_p
    h.sa_drv_ml_build_user.argtypes = [ctypes.c_void_p, ctypes.POINTER(sab.Params), ctypes.c_int, ctypes.c_int]
    return sab.Hierarchy(h.sa_drv_ml_build_user(pr.handle, ctypes.byref(p), 0, flags))


def test_getmatrix_only_provider_and_user_smoother():
    p = sab.default_params(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, partition_kind=1,
                           block=(4, 4, 4), coarse_block=2)
    pr = sab.Problem(3, 12, coef_kind=1)
    pr.partition(p)
    ref = _build(pr, p, 0)        # ElementMatrixDenseArray: batched view
    it0 = sab.ml_pcg(ref)
    brr0 = ref.get("pcg.brr")
    usr = _build(pr, p, 1)        # GetMatrix-only provider: packed by the host mirror
    it1 = sab.ml_pcg(usr)
    # (the SpGEMM of the setup accumulates with atomics: equal to round-off, not bitwise)
    assert it1 == it0 and np.allclose(usr.get("pcg.brr"), brr0, rtol=1e-9)
    sm = _build(pr, p, 3)         # + user smoother (host callback) on the finest level
    it2 = sab.ml_pcg(sm)
    calls = sab.host_lib().sa_drv_user_smoother_calls()
    assert calls == 2 * (it2 + 1), (calls, it2)  # pre + post per V-cycle, one V-cycle per iteration + 1
    assert it2 == it0 and np.allclose(sm.get("pcg.brr"), brr0, rtol=1e-9)
    for h in (ref, usr, sm):
        h.close()
    pr.close()


def test_parallel_coarse_provider_through_the_mirror():
    p = sab.default_params(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, partition_kind=1,
                           block=(4, 4, 4), coarse_block=2)
    pr = sab.Problem(3, 12, coef_kind=1)
    pr.partition(p)
    H = sab.ml_build(pr, p)
    sab.ml_download(H)
    h = sab.host_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    h.sa_drv_coarse_provider_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp,
                                               ctypes.POINTER(ctypes.c_int), ctypes.c_int, dp,
                                               ctypes.POINTER(ctypes.c_int)]
    cel, off = H.get("celmat", 0), H.get("celmat_off", 0)
    e2d_I, e2d_J = H.get("elem_to_dof.I", 1), H.get("elem_to_dof.J", 1)
    AEe_I, AEe_J = H.get("AE_to_elem.I", 1), H.get("AE_to_elem.J", 1)
    AEd_I, AEd_J = H.get("AE_to_dof.I", 1), H.get("AE_to_dof.J", 1)
    for elno, ae in ((0, 0), (len(off) - 2, len(AEd_I) - 2)):
        ne, n = ctypes.c_int(), ctypes.c_int()
        nce = e2d_I[elno + 1] - e2d_I[elno]
        nae = AEd_I[ae + 1] - AEd_I[ae]
        M = np.zeros(nce * nce)
        A = np.zeros(nae * nae)
        rc = h.sa_drv_coarse_provider_probe(H.handle, 1, elno, M.ctypes.data_as(dp), ctypes.byref(ne), ae,
                                            A.ctypes.data_as(dp), ctypes.byref(n))
        assert rc == 0 and ne.value == nce and n.value == nae
        # GetMatrix(elno) is the device block P_e^T A_AE P_e
        assert np.array_equal(M, cel[off[elno]:off[elno + 1]])
        # BuildAEStiff(ae) = sum of the AE's coarse element matrices (agg_build_AE_stiffm)
        dofs = AEd_J[AEd_I[ae]:AEd_I[ae + 1]]
        loc = {g: i for i, g in enumerate(dofs)}
        ref = np.zeros((nae, nae))
        for e in AEe_J[AEe_I[ae]:AEe_I[ae + 1]]:
            ed = e2d_J[e2d_I[e]:e2d_I[e + 1]]
            K = cel[off[e]:off[e + 1]].reshape(len(ed), len(ed)).T
            idx = [loc[g] for g in ed]
            ref[np.ix_(idx, idx)] += K
        got = A.reshape(nae, nae).T
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    H.close()
    pr.close()


@pytest.mark.parametrize("n", [7, 60, 150, 300])
def test_eigensolver_class_on_a_given_matrix(n):
    """Eigensolver::Solve (amg/inc/spectral.hpp:91-120) through the mirror: weighted-l1 B, all
    eigenpairs of A z = lambda B z with lambda <= theta, z^T B z = 1."""
    rng = np.random.default_rng(n)
    # a graph Laplacian-like SPD-ish matrix (positive diagonal), like an AE stiffness matrix
    W = np.abs(rng.normal(size=(n, n))) * (rng.random((n, n)) < 0.2)
    W = np.triu(W, 1)
    W = W + W.T
    A = np.diag(W.sum(axis=1) + 1e-3) - W
    h = sab.host_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    h.sa_drv_eigensolver_solve.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_double, ctypes.c_int, dp, dp, dp]
    cap = n
    ev, Z, B = np.zeros(cap), np.zeros(n * cap), np.zeros(n)
    theta = 0.05
    Af = np.asfortranarray(A)
    m = h.sa_drv_eigensolver_solve(0, n, Af.ctypes.data_as(dp), theta, cap, ev.ctypes.data_as(dp),
                                   Z.ctypes.data_as(dp), B.ctypes.data_as(dp))
    dg = np.diag(A)
    Bref = (np.abs(A) * np.sqrt(dg[:, None] / dg[None, :])).sum(axis=1)
    assert np.allclose(B, Bref, rtol=1e-13)
    w = np.linalg.eigvalsh(A / np.sqrt(Bref)[:, None] / np.sqrt(Bref)[None, :])
    assert m == max(1, int((w <= theta).sum()))
    assert np.allclose(ev[:m], w[:m], atol=1e-12)
    Zm = Z[: n * m].reshape(m, n).T
    assert np.allclose(Zm.T @ (Bref[:, None] * Zm), np.eye(m), atol=1e-11)
    assert np.abs(A @ Zm - (Bref[:, None] * Zm) * ev[None, :m]).max() <= 1e-11 * np.abs(A).max()
