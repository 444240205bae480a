"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and through size-independent
properties at larger sizes.  Tolerances are the north star's: maps/patterns bit-exact
(patterns up to numerical zeros, see parity._pattern_mismatch), eigenvalues 1e-10,
eigenspaces 1e-8 (sine of the largest principal angle), coarse operator 1e-9, PCG
iterations +-1."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

import fixtures
import golden_util
import oracle_util as ou
import parity
import saamge_b200 as sab
from saamge_b200 import cabi

pytestmark = pytest.mark.gpu


def _run(pr, p):
    Ho = ou.orc_build(pr, p)
    ito = ou.orc_pcg(Ho)
    Hg = sab.ml_build(pr, p)
    itg = sab.ml_pcg(Hg)
    sab.ml_download(Hg)
    return Hg, Ho, itg, ito


CONFIGS = [
    # dim, n, order, coef, levels, first_epa, epa, nu_pro, kind, block, cblock
    (3, 16, 1, 1, 3, 64, 8, 0, 1, 4, 2),
    (3, 16, 1, 1, 3, 64, 8, 1, 0, 4, 2),
    (2, 64, 1, 0, 2, 64, 64, 0, 0, 4, 2),
    (2, 64, 1, 1, 3, 64, 16, 1, 0, 4, 2),
    (3, 12, 2, 1, 2, 27, 8, 0, 1, 3, 2),
    (3, 24, 1, 1, 3, 52, 24, 0, 0, 4, 2),
    (3, 16, 1, 0, 3, 64, 8, 0, 1, 4, 2),
    (2, 6, 1, 1, 2, 4, 4, 0, 1, 2, 2),
    (3, 4, 1, 1, 2, 64, 8, 1, 1, 4, 2),  # a single AE
]


# configurations whose MIS bases differ from LAPACK's by a rotation on some level (degenerate
# singular values): number of levels that can be compared entry by entry
ROTATION_GAUGE_LEVELS = {}


@pytest.mark.parametrize("cfg", CONFIGS)
def test_parity_vs_oracle(cfg):
    dim, n, order, coef, levels, fepa, epa, nupro, kind, blk, cblk = cfg
    p = sab.default_params(num_levels=levels, first_elems_per_agg=fepa, elems_per_agg=epa,
                           first_nu_pro=nupro, nu_pro=nupro, partition_kind=kind,
                           block=(blk, blk, blk), coarse_block=cblk)
    pr = sab.Problem(dim, n, order=order, coef_kind=coef)
    pr.partition(p)
    Hg, Ho, itg, ito = _run(pr, p)
    # every coarsening is compared entry by entry -- no silently skipped levels -- except where the
    # configuration is known to produce a rotation gauge (see parity.compare_hierarchies)
    expect = ROTATION_GAUGE_LEVELS.get(cfg, levels - 1)
    res = parity.compare_hierarchies(Hg, Ho, expect_levels=expect)
    for l, m in enumerate(res):
        parity.assert_level_ok(m, l)
    for m in res[-1].get("coarser_invariants", []):
        assert m["maps_mismatch"] == [] and m["ND"][0] == m["ND"][1], m
        assert m.get("operator_spectrum_err", 0.0) <= 1e-9, m
    assert abs(itg - ito) <= 1
    assert itg > 0
    rg, ro = Hg.scalar("pcg.final_res_norm"), Ho.scalar("pcg.final_res_norm")
    assert abs(rg - ro) <= 1e-6 * max(ro, 1e-300) + 1e-12
    bg, bo = Hg.get("pcg.brr"), Ho.get("pcg.brr")
    k = min(len(bg), len(bo), 3)
    assert np.allclose(bg[:k], bo[:k], rtol=1e-6)
    for h in (Hg, Ho):
        h.close()
    pr.close()


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_gpu_matches_golden(name):
    pr, p = golden_util.make_golden.build(name)
    Hg = sab.ml_build(pr, p)
    sab.ml_pcg(Hg)
    sab.ml_download(Hg)
    golden_util.check_against_golden(Hg, p, golden_util.load(name))
    Hg.close()
    pr.close()


@pytest.mark.parametrize("order,levels,pinned", [(1, 3, 3), (2, 2, 4), (1, 2, 3)])
def test_reference_ctest_pins_on_gpu(order, levels, pinned):
    pr, p = fixtures.mltest_problem(order, levels)
    Hg, Ho, itg, ito = _run(pr, p)
    assert abs(itg - pinned) <= 1 and abs(itg - ito) <= 1
    res = parity.compare_hierarchies(Hg, Ho, expect_levels=levels - 1)
    # The fixture injects an all-ones vector that is not an eigenvector; with the 1e6
    # checkerboard its MIS restrictions have singular values down to ~1e-8 sigma_0, so
    # the kept singular vectors are only determined to eps / 1e-8 (in LAPACK as well).
    for l, m in enumerate(res):
        parity.assert_level_ok(m, l, space_tol=1e-6, ac_tol=1e-7)
    for h in (Hg, Ho):
        h.close()
    pr.close()


def test_atleast_one_fallback_and_tiny_theta():
    """theta so small that no eigenvalue qualifies: one vector per AE is still returned
    (xpacks_calc_lower_eigens_dense, amg/src/xpacks.cpp:270-288)."""
    p = sab.default_params(num_levels=2, first_elems_per_agg=64, first_theta=1e-30, theta=1e-30,
                           partition_kind=1, block=(4, 4, 4))
    pr = sab.Problem(3, 8, coef_kind=1)
    pr.partition(p)
    Hg, Ho, itg, ito = _run(pr, p)
    assert np.array_equal(Hg.get("ae_m", 0), Ho.get("ae_m", 0))
    assert np.all(Hg.get("ae_m", 0) >= 1)
    m, _ = parity.compare_level(Hg, Ho, 0)
    parity.assert_level_ok(m, 0)
    pr.close()


@pytest.mark.parametrize("env", [
    {"SA_GPU_LARGE_PATH": "twostage"},  # two-stage tridiagonalisation (the fallback of cholsi.cu) for every large AE
    {"SA_GPU_LARGE_PATH": "coop"},      # round 1's one-stage cooperative kernel (k_tridiag_coop_sym)
    {"SA_GPU_LARGE_PATH": "coop", "SA_GPU_COOP_SYM": "0"},  # ... its full-matrix form (k_tridiag_coop)
    {"SA_GPU_SQUARE_TILE": "1"},     # square shared-memory tile kernel (k_at_smem)
    {"SA_GPU_SMALL_PATH": "packed"}, # round 1's packed shared-memory kernel instead of k_tridiag_reg
    {"SA_GPU_NO_ASYNC_ALLOC": "1"},  # plain cudaMalloc instead of the stream-ordered pool
    {"SA_GPU_COARSE_BLOCKED_MIN": "1"},  # blocked coarsest factorisation even for tiny n
])
def test_parity_of_alternative_kernel_paths(env):
    """The switches are read once per process, so each variant runs in a fresh interpreter:
    3-level 24^3 (coarse-level AEs take the large-matrix path) against the oracle."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(here, "run_parity.py"),
                          "3", "24", "1", "1", "3", "52", "24", "0", "0"],
                         env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "PARITY OK" in out.stdout, out.stdout[-2000:]


def test_large_AE_chunk_falls_back_to_the_tridiagonalisation():
    """theta = 0.05: every coarse-level AE of the 24^3 three-level problem has 8 or more
    eigenvalues below theta, which the 8-vector subspace iteration of cholsi.cu reports (-2) -- the
    chunk is redone by the two-stage tridiagonalisation + Sturm counts, and the hierarchy still
    matches the oracle (counts exactly, operators to 1e-8 with ~10 vectors per AE)."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ)
    e.update({"SA_TEST_THETA": "0.05", "SA_TEST_AC_TOL": "1e-8", "SA_GPU_SPECTRAL_DEBUG": "1"})
    out = subprocess.run([sys.executable, os.path.join(here, "run_parity.py"),
                          "3", "24", "1", "1", "3", "52", "24", "0", "0"],
                         env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "PARITY OK" in out.stdout, out.stdout[-2000:]
    import re

    m = re.search(r"\[cholsi\] chunk of (\d+) AEs .*: (\d+) not done", out.stderr)
    assert m and int(m.group(2)) > 0, out.stderr[-1000:]


# ---------------------------------------------------------------- direct C ABI


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Context(0)
    yield c
    c.close()


def _problem(n=8, dim=3, blk=4, coef=1):
    p = sab.default_params(num_levels=2, first_elems_per_agg=blk ** dim, partition_kind=1, block=(blk, blk, blk))
    pr = sab.Problem(dim, n, coef_kind=coef)
    pr.partition(p)
    return pr, p


def _host_AE_matrix(pr, part):
    """A_AE = sum of the AE's element matrices (what agg_build_AE_stiffm computes)."""
    AEd_I, AEd_J = pr.get("AE_to_dof.I"), pr.get("AE_to_dof.J")
    AEe_I, AEe_J = pr.get("AE_to_elem.I"), pr.get("AE_to_elem.J")
    e2d_I, e2d_J = pr.get("elem_to_dof.I"), pr.get("elem_to_dof.J")
    ne = int(pr.scalar("ne"))
    el = pr.get("elmat")
    dofs = AEd_J[AEd_I[part] : AEd_I[part + 1]]
    loc = {g: i for i, g in enumerate(dofs)}
    M = np.zeros((len(dofs), len(dofs)))
    for e in AEe_J[AEe_I[part] : AEe_I[part + 1]]:
        ed = e2d_J[e2d_I[e] : e2d_I[e + 1]]
        K = el[e * ne * ne : (e + 1) * ne * ne].reshape(ne, ne).T
        idx = [loc[g] for g in ed]
        M[np.ix_(idx, idx)] += K
    return M


def test_abi_build_AE_stiff_both_modes(ctx):
    pr, p = _problem()
    flags = pr.get("agg_flags")
    for with_global in (False, True):
        lev = cabi.Level(ctx, pr, with_global=with_global)
        for part in (0, lev.nparts - 1):
            got = lev.build_AE_stiff(part)
            ref = _host_AE_matrix(pr, part)
            if with_global:
                # essential dofs: rows/cols copied from the eliminated global matrix
                dofs = pr.get("AE_to_dof.J")[pr.get("AE_to_dof.I")[part] : pr.get("AE_to_dof.I")[part + 1]]
                ess = (flags[dofs] & 2) != 0
                A = sp.csr_matrix((pr.get("A.A"), pr.get("A.J"), pr.get("A.I")))
                sub = A[dofs][:, dofs].toarray()
                iface = (flags[dofs] & 1) != 0
                for i in range(len(dofs)):
                    for j in range(len(dofs)):
                        if ess[i] or ess[j]:
                            if i == j and iface[i]:
                                continue  # essential interface diagonal is re-assembled
                            ref[i, j] = sub[i, j]
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
        lev.close()
    pr.close()


def test_abi_spectral_residuals_and_counts(ctx):
    pr, p = _problem(n=12, blk=4)
    lev = cabi.Level(ctx, pr)
    lev.local_spectral(0.003)
    m, ev, Z, D = lev.spectral()
    n = lev.ae_sizes()
    zo = np.concatenate([[0], np.cumsum(m.astype(np.int64) * n)])
    eo = np.concatenate([[0], np.cumsum(m)])
    do = np.concatenate([[0], np.cumsum(n)])
    for i in range(lev.nparts):
        A = lev.build_AE_stiff(i)
        Di = D[do[i] : do[i + 1]]
        Zi = Z[zo[i] : zo[i + 1]].reshape(m[i], n[i]).T
        lam = ev[eo[i] : eo[i + 1]]
        R = A @ Zi - (Di[:, None] * Zi) * lam[None, :]
        assert np.abs(R).max() <= 1e-12 * np.abs(A).max()
        assert np.allclose(Zi.T @ (Di[:, None] * Zi), np.eye(m[i]), atol=1e-12)
        # count: number of generalized eigenvalues <= theta (at least one)
        w = np.linalg.eigvalsh((A / np.sqrt(Di)[:, None]) / np.sqrt(Di)[None, :])
        assert m[i] == max(1, int(np.sum(w <= 0.003)))
        assert np.allclose(lam, w[: m[i]], atol=1e-12)
    lev.close()
    pr.close()


def test_abi_sparse_kernels(ctx):
    pr, p = _problem(n=10, dim=3, blk=5)
    lev = cabi.Level(ctx, pr)
    A = sp.csr_matrix((pr.get("A.A"), pr.get("A.J"), pr.get("A.I")))
    rng = np.random.default_rng(3)
    x = rng.normal(size=A.shape[0])
    assert np.allclose(lev.spmv(0, x), A @ x, rtol=1e-13, atol=1e-13)
    dneg = lev.build_Dinv_neg()
    dg = A.diagonal()
    ref = -1.0 / (np.sqrt(np.abs(dg)) * (abs(A) @ (1.0 / np.sqrt(np.abs(dg)))))
    assert np.allclose(dneg, ref, rtol=1e-13)
    # smpr_compute_poly (amg/inc/smpr.hpp:319-339)
    roots = np.array([1.0, 0.6, 0.3])
    b = rng.normal(size=A.shape[0])
    x0 = rng.normal(size=A.shape[0])
    xr = x0.copy()
    for r in roots:
        xr = xr + (1.0 / r) * (dneg * (A @ xr - b))
    assert np.allclose(lev.poly_smooth(b, x0, roots), xr, rtol=1e-12, atol=1e-12)
    # tentative P, smoothed P, transpose, RAP
    lev.local_spectral(0.003)
    ncd, NDc = lev.tentative_P(1)
    Pt = lev.csr(1)
    assert Pt.shape == (A.shape[0], NDc) and NDc == ncd.sum()
    assert abs(Pt.T @ Pt - sp.identity(NDc)).max() <= 1e-12
    sa_roots = np.array([np.sin(np.pi / 3.0) ** 2])  # smpr_sa_poly_roots, nu = 1
    lev.smooth_P(sa_roots)
    P = lev.csr(2)
    Pref = (sp.identity(A.shape[0]) + sp.diags(dneg / sa_roots[0]) @ A) @ Pt
    assert abs(P - Pref).max() <= 1e-13
    R = lev.csr(3)
    assert abs(R - P.T).max() == 0.0
    assert np.all(np.diff(R.indptr) >= 0) and R.has_sorted_indices
    lev.rap()
    Ac = lev.csr(4)
    assert abs(Ac - P.T @ A @ P).max() <= 1e-12 * abs(Ac).max()
    assert Ac.has_sorted_indices
    # V-cycle is a symmetric linear operator; PCG converges
    S = cabi.Solver(ctx, [lev], 3)
    r1, r2 = rng.normal(size=A.shape[0]), rng.normal(size=A.shape[0])
    z1, z2 = S.vcycle(r1), S.vcycle(r2)
    assert abs(z1 @ r2 - r1 @ z2) <= 1e-10 * abs(z1 @ r2)
    z12 = S.vcycle(r1 + 2 * r2)
    assert np.allclose(z12, z1 + 2 * z2, rtol=1e-10, atol=1e-12)
    xs, it, hist = S.pcg(pr.get("b"))
    assert 0 < it < 20 and hist[-1] < 1e-12 * hist[0]
    assert np.linalg.norm(A @ xs - pr.get("b")) <= 1e-5 * np.linalg.norm(pr.get("b"))
    S.close()
    lev.close()
    pr.close()


def test_abi_pipelined_upload_is_bit_identical(ctx):
    """desc.async_upload: the eigen stage starts on the first AEs while later slabs of the
    operator / element blocks are still in flight; results must not depend on it."""
    pr, p = _problem(n=16, blk=4)
    ref = cabi.Level(ctx, pr)
    ref.local_spectral(0.003)
    m0, ev0, Z0, D0 = ref.spectral()
    for first_call in ("spectral", "spmv", "wait"):
        lev = cabi.Level(ctx, pr, async_upload=True)
        if first_call == "spmv":  # any other entry point waits for the whole upload
            x = np.ones(lev.ND)
            assert np.array_equal(lev.spmv(0, x), ref.spmv(0, x))
        elif first_call == "wait":
            assert lev.lib.sa_gpu_level_upload_wait(lev.h) == 0
        lev.local_spectral(0.003)
        m1, ev1, Z1, D1 = lev.spectral()
        assert np.array_equal(m0, m1) and np.array_equal(ev0, ev1)
        assert np.array_equal(Z0, Z1) and np.array_equal(D0, D1)
        lev.close()
    # partial AE range on a pipelined level (the sharded path)
    lev = cabi.Level(ctx, pr, async_upload=True)
    h = lev.nparts // 2
    check = lev.lib.sa_gpu_local_spectral(lev.h, ctypes.c_double(0.003), h, lev.nparts, 0)
    assert check == 0
    lev.close()
    ref.close()
    pr.close()


def test_abi_errors_are_reported(ctx):
    pr, p = _problem()
    lev = cabi.Level(ctx, pr)
    ncd = np.zeros(lev.num_mises, dtype=np.int32)
    rc = lev.lib.sa_gpu_tentative_P(lev.h, 1, ncd.ctypes.data_as(cabi._i), None)
    assert rc != 0 and b"sa_gpu_local_spectral" in lev.lib.sa_gpu_last_error()
    assert lev.lib.sa_gpu_rap(lev.h) != 0
    lev.close()
    pr.close()


def test_properties_at_scale():
    """64^3-class sizes are too slow for the CPU oracle inside a test; check the
    size-independent properties instead (32^3 here keeps the GPU suite short)."""
    p = sab.default_params(num_levels=3, first_elems_per_agg=64, elems_per_agg=64,
                           partition_kind=1, block=(4, 4, 4), coarse_block=4)
    pr = sab.Problem(3, 32, coef_kind=1)
    pr.partition(p)
    Hg = sab.ml_build(pr, p)
    it = sab.ml_pcg(Hg)
    sab.ml_download(Hg)
    assert 0 < it < 30
    A = sp.csr_matrix((pr.get("A.A"), pr.get("A.J"), pr.get("A.I")))
    x = Hg.get("pcg.x")
    b = pr.get("b")
    assert np.linalg.norm(A @ x - b) <= 1e-5 * np.linalg.norm(b)
    Aprev = A
    for l in range(2):
        P = Hg.csr("interp", l)
        Pt = Hg.csr("tent_interp", l)
        Ac = Hg.csr("Ac", l)
        assert abs(Pt.T @ Pt - sp.identity(Pt.shape[1])).max() <= 1e-11
        G = (P.T @ Aprev @ P).tocsr()
        assert abs(G - Ac).max() <= 1e-11 * abs(Ac).max()
        assert abs(Ac - Ac.T).max() <= 1e-11 * abs(Ac).max()
        m = Hg.get("ae_m", l)
        assert np.all(m >= 1)
        Aprev = Ac
    Hg.close()
    pr.close()


def test_properties_at_full_size():
    """BASELINE configs[2] at full size (128^3, ~40k METIS AEs, 4 levels): the CPU oracle would
    need minutes, so only size-independent properties are checked: orthonormal tentative
    prolongators, symmetric Galerkin operators that equal P^T A P, at least one vector per AE,
    eigenvalues below theta, converging PCG."""
    p = sab.default_params(num_levels=4, first_elems_per_agg=52, elems_per_agg=64,
                           partition_kind=2, block=(32, 32, 32))
    pr = sab.Problem(3, 128, coef_kind=1)
    na = pr.partition(p)
    assert na == 40320
    Hg = sab.ml_build(pr, p)
    it = sab.ml_pcg(Hg)
    sab.ml_download(Hg)
    assert 0 < it <= 25
    A = sp.csr_matrix((pr.get("A.A"), pr.get("A.J"), pr.get("A.I")))
    x, b = Hg.get("pcg.x"), pr.get("b")
    assert np.linalg.norm(A @ x - b) <= 1e-5 * np.linalg.norm(b)
    Aprev = A
    for l in range(3):
        Pt = Hg.csr("tent_interp", l)
        P = Hg.csr("interp", l)
        Ac = Hg.csr("Ac", l)
        assert abs(Pt.T @ Pt - sp.identity(Pt.shape[1])).max() <= 1e-10
        assert abs(Ac - Ac.T).max() <= 1e-10 * abs(Ac).max()
        if l == 2:  # (the finer triple products take minutes in scipy; they are covered at 32^3)
            G = (P.T @ Aprev @ P).tocsr()
            assert abs(G - Ac).max() <= 1e-10 * abs(Ac).max()
        m = Hg.get("ae_m", l)
        assert np.all(m >= 1)
        ev = Hg.get("evals", l) if l == 0 else None
        if ev is not None:
            assert np.all(ev[np.isfinite(ev)] <= 0.003 + 1e-12) or np.sum(ev > 0.003) <= np.sum(m == 1)
        Aprev = Ac
    Hg.close()
    pr.close()


def test_contexts_can_be_reopened_and_share_the_allocation_stream():
    """the stream-ordered allocator must survive a context being destroyed and
    another one created (test fixtures do exactly that), and a second live context must not
    allocate on a stream its kernels do not run on."""
    pr, p = _problem()
    for _ in range(3):
        c1 = cabi.Context(0)
        c2 = cabi.Context(0)  # alive at the same time: shares c1's main stream
        l1, l2 = cabi.Level(c1, pr), cabi.Level(c2, pr)
        l1.local_spectral(0.003)
        l2.local_spectral(0.003)
        a, b = l1.spectral(), l2.spectral()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        l1.close()
        c1.close()  # the survivor keeps working
        l2.local_spectral(0.003)
        l2.close()
        c2.close()
    pr.close()


def test_smoothed_prolongator_drop_tolerance():
    """AltThreshold (amg/src/interp.cpp:86-170), the drop_tol branch of interp_smooth: entries of
    the smoothed P with |p| <= drop_tol are dropped before RAP -- same pattern and values as the
    oracle's restatement, and the tolerance really removes entries."""
    kw = dict(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, first_nu_pro=1, nu_pro=1,
              partition_kind=1, block=(4, 4, 4), coarse_block=2)
    pr = sab.Problem(3, 16, coef_kind=1)
    p0 = sab.default_params(**kw)
    pr.partition(p0)
    H0 = sab.ml_build(pr, p0)
    sab.ml_download(H0)
    nnz0 = H0.csr("interp", 0).nnz
    H0.close()
    p = sab.default_params(smooth_drop_tol=2e-2, **kw)
    Hg, Ho, itg, ito = _run(pr, p)
    assert Hg.csr("interp", 0).nnz < nnz0
    assert np.abs(Hg.csr("interp", 0).data).min() > 2e-2
    for l, m in enumerate(parity.compare_hierarchies(Hg, Ho, expect_levels=2)):
        parity.assert_level_ok(m, l)
    assert itg > 0 and abs(itg - ito) <= 1, (itg, ito)
    for h in (Hg, Ho):
        h.close()
    pr.close()


@pytest.mark.parametrize("resmooth", [True, False])
def test_operator_update_without_eigensolves(resmooth):
    """adapt_update_operators (amg/src/adapt.cpp:171-216): a new operator with the same pattern
    (here A + 0.3 diag(A): a reaction term switched on) -- smoothers, smoothed prolongators (from
    the KEPT tentative ones) and Galerkin operators of every level follow it on the device exactly
    as in the oracle; the spectral data are untouched; PCG on the new operator converges."""
    import scipy.sparse as sp

    p = sab.default_params(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, first_nu_pro=1, nu_pro=1,
                           partition_kind=1, block=(4, 4, 4), coarse_block=2)
    pr = sab.Problem(3, 16, coef_kind=1)
    pr.partition(p)
    Hg, Ho, _itg, _ito = _run(pr, p)
    tent_before = Hg.csr("tent_interp", 0).copy()
    interp_before = Hg.csr("interp", 0).copy()
    I, J, A = pr.get("A.I"), pr.get("A.J"), pr.get("A.A").copy()
    M = sp.csr_matrix((A, J, I))
    diag = M.diagonal()
    rows = np.repeat(np.arange(len(I) - 1), np.diff(I))
    A2 = A + 0.3 * diag[rows] * (rows == J)
    sab.set_operator_values(pr, A2)
    assert sab.ml_update_operators(Hg, resmooth) == 0
    assert ou.orc_update_operators(Ho, resmooth) == 0
    itg, ito = sab.ml_pcg(Hg), ou.orc_pcg(Ho)
    sab.ml_download(Hg)
    # stage by stage against the oracle (gauge-aware: P_g = S P_o Q, Ac_g = Q^T Ac_o Q, Dinv_neg)
    for l, m in enumerate(parity.compare_hierarchies(Hg, Ho, expect_levels=2)):
        parity.assert_level_ok(m, l)
    assert abs(Hg.csr("tent_interp", 0) - tent_before).max() == 0.0
    changed = abs(Hg.csr("interp", 0) - interp_before).max()
    assert (changed > 1e-6) if resmooth else (changed == 0.0)
    assert itg > 0 and abs(itg - ito) <= 1, (itg, ito)
    x = Hg.get("pcg.x")
    b = pr.get("b")
    M2 = sp.csr_matrix((A2, J, I))
    assert np.linalg.norm(M2 @ x - b) <= 1e-5 * np.linalg.norm(b)
    for h in (Hg, Ho):
        h.close()
    pr.close()


def test_correct_nullspace_coarsest_solver():
    """CorrectNullspace (amg/src/solve.cpp:52-164; ml_produce_hierarchy_from_level,
    amg/src/ml.cpp:225-235): below the last spectral level a two-grid cycle whose prolongator is
    the scaling P (one column per MIS: the coarse representation of the constant,
    amg/src/contrib.cpp:655-668, amg/src/interp.cpp:842-909).  The device builds it as one more level
    of the cycle; against the oracle's restatement (dgels per MIS): the scaling P up to the sign
    gauge of the coarse basis, its Galerkin operator (gauge invariant) and the PCG run."""
    # Two-level: the last spectral level is the finest one, where "1" is the constant FUNCTION.  (On
    # a coarser level the reference's b = 1.0 is the all-ones COEFFICIENT vector in whatever signs
    # LAPACK gave the coarse basis functions: not a gauge-invariant quantity, there is nothing to
    # compare entry by entry -- see the three-level run below.)
    p = sab.default_params(num_levels=2, first_elems_per_agg=64, first_nu_pro=1, nu_pro=1,
                           partition_kind=1, block=(4, 4, 4), correct_nullspace=1)
    pr = sab.Problem(3, 16, coef_kind=1)
    pr.partition(p)
    Hg, Ho, itg, ito = _run(pr, p)
    for l, m in enumerate(parity.compare_hierarchies(Hg, Ho, expect_levels=1)):
        parity.assert_level_ok(m, l)
    Pg, Po = Hg.csr("cn_P"), Ho.csr("cn_P")
    assert Pg.shape == Po.shape and Pg.shape[0] == Hg.csr("Ac", 0).shape[0]
    assert np.array_equal(Pg.indices, Po.indices) and np.array_equal(Pg.indptr, Po.indptr)
    # (the entries themselves are gauge dependent: a MIS basis is fixed up to an orthogonal Q_mis, and
    # the representation of the constant turns with it, x_g = Q_mis^T x_o -- its norm per MIS and the
    # Galerkin operator P_s^T Ac P_s do not)
    nrm = lambda P: np.sqrt(np.asarray(P.multiply(P).sum(axis=0)).ravel())
    assert np.allclose(nrm(Pg), nrm(Po), atol=1e-12)
    # every column has unit norm (the representation is normalised per MIS)
    assert np.allclose(np.asarray(Pg.multiply(Pg).sum(axis=0)).ravel(), 1.0, atol=1e-12)
    Ag, Ao = Hg.csr("cn_Ac"), Ho.csr("cn_Ac")
    assert Ag.shape == Ao.shape == (Pg.shape[1], Pg.shape[1])
    assert abs(Ag - Ao).max() <= 1e-9 * abs(Ao).max()
    assert itg > 0 and abs(itg - ito) <= 1, (itg, ito)
    bg, bo = Hg.get("pcg.brr"), Ho.get("pcg.brr")
    k = min(len(bg), len(bo), 3)
    assert np.allclose(bg[:k], bo[:k], rtol=1e-6)
    for h in (Hg, Ho):
        h.close()
    # three levels: the cycle with the extra level converges like the oracle's
    p3 = sab.default_params(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, first_nu_pro=1, nu_pro=1,
                            partition_kind=1, block=(4, 4, 4), coarse_block=2, correct_nullspace=1)
    Hg, Ho, itg, ito = _run(pr, p3)
    assert Hg.csr("cn_P").shape == Ho.csr("cn_P").shape
    assert itg > 0 and abs(itg - ito) <= 2, (itg, ito)
    for h in (Hg, Ho):
        h.close()
    pr.close()
