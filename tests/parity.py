"""Parity metrics between two hierarchies (normally: B200 path vs CPU oracle).

Gauge freedom.  Eigenvectors are defined up to sign (rotation inside multiple
eigenvalues) and the SVD basis of every MIS up to an orthogonal k x k matrix Q_mis.
What both implementations must agree on is therefore: integer maps and sparsity
patterns (bit-exact), eigenvalues, the eigen*spaces* per AE, the column *space* of
the tentative prolongator per MIS, and every operator after the change of coarse
basis Q = blockdiag(Q_mis):  P_g = S P_o Q,  Ac_g = Q^T Ac_o Q, where S is the
(diagonal, +-1) basis change of the previous level (identity on the finest level).
Tolerances are the ones BASELINE.json's north_star states.
"""
import numpy as np
import scipy.sparse as sp

MAPS = [
    "partitioning",
    "elem_to_dof.I",
    "elem_to_dof.J",
    "AE_to_elem.I",
    "AE_to_elem.J",
    "AE_to_dof.I",
    "AE_to_dof.J",
    "dof_to_AE.I",
    "dof_to_AE.J",
    "dof_id_inAE",
    "agg_flags",
    "mises",
    "mises_size",
    "mis_to_dof.I",
    "mis_to_dof.J",
    "mis_to_AE.I",
    "mis_to_AE.J",
    "AE_to_mis.I",
    "AE_to_mis.J",
]


def _pattern_mismatch(Mg, Mo):
    """Largest |value| (relative to the largest entry) sitting at a position that is
    structurally present in only one of the two matrices.  0 means bit-identical
    patterns; a value below the entry tolerance means the patterns agree up to
    numerical zeros (LAPACK leaves round-off dust where the Jacobi SVD leaves exact
    zeros, and the reference drops exact zeros, amg/src/contrib.cpp:188)."""
    Mg = Mg.tocsr()
    Mo = Mo.tocsr()
    Sg = sp.csr_matrix((np.ones(Mg.nnz), Mg.indices, Mg.indptr), shape=Mg.shape)
    So = sp.csr_matrix((np.ones(Mo.nnz), Mo.indices, Mo.indptr), shape=Mo.shape)
    only_g = (Sg - Sg.multiply(So)).tocsr()
    only_g.eliminate_zeros()
    only_o = (So - So.multiply(Sg)).tocsr()
    only_o.eliminate_zeros()
    scale = max(np.abs(Mo.data).max() if Mo.nnz else 0.0, 1e-300)
    worst = 0.0
    if only_g.nnz:
        worst = max(worst, np.abs(Mg.multiply(only_g).data).max(initial=0.0))
    if only_o.nnz:
        worst = max(worst, np.abs(Mo.multiply(only_o).data).max(initial=0.0))
    return float(worst / scale), int(only_g.nnz), int(only_o.nnz)


def _mis_allowance(Ho, level, mis, Zo, oo, mo, AEI, AEJ, input_diff=0.0):
    """How well the DATA determine the column space kept for one MIS.  The kept space consists of
    the left singular vectors with sigma_i > 1e-10 sigma_0 (amg/src/xpacks.cpp:609-610); by Wedin's
    theorem a perturbation dM of the gathered matrix moves it by ~ ||dM|| / gap, where gap is the
    distance from the smallest kept singular value to the next one (or to zero).  The inputs of
    the two SVDs are the eigenvectors of the two eigensolvers, which are themselves only determined
    to ~ eps ||A^|| / gap_lambda (the eigenvalue nearest theta is ~1e-3 from it: 1e-13 .. 1e-11
    between two backward-stable solvers, tests/test_cholsi.py), so the subspace tolerance is
        max(1e-8, 4 (sigma_0 / gap) amp max(2048 eps, input_diff))
    with input_diff = the eigenspace difference actually MEASURED on this level between the CUDA
    path and the oracle (`eigenspace_sin`, itself held to 1e-8), amp = max ||z|| / ||z restricted to
    the MIS|| over the gathered columns (they are normalised AFTER the restriction), 4 = norm
    conversions (several unit columns per matrix).  A MIS above 1e-8 is counted (`mis_ill_conditioned`).
    Recomputed here from the oracle's eigenvectors (the same gather / boundary filter / column
    normalisation as contrib.cpp:492-687)."""
    m2a_I, m2a_J = Ho.get("mis_to_AE.I", level), Ho.get("mis_to_AE.J", level)
    m2d_I, m2d_J = Ho.get("mis_to_dof.I", level), Ho.get("mis_to_dof.J", level)
    flags = Ho.get("agg_flags", level)
    dofs = m2d_J[m2d_I[mis] : m2d_I[mis + 1]]
    cols, full = [], []
    for ae in m2a_J[m2a_I[mis] : m2a_I[mis + 1]]:
        n = AEI[ae + 1] - AEI[ae]
        k = mo[ae]
        Z = Zo[oo[ae] : oo[ae + 1]].reshape(k, n).T
        loc = {g: i for i, g in enumerate(AEJ[AEI[ae] : AEI[ae + 1]])}
        cols.append(Z[[loc[g] for g in dofs], :])
        full.append(np.linalg.norm(Z, axis=0))
    M = np.concatenate(cols, axis=1)
    M[(flags[dofs] & 2) != 0, :] = 0.0
    nrm = np.linalg.norm(M, axis=0)
    keep = nrm > 1e-10
    # a column is the restriction of an eigenvector to the MIS, NORMALISED: a difference of the
    # eigenvector (relative to its own norm) is magnified by ||z|| / ||z restricted to the MIS||
    amp = float(np.max(np.concatenate(full)[keep] / nrm[keep])) if np.any(keep) else 1.0
    M = M[:, keep] / nrm[keep]
    if M.shape[1] == 0:
        return 1e-8
    sv = np.linalg.svd(M, compute_uv=False)
    kept = sv[sv > 1e-10 * sv[0]]
    nxt = sv[len(kept)] if len(kept) < len(sv) else 0.0
    gap = max(kept[-1] - nxt, 1e-300)
    return max(1e-8, 4.0 * (sv[0] / gap) * amp * max(2048 * np.finfo(float).eps, input_diff))


def compare_level(Hg, Ho, level, S_prev=None, check_celmat=True):
    """Returns (metrics dict, S for the next level or None if the gauge is not diagonal)."""
    m = {}
    # ---- integer maps: bit-exact
    bad = [k for k in MAPS if not np.array_equal(Hg.get(k, level), Ho.get(k, level))]
    m["maps_mismatch"] = bad
    AEI = Ho.get("AE_to_dof.I", level)
    AEJ = Ho.get("AE_to_dof.J", level)
    nparts = len(AEI) - 1
    ND = int(Ho.scalar("ND", level))
    if S_prev is None:
        S_prev = np.ones(ND)

    # ---- per-AE spectral data
    mg, mo = Hg.get("ae_m", level), Ho.get("ae_m", level)
    m["ae_m_mismatch"] = int(np.sum(mg != mo))
    Dg, Do = Hg.get("ae_D", level), Ho.get("ae_D", level)
    m["D_relerr"] = float(np.max(np.abs(Dg - Do) / np.abs(Do))) if len(Do) else 0.0
    eg, eo = Hg.get("evals", level), Ho.get("evals", level)
    if len(eg) == len(eo):
        m["eval_err"] = float(np.max(np.abs(eg - eo) / np.maximum(np.abs(eo), 1.0))) if len(eo) else 0.0
    else:
        m["eval_err"] = float("inf")
    og, oo = Hg.get("ae_evect_off", level), Ho.get("ae_evect_off", level)
    Zg, Zo = Hg.get("evects", level), Ho.get("evects", level)
    worst = 0.0
    dnorm_worst = 0.0
    if m["ae_m_mismatch"] == 0:
        for i in range(nparts):
            n = AEI[i + 1] - AEI[i]
            k = mo[i]
            if k == 0:
                continue
            dofs = AEJ[AEI[i] : AEI[i + 1]]
            s = S_prev[dofs]
            D = Do[AEI[i] : AEI[i + 1]]
            A = Zg[og[i] : og[i + 1]].reshape(k, n).T * s[:, None]
            B = Zo[oo[i] : oo[i + 1]].reshape(k, n).T
            # D-orthonormalise both (they are D-orthonormal up to roundoff / injected vector)
            sq = np.sqrt(D)[:, None]
            Qa, _ = np.linalg.qr(A * sq)
            Qb, _ = np.linalg.qr(B * sq)
            sin = np.linalg.norm(Qa - Qb @ (Qb.T @ Qa), 2)
            worst = max(worst, sin)
            # normalisation z^T D z = 1 of the GPU vectors (skip an injected all-ones column)
            nd = np.abs(np.sum(A * A * D[:, None], axis=0) - 1.0)
            if i == 0 and len(nd) and np.allclose(np.abs(A[:, -1]), 1.0):
                nd = nd[:-1]
            if len(nd):
                dnorm_worst = max(dnorm_worst, float(nd.max()))
    else:
        worst = float("inf")
    m["eigenspace_sin"] = float(worst)
    m["evect_Dnorm_err"] = float(dnorm_worst)

    # ---- per-MIS blocks
    kg, ko = Hg.get("mis_numcoarsedof", level), Ho.get("mis_numcoarsedof", level)
    m["mis_ncd_mismatch"] = int(np.sum(kg != ko))
    m["NDc"] = int(ko.sum())
    S_next = None
    if m["mis_ncd_mismatch"] == 0:
        misI = Ho.get("mis_to_dof.I", level)
        misJ = Ho.get("mis_to_dof.J", level)
        offg, offo = Hg.get("mis_off", level), Ho.get("mis_off", level)
        Tg, To = Hg.get("mis_tent", level), Ho.get("mis_tent", level)
        nmis = len(ko)
        worst = 0.0
        orth = 0.0
        rows, cols, vals = [], [], []
        col0 = 0
        diag_gauge = True
        excess = 0.0       # worst sin / conditioning-aware allowance (see _mis_allowance)
        allowance = 0.0    # largest allowance granted
        ill_conditioned = 0  # MISes whose kept space is determined worse than 1e-8 by the data
        for mis in range(nmis):
            k = ko[mis]
            s = misI[mis + 1] - misI[mis]
            if k == 0:
                continue
            sg = S_prev[misJ[misI[mis] : misI[mis + 1]]]
            Ug = Tg[offg[mis] : offg[mis + 1]].reshape(k, s).T * sg[:, None]
            Uo = To[offo[mis] : offo[mis + 1]].reshape(k, s).T
            Q = Uo.T @ Ug
            sin = np.linalg.norm(Ug - Uo @ Q, 2)
            if sin > 1e-8:
                allow = _mis_allowance(Ho, level, mis, Zo, oo, mo, AEI, AEJ, m["eigenspace_sin"])
                ill_conditioned += 1
                excess = max(excess, sin / allow)
                allowance = max(allowance, allow)
            else:
                worst = max(worst, sin)
            orth = max(orth, np.linalg.norm(Ug.T @ Ug - np.eye(k), 2))
            if np.max(np.abs(np.abs(Q) - np.eye(k))) > 1e-6:
                diag_gauge = False
            for a in range(k):
                for b in range(k):
                    rows.append(col0 + a)
                    cols.append(col0 + b)
                    vals.append(Q[a, b])
            col0 += k
        m["mis_ill_conditioned"] = int(ill_conditioned)
        m["mis_space_excess"] = float(excess)
        m["space_allowance"] = float(allowance)
        m["mis_space_sin"] = float(worst)
        m["mis_orth_err"] = float(orth)
        m["gauge_is_signs"] = bool(diag_gauge)
        NDc = col0
        Q = sp.csr_matrix((vals, (rows, cols)), shape=(NDc, NDc))
        Sd = sp.diags(S_prev)
        # ---- tentative and final prolongator
        for name in ("tent_interp", "interp"):
            Pg, Po = Hg.csr(name, level), Ho.csr(name, level)
            pm = _pattern_mismatch(Pg, Po)
            m[name + "_pattern_mismatch"] = pm[0]
            m[name + "_pattern_only"] = [pm[1], pm[2]]
            diff = (Pg - Sd @ Po @ Q).tocsr()
            scale = max(1e-300, np.abs(Po.data).max())
            m[name + "_err"] = float(np.abs(diff.data).max() / scale) if diff.nnz else 0.0
        # ---- coarse operator
        Ag, Ao = Hg.csr("Ac", level), Ho.csr("Ac", level)
        pm = _pattern_mismatch(Ag, Ao)
        m["Ac_pattern_mismatch"] = pm[0]
        m["Ac_pattern_only"] = [pm[1], pm[2]]
        diff = (Ag - Q.T @ Ao @ Q).tocsr()
        m["Ac_err"] = float(np.abs(diff.data).max() / np.abs(Ao.data).max()) if diff.nnz else 0.0
        m["Ac_nnz"] = int(Ao.nnz)
        # relative error entry by entry on entries that are not tiny
        A2 = (Q.T @ Ao @ Q).tocsr()
        big = np.abs(A2.data) > 1e-8 * np.abs(Ao.data).max()
        if big.any():
            Agd = Ag.tocsr()
            r, c = A2.nonzero()
            ref = np.asarray(A2[r, c]).ravel()
            got = np.asarray(Agd[r, c]).ravel()
            sel = np.abs(ref) > 1e-8 * np.abs(Ao.data).max()
            m["Ac_entry_relerr"] = float(np.max(np.abs(got[sel] - ref[sel]) / np.abs(ref[sel])))
        # ---- Dinv_neg
        dg, do = Hg.get("Dinv_neg", level), Ho.get("Dinv_neg", level)
        m["Dinv_relerr"] = float(np.max(np.abs(dg - do) / np.abs(do)))
        if diag_gauge:
            S_next = np.asarray(Q.diagonal()).copy()
            S_next = np.where(S_next >= 0, 1.0, -1.0)
        # ---- coarse element matrices (exist when a coarser level was built)
        cg = Hg.get("celmat", level)
        co = Ho.get("celmat", level)
        if check_celmat and len(co) and len(cg) == len(co):
            offs = Ho.get("celmat_off", level)
            ceI = Ho.get("elem_to_dof.I", level + 1)
            ceJ = Ho.get("elem_to_dof.J", level + 1)
            Qd = Q.toarray() if NDc <= 4000 else None
            worst = 0.0
            scale = np.abs(co).max()
            for e in range(nparts):
                cd = ceJ[ceI[e] : ceI[e + 1]]
                nc = len(cd)
                if nc == 0:
                    continue
                Mg = cg[offs[e] : offs[e + 1]].reshape(nc, nc).T
                Mo = co[offs[e] : offs[e + 1]].reshape(nc, nc).T
                Qe = Qd[np.ix_(cd, cd)] if Qd is not None else Q[cd][:, cd].toarray()
                worst = max(worst, np.abs(Mg - Qe.T @ Mo @ Qe).max() / scale)
            m["celmat_err"] = float(worst)
        elif len(co) != len(cg):
            m["celmat_err"] = float("inf")
    return m, S_next


def compare_hierarchies(Hg, Ho, expect_levels=None):
    """Compares level by level.  Returns one metrics dict per level that could be compared entry
    by entry.  A level can only be compared when the coarse bases of the two hierarchies differ by
    signs (S diagonal): a ROTATION inside a MIS block (degenerate singular values, e.g. constant
    coefficients on symmetric agglomerates -- LAPACK's basis is arbitrary there as well) changes
    the weighted-l1 matrix D of the next level, which is not invariant under rotations, so the two
    next-level eigenproblems are genuinely different problems.  In that case the remaining levels
    are compared through what IS invariant (sizes, integer maps, spectrum of the operator) and
    the returned list stops; callers state how many levels they expect (`expect_levels`)."""
    out = []
    S = None
    nl = int(Ho.scalar("num_coarsenings", 0))
    for l in range(nl):
        m, S = compare_level(Hg, Ho, l, S)
        m["level"] = l
        out.append(m)
        if S is None and l + 1 < nl:
            m["note"] = "gauge not diagonal: coarser levels compared through invariants only"
            m["coarser_invariants"] = [compare_level_invariants(Hg, Ho, k) for k in range(l + 1, nl)]
            break
    if expect_levels is not None:
        assert len(out) == expect_levels, ("levels compared entry-wise", len(out), "expected", expect_levels,
                                            [m.get("note") for m in out])
    return out


def compare_level_invariants(Hg, Ho, level):
    """What two hierarchies share on a level whose basis differs by a block rotation: the integer
    maps (bit-exact), the sizes, and the spectrum of the level's operator (A_g = Q^T A_o Q)."""
    m = {"level": level}
    m["maps_mismatch"] = [k for k in MAPS if not np.array_equal(Hg.get(k, level), Ho.get(k, level))]
    Ag, Ao = Hg.csr("Ac", level - 1), Ho.csr("Ac", level - 1)
    m["ND"] = [Ag.shape[0], Ao.shape[0]]
    if Ag.shape == Ao.shape and Ag.shape[0] <= 3000:
        wg = np.linalg.eigvalsh(Ag.toarray())
        wo = np.linalg.eigvalsh(Ao.toarray())
        m["operator_spectrum_err"] = float(np.max(np.abs(wg - wo)) / max(np.abs(wo).max(), 1e-300))
    return m


def pattern_counts(res):
    """Entries present in only one of the two patterns, per level: the deviation from
    'bit-exact sparsity patterns' (numerical zeros, see _pattern_mismatch)."""
    return [{k: m[k] for k in m if k.endswith("_pattern_only")} for m in res]


def assert_level_ok(m, level, eig_tol=1e-10, space_tol=1e-8, ac_tol=1e-9):
    assert m["maps_mismatch"] == [], (level, m["maps_mismatch"])
    assert m["ae_m_mismatch"] == 0, (level, "ae_m", m["ae_m_mismatch"])
    assert m["D_relerr"] <= (1e-11 if level == 0 else 1e-7), (level, "D", m["D_relerr"])
    assert m["eval_err"] <= eig_tol, (level, "eval", m["eval_err"])
    assert m["eigenspace_sin"] <= space_tol, (level, "eigenspace", m["eigenspace_sin"])
    assert m["evect_Dnorm_err"] <= 1e-10, (level, "Dnorm", m["evect_Dnorm_err"])
    assert m["mis_ncd_mismatch"] == 0, (level, "mis_ncd", m["mis_ncd_mismatch"])
    # MIS column spaces: 1e-8, except for the MISes whose kept singular vectors the data determine
    # worse than that (count reported; each must stay within its own Wedin bound)
    assert m["mis_space_sin"] <= space_tol, (level, "mis_space", m["mis_space_sin"])
    assert m.get("mis_space_excess", 0.0) <= 1.0, (level, "mis_space beyond its conditioning bound",
                                                   m["mis_space_excess"], m["mis_ill_conditioned"])
    space_tol = max(space_tol, m.get("space_allowance", 0.0))
    assert m["mis_orth_err"] <= 1e-10, (level, "mis_orth", m["mis_orth_err"])
    # patterns: identical up to entries that are numerical zeros (see _pattern_mismatch)
    assert m["tent_interp_pattern_mismatch"] <= space_tol, (level, "tent pattern", m["tent_interp_pattern_mismatch"])
    assert m["interp_pattern_mismatch"] <= space_tol, (level, "interp pattern", m["interp_pattern_mismatch"])
    assert m["Ac_pattern_mismatch"] <= ac_tol, (level, "Ac pattern", m["Ac_pattern_mismatch"])
    assert m["tent_interp_err"] <= space_tol, (level, "tent", m["tent_interp_err"])
    assert m["interp_err"] <= space_tol, (level, "interp", m["interp_err"])
    assert m["Ac_err"] <= ac_tol, (level, "Ac", m["Ac_err"])
    assert m["Dinv_relerr"] <= (ac_tol if level == 0 else 1e-7), (level, "Dinv", m["Dinv_relerr"])
    if "celmat_err" in m:
        assert m["celmat_err"] <= ac_tol, (level, "celmat", m["celmat_err"])
