"""Generates the golden fixtures of tests/golden/ from the CPU oracle (run on the
CPU box: python tests/golden/make_golden.py).  Only gauge-invariant quantities are
stored (they do not depend on eigenvector signs / SVD bases): counts, eigenvalues,
weighted-l1 diagonals, the spectrum of every coarse operator, PCG histories."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fixtures  # noqa: E402
import oracle_util as ou  # noqa: E402
import saamge_b200 as sab  # noqa: E402

CONFIGS = {
    "hex8_lognormal_3level": dict(dim=3, n=8, coef=1, params=dict(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, first_nu_pro=1, nu_pro=1, partition_kind=1, block=(4, 4, 4), coarse_block=2)),
    "quad32_poisson_metis_2level": dict(dim=2, n=32, coef=0, params=dict(num_levels=2, first_elems_per_agg=64, elems_per_agg=64)),
    "hex12_lognormal_metis_3level": dict(dim=3, n=12, coef=1, params=dict(num_levels=3, first_elems_per_agg=48, elems_per_agg=6)),
    "mltest_3level": None,
}


def build(name):
    cfg = CONFIGS[name]
    if cfg is None:
        pr, p = fixtures.mltest_problem(1, 3)
    else:
        p = sab.default_params(**cfg["params"])
        pr = sab.Problem(cfg["dim"], cfg["n"], coef_kind=cfg["coef"])
        pr.partition(p)
    return pr, p


def invariants(H, nlev):
    out = {}
    for l in range(nlev):
        out["l%d_ae_m" % l] = H.get("ae_m", l)
        out["l%d_evals" % l] = H.get("evals", l)
        out["l%d_mis_ncd" % l] = H.get("mis_numcoarsedof", l)
        out["l%d_partitioning" % l] = H.get("partitioning", l)
        out["l%d_mises" % l] = H.get("mises", l)
        if l == 0:
            out["l0_ae_D"] = H.get("ae_D", 0)
        Ac = H.csr("Ac", l).toarray()
        out["l%d_Ac_spectrum" % l] = np.linalg.eigvalsh(0.5 * (Ac + Ac.T))
        out["l%d_Dinv_neg" % l] = H.get("Dinv_neg", l)
    out["pcg_iterations"] = np.array([int(H.scalar("pcg.iterations"))])
    out["pcg_brr"] = H.get("pcg.brr")
    return out


if __name__ == "__main__":
    for name in CONFIGS:
        pr, p = build(name)
        H = ou.orc_build(pr, p)
        ou.orc_pcg(H)
        inv = invariants(H, p.num_levels - 1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **inv)
        print(name, "iters", inv["pcg_iterations"], "NDc", [len(inv["l%d_Ac_spectrum" % l]) for l in range(p.num_levels - 1)])
