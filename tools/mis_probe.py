"""Development probe: conditioning of the MIS whose kept space differs most between the CUDA
path and the oracle (config cfg1 of tests/test_gpu.py)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import saamge_b200 as sab, oracle_util as ou, parity

p = sab.default_params(num_levels=3, first_elems_per_agg=64, elems_per_agg=8, first_nu_pro=1, nu_pro=1,
                       partition_kind=0, block=(4, 4, 4), coarse_block=2)
pr = sab.Problem(3, 16, order=1, coef_kind=1); pr.partition(p)
Ho = ou.orc_build(pr, p); Hg = sab.ml_build(pr, p); sab.ml_download(Hg)
orig = parity._mis_allowance
def wrapped(*a):
    r = orig(*a)
    print("MIS", a[2], "allowance", r, flush=True)
    return r
parity._mis_allowance = wrapped
res = parity.compare_hierarchies(Hg, Ho, expect_levels=2)
for m in res:
    print({k: v for k, v in m.items() if k in ("level", "eigenspace_sin", "eval_err", "mis_space_sin", "mis_space_excess", "mis_ill_conditioned", "space_allowance", "ae_m_mismatch")})
# eigenvalues near theta on level 1
ev = Ho.get("evals", 1); print("level-1 oracle eigenvalues closest to theta:", np.sort(np.abs(ev - 0.003))[:5])
