"""CLI: build a hierarchy with the B200 path and with the oracle, print parity metrics.
usage: python tests/run_parity.py dim n order coef levels first_epa epa nu_pro [kind] [--self]"""
import json
import os
import sys
import time

import conftest  # noqa: F401  (sys.path)
import oracle_util as ou
import parity
import saamge_b200 as sab


def main():
    a = sys.argv[1:]
    self_mode = "--self" in a
    a = [x for x in a if not x.startswith("--")]
    dim, n, order, coef, levels, fepa, epa, nupro = [int(x) for x in a[:8]]
    kind = int(a[8]) if len(a) > 8 else 0
    blk = int(a[9]) if len(a) > 9 else 4
    cblk = int(a[10]) if len(a) > 10 else 2
    p = sab.default_params(num_levels=levels, first_elems_per_agg=fepa, elems_per_agg=epa,
                           first_nu_pro=nupro, nu_pro=nupro, partition_kind=kind,
                           block=(blk, blk, blk), coarse_block=cblk)
    if os.environ.get("SA_TEST_THETA"):  # spectral threshold of every level (default 0.003)
        p.first_theta = p.theta = float(os.environ["SA_TEST_THETA"])
    pr = sab.Problem(dim, n, order=order, coef_kind=coef)
    na = pr.partition(p)
    print("AEs", na, "ND", pr.scalar("ND"), "mises", pr.scalar("num_mises"), flush=True)
    t = time.time()
    Ho = ou.orc_build(pr, p)
    print("oracle build %.3fs" % (time.time() - t), flush=True)
    ito = ou.orc_pcg(Ho)
    t = time.time()
    if self_mode:
        Hg = ou.orc_build(pr, p)
        itg = ou.orc_pcg(Hg)
    else:
        Hg = sab.ml_build(pr, p)
        print("gpu build %.3fs" % (time.time() - t), flush=True)
        itg = sab.ml_pcg(Hg)
        sab.ml_download(Hg)
        print("gpu times", {k: round(v, 4) for k, v in Hg.times().items()})
        print("orc times", {k: round(v, 4) for k, v in Ho.times().items()})
    print("pcg iters gpu", itg, "oracle", ito)
    print("brr gpu", Hg.get("pcg.brr")[:8], "\nbrr orc", Ho.get("pcg.brr")[:8])
    print("final res gpu", Hg.scalar("pcg.final_res_norm"), "orc", Ho.scalar("pcg.final_res_norm"))
    res = parity.compare_hierarchies(Hg, Ho, expect_levels=levels - 1)
    for l, m in enumerate(res):
        print("level", l, json.dumps(m, indent=1))
    ok = True
    for l, m in enumerate(res):
        try:
            parity.assert_level_ok(m, l, ac_tol=float(os.environ.get("SA_TEST_AC_TOL", "1e-9")))
        except AssertionError as e:
            ok = False
            print("FAIL", e)
    print("PARITY", "OK" if ok and abs(itg - ito) <= 1 else "FAILED")


if __name__ == "__main__":
    main()
