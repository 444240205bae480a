import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

NAMES = list(make_golden.CONFIGS.keys())


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def check_against_golden(H, p, gold, eval_tol=1e-10, spec_tol=1e-9):
    nlev = p.num_levels - 1
    inv = make_golden.invariants(H, nlev)
    for l in range(nlev):
        for key in ("ae_m", "mis_ncd", "partitioning", "mises"):
            k = "l%d_%s" % (l, key)
            assert np.array_equal(inv[k], gold[k]), k
        k = "l%d_evals" % l
        assert np.max(np.abs(inv[k] - gold[k]) / np.maximum(1.0, np.abs(gold[k]))) <= eval_tol, k
        k = "l%d_Ac_spectrum" % l
        scale = np.abs(gold[k]).max()
        assert np.max(np.abs(inv[k] - gold[k])) <= spec_tol * scale, k
        k = "l%d_Dinv_neg" % l
        assert np.max(np.abs(inv[k] - gold[k]) / np.abs(gold[k])) <= 1e-8, k
    assert np.max(np.abs(inv["l0_ae_D"] - gold["l0_ae_D"]) / gold["l0_ae_D"]) <= 1e-11
    assert abs(int(inv["pcg_iterations"][0]) - int(gold["pcg_iterations"][0])) <= 1
    n = min(len(inv["pcg_brr"]), len(gold["pcg_brr"]), 2)
    assert np.allclose(inv["pcg_brr"][:n], gold["pcg_brr"][:n], rtol=1e-6)
