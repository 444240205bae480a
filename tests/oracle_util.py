"""Loader for the CPU oracle (oracle/liboracle.so) -- test infrastructure only."""
import ctypes
import glob
import os

import saamge_b200 as sab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
_orc = None


def lapack_path():
    import scipy

    pats = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so")
    libs = glob.glob(pats)
    if not libs:
        raise RuntimeError("scipy's bundled OpenBLAS not found")
    return os.path.abspath(libs[0])


def oracle(num_threads=0):
    global _orc
    if _orc is None:
        sab.host_lib()
        o = ctypes.CDLL(ORACLE_LIB)
        o.sa_orc_init.argtypes = [ctypes.c_char_p, ctypes.c_int]
        o.sa_orc_ml_build.restype = ctypes.c_void_p
        o.sa_orc_ml_build.argtypes = [ctypes.c_void_p, ctypes.POINTER(sab.Params)]
        o.sa_orc_ml_pcg.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        o.sa_orc_time_local_spectral.restype = ctypes.c_double
        o.sa_orc_time_local_spectral.argtypes = [ctypes.c_void_p, ctypes.POINTER(sab.Params), ctypes.c_int, ctypes.c_int]
        o.sa_orc_check_mises.argtypes = [ctypes.c_void_p]
        o.sa_orc_init(lapack_path().encode(), num_threads)
        _orc = o
    elif num_threads > 0:
        _orc.sa_orc_init(lapack_path().encode(), num_threads)  # re-set the OpenMP thread count
    return _orc


def orc_build(problem, params):
    return sab.Hierarchy(oracle().sa_orc_ml_build(problem.handle, ctypes.byref(params)))


def orc_pcg(hier, maxiter=1000, rtol=1e-12, atol=0.0):
    return oracle().sa_orc_ml_pcg(hier.handle, maxiter, rtol, atol)


def orc_update_operators(hier, resmooth_interp=True):
    """adapt_update_operators on the oracle's hierarchy (the problem's current operator values)."""
    o = oracle()
    o.sa_orc_ml_update_operators.argtypes = [ctypes.c_void_p, ctypes.c_int]
    return o.sa_orc_ml_update_operators(hier.handle, 1 if resmooth_interp else 0)


def orc_build_algebraic(problem, params):
    o = oracle()
    o.sa_orc_ml_build_algebraic.restype = ctypes.c_void_p
    o.sa_orc_ml_build_algebraic.argtypes = [ctypes.c_void_p, ctypes.POINTER(sab.Params)]
    return sab.Hierarchy(o.sa_orc_ml_build_algebraic(problem.handle, ctypes.byref(params)))
