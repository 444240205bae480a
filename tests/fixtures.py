"""Reference fixtures restated for the tests."""
import ctypes

import numpy as np

import saamge_b200 as sab


def mltest_problem(order=1, levels=2):
    """The 12-quad mltest.mesh fixture of the reference's CTest drivers: 4 x 3 cells on
    the unit square (amg/test/mltest.mesh), essential boundary = attribute 4 = side x=0
    (amg/test/mltest/mltest.cpp:476-479), checkerboard coefficient 1e6/1 evaluated at
    element centres (:156-175, projected on piecewise constants :609-611), rhs 1,
    hard-coded AE partition (:221-229), extra all-ones vector on AE 0
    (amg/src/interp.cpp:510-524), coarse partition {0,0,1,1} (amg/src/aggregates.cpp:1781-1783).
    Pinned results: 'Outer PCG converged in 3 iterations' for Q1 2-level and 3-level,
    4 iterations for order 2 (amg/CMakeLists.txt:191-217)."""
    h = sab.host_lib()
    h.sa_drv_problem_create_ex.restype = ctypes.c_void_p
    h.sa_drv_problem_create_ex.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_uint64, ctypes.c_int]
    h.sa_drv_problem_partition_array.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    h.sa_drv_problem_set_coarse_partition.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    pr = sab.Problem.__new__(sab.Problem)
    pr.handle = h.sa_drv_problem_create_ex(2, 4, 3, 1, order, 3, 1e6, 0, 1)
    pr.dim, pr.n, pr.order = 2, 4, order
    part = np.array([0, 0, 1, 1, 0, 0, 2, 2, 3, 3, 3, 2], dtype=np.int32)
    h.sa_drv_problem_partition_array(pr.handle, part.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 4)
    cpart = np.array([0, 0, 1, 1], dtype=np.int32)
    h.sa_drv_problem_set_coarse_partition(pr.handle, 1, cpart.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 4)
    p = sab.default_params(num_levels=levels, first_elems_per_agg=3, elems_per_agg=2, testmesh_inject=1)
    return pr, p
