"""ctypes binding of the CUDA library's C ABI (include/saamge_b200.h), one call per
entry point.  Used by the GPU tests to exercise the drop-in boundary directly, without
the C++ host mirror in between."""
import ctypes

import numpy as np

from . import gpu_lib

_i = ctypes.POINTER(ctypes.c_int)
_d = ctypes.POINTER(ctypes.c_double)
_l = ctypes.POINTER(ctypes.c_int64)
_c = ctypes.c_char_p


class LevelDesc(ctypes.Structure):
    _fields_ = [
        ("ND", ctypes.c_int), ("NE", ctypes.c_int), ("nparts", ctypes.c_int), ("num_mises", ctypes.c_int),
        ("elem_to_dof_I", _i), ("elem_to_dof_J", _i),
        ("dof_to_elem_I", _i), ("dof_to_elem_J", _i),
        ("AE_to_elem_I", _i), ("AE_to_elem_J", _i),
        ("AE_to_dof_I", _i), ("AE_to_dof_J", _i),
        ("dof_to_AE_I", _i), ("dof_to_AE_J", _i), ("dof_id_inAE", _i),
        ("partitioning", _i), ("agg_flags", ctypes.c_void_p),
        ("mis_to_dof_I", _i), ("mis_to_dof_J", _i),
        ("mis_to_AE_I", _i), ("mis_to_AE_J", _i),
        ("AE_to_mis_I", _i), ("AE_to_mis_J", _i),
        ("mises", _i),
        ("A_I", _i), ("A_J", _i), ("A_data", _d),
        ("elmat", _d), ("elmat_off", _l),
        ("assemble_with_global", ctypes.c_int),
        ("mis_coarsedofoffsets", _i),
        ("async_upload", ctypes.c_int),
    ]


def _ip(a):
    return a.ctypes.data_as(_i)


def _dp(a):
    return a.ctypes.data_as(_d)


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, gpu_lib().sa_gpu_last_error().decode()))


class Context:
    def __init__(self, device=0):
        self.lib = gpu_lib()
        self.h = ctypes.c_void_p()
        check(self.lib.sa_gpu_ctx_create(device, ctypes.byref(self.h)), "sa_gpu_ctx_create")

    def close(self):
        if self.h:
            self.lib.sa_gpu_ctx_destroy(self.h)
            self.h = None


class Level:
    """Finest level built from a saamge_b200.Problem (relations must exist)."""

    def __init__(self, ctx, problem, with_global=True, give_operator=True, async_upload=False):
        self.ctx, self.lib = ctx, ctx.lib
        g = problem.get
        self.keep = {}
        d = LevelDesc()
        names = ["elem_to_dof", "dof_to_elem", "AE_to_elem", "AE_to_dof", "dof_to_AE", "mis_to_dof", "mis_to_AE", "AE_to_mis"]
        for nm in names:
            for f in ("I", "J"):
                a = np.ascontiguousarray(g("%s.%s" % (nm, f)), dtype=np.int32)
                if len(a) == 0:
                    a = np.zeros(1, dtype=np.int32)
                self.keep[nm + f] = a
                setattr(d, "%s_%s" % (nm, f), _ip(a))
        for nm in ("dof_id_inAE", "partitioning", "mises"):
            a = np.ascontiguousarray(g(nm), dtype=np.int32)
            self.keep[nm] = a
            setattr(d, nm, _ip(a))
        fl = np.ascontiguousarray(g("agg_flags"), dtype=np.int8)
        self.keep["flags"] = fl
        d.agg_flags = fl.ctypes.data
        d.ND = int(problem.scalar("ND"))
        d.NE = len(self.keep["elem_to_dofI"]) - 1
        d.nparts = int(problem.scalar("nparts"))
        d.num_mises = int(problem.scalar("num_mises"))
        self.AI, self.AJ, self.AA = g("A.I"), g("A.J"), g("A.A")
        if give_operator:
            d.A_I, d.A_J, d.A_data = _ip(self.AI), _ip(self.AJ), _dp(self.AA)
        ne = int(problem.scalar("ne"))
        self.elmat = g("elmat")
        self.elmat_off = (np.arange(d.NE + 1, dtype=np.int64) * ne * ne)
        d.elmat = _dp(self.elmat)
        d.elmat_off = self.elmat_off.ctypes.data_as(_l)
        d.assemble_with_global = 1 if with_global else 0
        # pipelined upload: the numpy arrays in self.keep / self.elmat stay alive with the object
        d.async_upload = 1 if async_upload else 0
        self.desc = d
        self.nparts, self.ND, self.num_mises = d.nparts, d.ND, d.num_mises
        self.h = ctypes.c_void_p()
        check(self.lib.sa_gpu_level_create(ctx.h, ctypes.byref(d), None, ctypes.byref(self.h)), "sa_gpu_level_create")

    def ae_sizes(self):
        I = self.keep["AE_to_dofI"]
        return I[1:] - I[:-1]

    def local_spectral(self, theta, a0=0, a1=None, inject=0):
        a1 = self.nparts if a1 is None else a1
        check(self.lib.sa_gpu_local_spectral(self.h, ctypes.c_double(theta), a0, a1, inject), "sa_gpu_local_spectral")

    def spectral(self):
        m = np.zeros(self.nparts, dtype=np.int32)
        check(self.lib.sa_gpu_get_spectral_counts(self.h, _ip(m)), "counts")
        n = self.ae_sizes()
        ev = np.zeros(int(m.sum()))
        Z = np.zeros(int((m.astype(np.int64) * n).sum()))
        D = np.zeros(int(n.sum()))
        check(self.lib.sa_gpu_get_spectral(self.h, _dp(ev), _dp(Z), _dp(D)), "get_spectral")
        return m, ev, Z, D

    def set_spectral(self, m, evals, evects, D):
        m = np.ascontiguousarray(m, dtype=np.int32)
        evals = np.ascontiguousarray(evals, dtype=np.float64)
        evects = np.ascontiguousarray(evects, dtype=np.float64)
        D = np.ascontiguousarray(D, dtype=np.float64)
        check(self.lib.sa_gpu_set_spectral(self.h, 0, self.nparts, _ip(m), _dp(evals), _dp(evects), _dp(D)),
              "set_spectral")

    def local_spectral_sharded(self, theta, dist, group=None):
        """Local spectral stage sharded over the ranks of `dist` (contiguous AE ranges
        balanced on n^3), results all-gathered and installed on every rank."""
        from . import sharding

        n = self.ae_sizes()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        a, b = sharding.shard_ranges(n, world)[rank]
        self.local_spectral(theta, a, b)
        m, ev, Z, D = self.spectral()
        do = np.concatenate([[0], np.cumsum(n)])
        full = sharding.allgather_spectral((m[a:b], ev, Z, D[do[a]:do[b]]), dist, group)
        self.set_spectral(*full)
        return a, b

    def build_AE_stiff(self, part):
        n = int(self.ae_sizes()[part])
        out = np.zeros(n * n)
        check(self.lib.sa_gpu_build_AE_stiff(self.h, part, _dp(out)), "build_AE_stiff")
        return out.reshape(n, n).T

    def tentative_P(self, avoid_ess=1):
        ncd = np.zeros(self.num_mises, dtype=np.int32)
        NDc = ctypes.c_int()
        check(self.lib.sa_gpu_tentative_P(self.h, avoid_ess, _ip(ncd), ctypes.byref(NDc)), "tentative_P")
        return ncd, NDc.value

    def build_Dinv_neg(self):
        check(self.lib.sa_gpu_build_Dinv_neg(self.h), "Dinv")
        out = np.zeros(self.ND)
        check(self.lib.sa_gpu_get_Dinv_neg(self.h, _dp(out)), "get Dinv")
        return out

    def smooth_P(self, roots):
        r = np.ascontiguousarray(roots, dtype=np.float64)
        check(self.lib.sa_gpu_smooth_P(self.h, len(r), _dp(r) if len(r) else None), "smooth_P")

    def rap(self):
        check(self.lib.sa_gpu_rap(self.h), "rap")

    def csr(self, which):
        import scipy.sparse as sp

        r, c, nnz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(self.lib.sa_gpu_get_csr_sizes(self.h, which, ctypes.byref(r), ctypes.byref(c), ctypes.byref(nnz)), "sizes")
        I = np.zeros(r.value + 1, dtype=np.int32)
        J = np.zeros(max(nnz.value, 1), dtype=np.int32)
        A = np.zeros(max(nnz.value, 1))
        check(self.lib.sa_gpu_get_csr(self.h, which, _ip(I), _ip(J), _dp(A)), "get_csr")
        return sp.csr_matrix((A[: nnz.value], J[: nnz.value], I), shape=(r.value, c.value))

    def spmv(self, which, x):
        M = self.csr(which)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(M.shape[0])
        check(self.lib.sa_gpu_spmv(self.h, which, _dp(x), _dp(y)), "spmv")
        return y

    def poly_smooth(self, b, x, roots):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        r = np.ascontiguousarray(roots, dtype=np.float64)
        check(self.lib.sa_gpu_poly_smooth(self.h, _dp(b), _dp(x), len(r), _dp(r)), "poly_smooth")
        return x

    def close(self):
        if self.h:
            self.lib.sa_gpu_level_destroy(self.h)
            self.h = None


class Solver:
    def __init__(self, ctx, levels, nu_relax=3):
        self.lib = ctx.lib
        arr = (ctypes.c_void_p * len(levels))(*[l.h for l in levels])
        self.h = ctypes.c_void_p()
        check(self.lib.sa_gpu_solver_create(ctx.h, arr, len(levels), nu_relax, ctypes.byref(self.h)), "solver_create")
        self.n = levels[0].ND

    def vcycle(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n)
        check(self.lib.sa_gpu_vcycle(self.h, _dp(b), _dp(x)), "vcycle")
        return x

    def pcg(self, b, x0=None, maxiter=1000, rtol=1e-12, atol=0.0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        it, hl = ctypes.c_int(), ctypes.c_int()
        hist = np.zeros(maxiter + 2)
        check(self.lib.sa_gpu_pcg(self.h, _dp(b), _dp(x), maxiter, ctypes.c_double(rtol), ctypes.c_double(atol),
                                  ctypes.byref(it), _dp(hist), len(hist), ctypes.byref(hl)), "pcg")
        return x, it.value, hist[: hl.value]

    def close(self):
        if self.h:
            self.lib.sa_gpu_solver_destroy(self.h)
            self.h = None
