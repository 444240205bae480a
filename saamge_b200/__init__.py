"""saamge_b200 -- B200-native spectral AMGe hot path (LLNL/saamge's local spectral
stage + V-cycle/PCG solve), Python loader.

The product is the CUDA library ``lib/libsaamge_b200.so`` (C ABI in
``include/saamge_b200.h``) and the C++ host mirror of the reference's tg_/ml_
interface ``lib/libsaamge_host.so``.  This module only loads them through ctypes
and wraps the driver C API (``include/saamge_b200_driver.h``) that tests and
``bench.py`` use.  There is no CPU fallback: building a hierarchy without a CUDA
device aborts with the library's error message.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
GPU_LIB = os.path.join(LIB_DIR, "libsaamge_b200.so")
HOST_LIB = os.path.join(LIB_DIR, "libsaamge_host.so")

_gpu = None
_host = None


class Params(ctypes.Structure):
    """sa_drv_params_t (include/saamge_b200_driver.h)."""

    _fields_ = [
        ("num_levels", ctypes.c_int),
        ("first_elems_per_agg", ctypes.c_int),
        ("elems_per_agg", ctypes.c_int),
        ("first_nu_pro", ctypes.c_int),
        ("nu_pro", ctypes.c_int),
        ("nu_relax", ctypes.c_int),
        ("first_theta", ctypes.c_double),
        ("theta", ctypes.c_double),
        ("avoid_ess_bdr_dofs", ctypes.c_int),
        ("partition_kind", ctypes.c_int),
        ("block", ctypes.c_int * 3),
        ("coarse_block", ctypes.c_int),
        ("testmesh_inject", ctypes.c_int),
        ("smooth_drop_tol", ctypes.c_double),
        ("correct_nullspace", ctypes.c_int),
    ]


def gpu_lib():
    """The CUDA library (C ABI).  Raises if it has not been built."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_LIB):
            raise RuntimeError(
                "saamge_b200: %s is missing -- run `make` (or __graft_entry__.build()); "
                "there is no fallback path" % GPU_LIB
            )
        _gpu = ctypes.CDLL(GPU_LIB, mode=ctypes.RTLD_GLOBAL)
        _gpu.sa_gpu_last_error.restype = ctypes.c_char_p
        _gpu.sa_gpu_ctx_launch_count.restype = ctypes.c_int64
        _gpu.sa_gpu_ctx_launch_count.argtypes = [ctypes.c_void_p]
        _gpu.sa_gpu_ctx_timer.restype = ctypes.c_double
        _gpu.sa_gpu_ctx_timer.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _gpu.sa_gpu_bench_spmv.restype = ctypes.c_double
        _gpu.sa_gpu_bench_spmv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _gpu.sa_gpu_bench_smoother.restype = ctypes.c_double
        _gpu.sa_gpu_bench_smoother.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _gpu.sa_gpu_bench_fp64_peak.restype = ctypes.c_double
        _gpu.sa_gpu_bench_fp64_peak.argtypes = [ctypes.c_void_p]
    return _gpu


def host_lib():
    global _host
    if _host is None:
        gpu_lib()
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("saamge_b200: %s is missing -- run `make`" % HOST_LIB)
        h = ctypes.CDLL(HOST_LIB, mode=ctypes.RTLD_GLOBAL)
        h.sa_drv_default_params.argtypes = [ctypes.POINTER(Params)]
        h.sa_drv_problem_create.restype = ctypes.c_void_p
        h.sa_drv_problem_create.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_uint64]
        h.sa_drv_problem_destroy.argtypes = [ctypes.c_void_p]
        h.sa_drv_problem_partition.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params)]
        h.sa_drv_ml_build.restype = ctypes.c_void_p
        h.sa_drv_ml_build.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_int]
        h.sa_drv_ml_pcg.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        h.sa_drv_ml_download.argtypes = [ctypes.c_void_p]
        h.sa_drv_hier_destroy.argtypes = [ctypes.c_void_p]
        h.sa_drv_get.argtypes = [
            ctypes.c_void_p,
            ctypes.c_char_p,
            ctypes.c_int,
            ctypes.POINTER(ctypes.c_void_p),
            ctypes.POINTER(ctypes.c_int64),
            ctypes.POINTER(ctypes.c_int),
        ]
        h.sa_drv_get_scalar.restype = ctypes.c_double
        h.sa_drv_get_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        if hasattr(h, "sa_drv_bench_create"):
            h.sa_drv_bench_create.restype = ctypes.c_void_p
            h.sa_drv_bench_create.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_int]
            h.sa_drv_bench_destroy.argtypes = [ctypes.c_void_p]
            h.sa_drv_bench_step.restype = ctypes.c_double
            h.sa_drv_bench_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            h.sa_drv_bench_scalar.restype = ctypes.c_double
            h.sa_drv_bench_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        _host = h
    return _host


_DTYPES = {0: np.int32, 1: np.int64, 2: np.float64, 3: np.int8}


def _get(handle, name, level=0):
    h = host_lib()
    ptr = ctypes.c_void_p()
    cnt = ctypes.c_int64()
    dt = ctypes.c_int()
    rc = h.sa_drv_get(handle, name.encode(), level, ctypes.byref(ptr), ctypes.byref(cnt), ctypes.byref(dt))
    if rc != 0:
        raise KeyError("%s (level %d): rc=%d" % (name, level, rc))
    n = cnt.value
    dtype = _DTYPES[dt.value]
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dtype).copy()


def default_params(**kw):
    p = Params()
    host_lib().sa_drv_default_params(ctypes.byref(p))
    for k, v in kw.items():
        if k == "block":
            for i in range(3):
                p.block[i] = v[i]
        else:
            setattr(p, k, v)
    return p


class Problem:
    """Structured Q1/Q2 diffusion problem + finest-level partition (host inputs)."""

    def __init__(self, dim, n, order=1, coef_kind=0, contrast=1e6, seed=12345, nyz=None):
        ny, nz = (n, n) if nyz is None else nyz
        self.handle = host_lib().sa_drv_problem_create(dim, n, ny, nz, order, coef_kind, contrast, seed)
        self.dim, self.n, self.order = dim, n, order

    @classmethod
    def from_matrix(cls, A, elems_per_agg, b=None, isolated=()):
        """Algebraic entry (amg/test/algebraic/algebraic.cpp:219-262): cells = dofs, agglomerates
        = METIS parts of the matrix graph, `isolated` dofs become agglomerates of their own.
        A: scipy CSR.  The problem comes back already partitioned."""
        import scipy.sparse as sp

        A = sp.csr_matrix(A)
        A.sort_indices()
        h = host_lib()
        h.sa_drv_problem_from_matrix.restype = ctypes.c_void_p
        h.sa_drv_problem_from_matrix.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        I = np.ascontiguousarray(A.indptr, dtype=np.int32)
        J = np.ascontiguousarray(A.indices, dtype=np.int32)
        V = np.ascontiguousarray(A.data, dtype=np.float64)
        iso = np.ascontiguousarray(list(isolated), dtype=np.int32)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        self = cls.__new__(cls)
        self.handle = h.sa_drv_problem_from_matrix(A.shape[0], I.ctypes.data, J.ctypes.data, V.ctypes.data,
                                                   None if bb is None else bb.ctypes.data, elems_per_agg,
                                                   iso.ctypes.data if len(iso) else None, len(iso))
        self.dim, self.n, self.order = 0, A.shape[0], 1
        return self

    def partition(self, params):
        return host_lib().sa_drv_problem_partition(self.handle, ctypes.byref(params))

    def get(self, name, level=0):
        return _get(self.handle, name, level)

    def scalar(self, name, level=0):
        return host_lib().sa_drv_get_scalar(self.handle, name.encode(), level)

    def close(self):
        if self.handle:
            host_lib().sa_drv_problem_destroy(self.handle)
            self.handle = None


class Hierarchy:
    """Handle on a built hierarchy (product or oracle) -- read access only."""

    def __init__(self, handle):
        self.handle = handle

    def get(self, name, level=0):
        return _get(self.handle, name, level)

    def scalar(self, name, level=0):
        return host_lib().sa_drv_get_scalar(self.handle, name.encode(), level)

    def times(self):
        keys = bytes(self.get("time_keys").astype(np.uint8)).decode().split("\n")
        return {k: self.scalar("time." + k) for k in keys if k}

    def csr(self, name, level=0):
        import scipy.sparse as sp

        I = self.get(name + ".I", level)
        J = self.get(name + ".J", level)
        A = self.get(name + ".A", level)
        ncols = int(J.max()) + 1 if len(J) else 0
        if name in ("tent_interp", "interp"):
            ncols = int(self.scalar("NDc", level))
        elif name in ("Ac", "cn_Ac"):
            ncols = len(I) - 1
        return sp.csr_matrix((A, J, I), shape=(len(I) - 1, ncols))

    def close(self):
        if self.handle:
            host_lib().sa_drv_hier_destroy(self.handle)
            self.handle = None


_exchange_keepalive = []


class _DevArray:
    """Zero-copy view of device memory owned by the CUDA library for torch
    (torch.as_tensor understands __cuda_array_interface__)."""

    def __init__(self, ptr, count, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def make_exchange(dist, group=None):
    """exchange(gpu_level_pointer, a, b, nparts): combines the per-AE results of a local spectral
    stage every rank ran on its own AE range [a, b) so that all ranks hold the full set."""
    from . import sharding

    g = gpu_lib()
    world = dist.get_world_size(group)
    ip = ctypes.POINTER(ctypes.c_int)
    dp = ctypes.POINTER(ctypes.c_double)

    def exchange_host(level, a, b, nparts):
        # any backend (gloo in the CPU tests): device -> host -> collective -> host -> device
        lev = ctypes.c_void_p(level)
        n = np.zeros(nparts, dtype=np.int32)
        m = np.zeros(nparts, dtype=np.int32)
        g.sa_gpu_get_AE_sizes(lev, n.ctypes.data_as(ip))
        g.sa_gpu_get_spectral_counts(lev, m.ctypes.data_as(ip))
        m[:a] = 0
        m[b:] = 0
        ev = np.zeros(int(m.sum()))
        Z = np.zeros(int((m.astype(np.int64) * n).sum()))
        D = np.zeros(int(n.sum()))
        rc = g.sa_gpu_get_spectral(lev, ev.ctypes.data_as(dp), Z.ctypes.data_as(dp), D.ctypes.data_as(dp))
        assert rc == 0, g.sa_gpu_last_error()
        do = np.concatenate([[0], np.cumsum(n.astype(np.int64))])
        full = sharding.allgather_spectral((m[a:b], ev, Z, D[do[a]:do[b]]), dist, group)
        fm, fev, fZ, fD = [np.ascontiguousarray(x) for x in full]
        rc = g.sa_gpu_set_spectral(lev, 0, nparts, fm.astype(np.int32).ctypes.data_as(ip), fev.ctypes.data_as(dp),
                                   fZ.ctypes.data_as(dp), fD.ctypes.data_as(dp))
        assert rc == 0, g.sa_gpu_last_error()

    def exchange_device(level, a, b, nparts):
        # NCCL on the level's own device arrays (sa_gpu_spectral_gather_begin): the counts are
        # combined with one all-reduce, then every rank broadcasts its slice of the eigenvalues,
        # eigenvectors and D straight into the other ranks' arrays over NVLink
        import torch

        lev = ctypes.c_void_p(level)
        n = np.zeros(nparts, dtype=np.int32)
        m = np.zeros(nparts, dtype=np.int32)
        g.sa_gpu_get_AE_sizes(lev, n.ctypes.data_as(ip))
        g.sa_gpu_get_spectral_counts(lev, m.ctypes.data_as(ip))
        m[:a] = 0
        m[b:] = 0
        tm = torch.from_numpy(m).cuda()
        dist.all_reduce(tm, group=group)
        rng = torch.tensor([a, b], dtype=torch.int64, device="cuda")
        rngs = [torch.zeros_like(rng) for _ in range(world)]
        dist.all_gather(rngs, rng, group=group)
        mfull = np.ascontiguousarray(tm.cpu().numpy(), dtype=np.int32)
        ranges = [(int(r[0]), int(r[1])) for r in torch.stack(rngs).cpu().numpy()]
        pe, pz, pd = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        rc = g.sa_gpu_spectral_gather_begin(lev, a, b, mfull.ctypes.data_as(ip), ctypes.byref(pe), ctypes.byref(pz),
                                            ctypes.byref(pd))
        assert rc == 0, g.sa_gpu_last_error()
        n64, m64 = n.astype(np.int64), mfull.astype(np.int64)
        offs = (np.concatenate([[0], np.cumsum(m64)]), np.concatenate([[0], np.cumsum(m64 * n64)]),
                np.concatenate([[0], np.cumsum(n64)]))
        for ptr, off in zip((pe, pz, pd), offs):
            if not ptr.value or off[-1] == 0:
                continue
            T = torch.as_tensor(_DevArray(ptr.value, int(off[-1])), device="cuda")
            for r, (ar, br) in enumerate(ranges):
                if off[br] > off[ar]:
                    dist.broadcast(T[int(off[ar]):int(off[br])], src=r, group=group)
        torch.cuda.synchronize()

    def exchange(level, a, b, nparts):
        if dist.get_backend(group) == "nccl" and not os.environ.get("SA_SHARD_HOST_EXCHANGE"):
            exchange_device(level, a, b, nparts)
        else:
            exchange_host(level, a, b, nparts)

    return exchange


def library_comm(dist, group=None, device=None):
    """sa_gpu_comm (NCCL communicator owned by the CUDA library) over the ranks of `dist`:
    rank 0 makes the 128-byte unique id, the launcher's process group hands it to everyone.
    Collective.  `device`: this rank's GPU (default: torch's current device) -- the library's
    context is created there if the process has none yet.  Returns a ctypes.c_void_p; free with
    gpu_lib().sa_gpu_comm_destroy."""
    g, h = gpu_lib(), host_lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        import torch

        device = torch.cuda.current_device()
    h.sa_drv_ctx_on.restype = ctypes.c_void_p
    h.sa_drv_ctx_on.argtypes = [ctypes.c_int]
    h.sa_drv_ctx_on(int(device))
    h.sa_drv_ctx.restype = ctypes.c_void_p
    g.sa_gpu_nccl_unique_id.argtypes = [ctypes.c_void_p]
    g.sa_gpu_comm_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p)]
    g.sa_gpu_comm_destroy.argtypes = [ctypes.c_void_p]
    g.sa_gpu_comm_destroy.restype = None
    ident = ctypes.create_string_buffer(128)
    box = [None]
    if rank == 0:
        assert g.sa_gpu_nccl_unique_id(ident) == 0, g.sa_gpu_last_error()
        box = [ident.raw]
    dist.broadcast_object_list(box, src=0, group=group)
    ident = ctypes.create_string_buffer(box[0], 128)
    comm = ctypes.c_void_p()
    rc = g.sa_gpu_comm_create(ctypes.c_void_p(h.sa_drv_ctx()), ident, world, rank, ctypes.byref(comm))
    assert rc == 0, g.sa_gpu_last_error()
    return comm


_owner_comm = []


def enable_sharding(dist, group=None, mode="owner", device=None):
    """Shard the setup over the ranks of `dist` (torch.distributed; NCCL on GPUs).
    mode "owner" (default with NCCL): every rank computes its AE range and keeps its
    eigenvectors; the tentative prolongator is built by the MIS owners after one all-to-all-v of
    MIS-restricted blocks, coarse element matrices by AE range, P smoothing / RAP by row blocks --
    all inside the CUDA library over its own NCCL communicator (sa_gpu_dist_*).
    mode "replicate": only the AE loop is sharded; the per-AE results are all-gathered
    (make_exchange) and every later stage runs replicated on every rank (round 1's scheme; the
    only one available with the gloo backend)."""
    h = host_lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if mode == "owner" and dist.get_backend(group) == "nccl" and world > 1 \
            and not os.environ.get("SA_SHARD_REPLICATE"):
        comm = library_comm(dist, group, device)
        _owner_comm.append(comm)
        h.sa_drv_set_sharding_comm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        h.sa_drv_set_sharding_comm(comm, rank, world)
        return
    CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int)
    cb = CB(make_exchange(dist, group))
    _exchange_keepalive.append(cb)
    h.sa_drv_set_sharding.argtypes = [ctypes.c_int, ctypes.c_int, CB]
    h.sa_drv_set_sharding(rank, world, cb)


def sharding_stats():
    """Bytes moved by the owner-sharded setup stages of this rank since enable_sharding:
    MIS blocks sent / received, gathered MIS bases, product rows received."""
    h = host_lib()
    out = (ctypes.c_double * 8)()
    h.sa_drv_sharding_stats(out)
    return {"mis_blocks_sent": out[0], "mis_blocks_received": out[1], "mis_bases_allreduced": out[2],
            "spgemm_rows_received": out[3]}


def disable_sharding():
    h = host_lib()
    h.sa_drv_set_sharding_comm.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    h.sa_drv_set_sharding_comm(None, 0, 1)
    while _owner_comm:
        gpu_lib().sa_gpu_comm_destroy(_owner_comm.pop())
    CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int)
    h.sa_drv_set_sharding.argtypes = [ctypes.c_int, ctypes.c_int, CB]
    h.sa_drv_set_sharding(0, 1, CB())


def ml_build(problem, params, device=0):
    """ml_produce_data through the B200 path; returns a Hierarchy."""
    h = host_lib().sa_drv_ml_build(problem.handle, ctypes.byref(params), device)
    return Hierarchy(h)


def ml_build_algebraic(problem, params, device=0):
    """tg_produce_data_algebraic / ml form (amg/src/tg.cpp:862-886) for a Problem.from_matrix."""
    h = host_lib()
    h.sa_drv_ml_build_algebraic.restype = ctypes.c_void_p
    h.sa_drv_ml_build_algebraic.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_int]
    return Hierarchy(h.sa_drv_ml_build_algebraic(problem.handle, ctypes.byref(params), device))


def ml_pcg(hier, maxiter=1000, rtol=1e-12, atol=0.0):
    return host_lib().sa_drv_ml_pcg(hier.handle, maxiter, rtol, atol)


def set_operator_values(problem, values):
    """New values (same sparsity pattern) for the problem's operator -- the input of
    ml_update_operators (a time-dependent coefficient)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    h = host_lib()
    h.sa_drv_problem_set_A_values.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    rc = h.sa_drv_problem_set_A_values(problem.handle, v.ctypes.data, len(v))
    assert rc == 0, "operator values: wrong length"


def ml_update_operators(hier, resmooth_interp=True):
    """adapt_update_operators (amg/src/adapt.cpp:171-216): the hierarchy follows the problem's
    current operator values without new eigensolves."""
    h = host_lib()
    h.sa_drv_ml_update_operators.argtypes = [ctypes.c_void_p, ctypes.c_int]
    return h.sa_drv_ml_update_operators(hier.handle, 1 if resmooth_interp else 0)


def ml_download(hier):
    return host_lib().sa_drv_ml_download(hier.handle)
