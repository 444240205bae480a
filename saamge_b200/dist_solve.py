"""Row-partitioned multi-GPU V-cycle / PCG solve -- thin wrapper.

The solve itself lives in the CUDA library (saamge_b200/csrc/dist.cu, C ABI `sa_gpu_dist_*`):
every rank holds the hierarchy (built with the AE loop sharded over the ranks) and owns a
contiguous row range of every level's vectors; every SpMV / smoother step / restriction /
prolongation runs on the rank's rows after a halo exchange of packed boundary lists (grouped
ncclSend / ncclRecv over NVLink), the dots are all-reduced into device scalars and the host
reads one scalar per iteration.  The algorithm is kalchev_pcg + tg_cycle_atb unchanged
(amg/src/mfem_addons.cpp:106-248, amg/src/tg.cpp:91-132), so iteration counts match the
single-GPU solve.  torch.distributed is only used to hand the NCCL unique id to the ranks."""
import ctypes

import numpy as np

from . import gpu_lib, host_lib

_ip = ctypes.POINTER(ctypes.c_int)
_dp = ctypes.POINTER(ctypes.c_double)


def _ranges(n, world):
    """Row ranges of the ranks: [n q / world, n (q + 1) / world)."""
    return [(n * q) // world for q in range(world + 1)]


def halo_plan_lists(I, J, world, rank, row_part, col_part):
    """Exchange lists of `rank` for the CSR pattern (I, J): (send, recv) with send[q] / recv[q]
    the ascending column indices sent to / received from rank q (host-only entry point
    sa_gpu_halo_plan; needs no GPU)."""
    g = gpu_lib()
    I = np.ascontiguousarray(I, dtype=np.int32)
    J = np.ascontiguousarray(J, dtype=np.int32)
    rp = np.ascontiguousarray(row_part, dtype=np.int32)
    cp = np.ascontiguousarray(col_part, dtype=np.int32)
    sc = np.zeros(world, dtype=np.int32)
    rc = np.zeros(world, dtype=np.int32)
    ns, nr = ctypes.c_int(), ctypes.c_int()
    dummy = np.zeros(1, dtype=np.int32)

    def call(si, scap, ri, rcap):
        return g.sa_gpu_halo_plan(len(I) - 1, I.ctypes.data_as(_ip), J.ctypes.data_as(_ip), world, rank,
                                  rp.ctypes.data_as(_ip), cp.ctypes.data_as(_ip), sc.ctypes.data_as(_ip),
                                  rc.ctypes.data_as(_ip), si.ctypes.data_as(_ip), scap,
                                  ri.ctypes.data_as(_ip), rcap, ctypes.byref(ns), ctypes.byref(nr))

    r = call(dummy, 0, dummy, 0)
    assert r in (0, 2), g.sa_gpu_last_error()
    si = np.zeros(max(1, ns.value), dtype=np.int32)
    ri = np.zeros(max(1, nr.value), dtype=np.int32)
    r = call(si, ns.value, ri, nr.value)
    assert r == 0, g.sa_gpu_last_error()
    so = np.concatenate([[0], np.cumsum(sc)])
    ro = np.concatenate([[0], np.cumsum(rc)])
    return ([si[so[q]:so[q + 1]].copy() for q in range(world)],
            [ri[ro[q]:ro[q + 1]].copy() for q in range(world)])


class DistSolver:
    """sa_gpu_dist_solver over the hierarchy `hier` of this rank; `dist` = torch.distributed
    (initialised, NCCL backend) or None for a single rank."""

    def __init__(self, hier, dist, group=None):
        self.g = g = gpu_lib()
        self.h = h = host_lib()
        self.dist, self.group = dist, group
        # (dist=None: a single rank -- the same device-scalar PCG without any exchange)
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if dist else (0, 1)
        h.sa_drv_ml_gpu_solver.restype = ctypes.c_void_p
        h.sa_drv_ml_gpu_solver.argtypes = [ctypes.c_void_p]
        h.sa_drv_ctx.restype = ctypes.c_void_p
        self.ctx = ctypes.c_void_p(h.sa_drv_ctx())
        self.solver = ctypes.c_void_p(h.sa_drv_ml_gpu_solver(hier.handle))
        self.n = int(hier.scalar("ND", 0))
        g.sa_gpu_nccl_unique_id.argtypes = [ctypes.c_void_p]
        g.sa_gpu_comm_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_void_p)]
        g.sa_gpu_comm_destroy.argtypes = [ctypes.c_void_p]
        g.sa_gpu_comm_destroy.restype = None
        g.sa_gpu_dist_solver_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.POINTER(ctypes.c_void_p)]
        g.sa_gpu_dist_solver_destroy.argtypes = [ctypes.c_void_p]
        g.sa_gpu_dist_solver_destroy.restype = None
        g.sa_gpu_dist_pcg.argtypes = [ctypes.c_void_p, _dp, _dp, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_int, _ip, _dp, ctypes.c_int, _ip, _dp]
        g.sa_gpu_dist_solver_stats.argtypes = [ctypes.c_void_p, _ip, _ip, ctypes.POINTER(ctypes.c_long),
                                               ctypes.POINTER(ctypes.c_long)]
        # the NCCL unique id goes from rank 0 to everyone through the launcher's process group
        ident = ctypes.create_string_buffer(128)
        if self.world > 1:
            box = [None]
            if self.rank == 0:
                assert g.sa_gpu_nccl_unique_id(ident) == 0, g.sa_gpu_last_error()
                box = [ident.raw]
            dist.broadcast_object_list(box, src=0, group=group)
            ident = ctypes.create_string_buffer(box[0], 128)
        comm = ctypes.c_void_p()
        rc = g.sa_gpu_comm_create(self.ctx, ident, self.world, self.rank, ctypes.byref(comm))
        assert rc == 0, g.sa_gpu_last_error()
        self.comm = comm
        d = ctypes.c_void_p()
        rc = g.sa_gpu_dist_solver_create(self.solver, self.comm, ctypes.byref(d))
        assert rc == 0, g.sa_gpu_last_error()
        self.d = d
        self.solve_seconds = 0.0

    @property
    def halo_calls(self):
        return self.stats()[2]

    def stats(self):
        r0, r1 = ctypes.c_int(), ctypes.c_int()
        hc, hd = ctypes.c_long(), ctypes.c_long()
        rc = self.g.sa_gpu_dist_solver_stats(self.d, ctypes.byref(r0), ctypes.byref(r1), ctypes.byref(hc),
                                             ctypes.byref(hd))
        assert rc == 0, self.g.sa_gpu_last_error()
        return r0.value, r1.value, hc.value, hd.value

    def level_info(self):
        """[(rows, nnz(A), nnz(P))] per level."""
        out = []
        self.g.sa_gpu_dist_solver_level_info.argtypes = [ctypes.c_void_p, ctypes.c_int, _ip,
                                                         ctypes.POINTER(ctypes.c_long),
                                                         ctypes.POINTER(ctypes.c_long)]
        l = 0
        while True:
            n, a, p = ctypes.c_int(), ctypes.c_long(), ctypes.c_long()
            if self.g.sa_gpu_dist_solver_level_info(self.d, l, ctypes.byref(n), ctypes.byref(a), ctypes.byref(p)):
                break
            out.append((n.value, a.value, p.value))
            l += 1
        return out

    def bytes_per_iteration(self, degree=10):
        """Algorithmic HBM bytes of one PCG iteration: per level
        (2 deg + 1) A-SpMV equivalents (12 nnz + 20 rows, + 24 rows per fused smoother step) and
        two P-SpMVs; plus one A-SpMV and ~56 rows of vector traffic on the finest level."""
        tot = 0.0
        info = self.level_info()
        for (n, a, p) in info:
            spmv = 12.0 * a + 20.0 * n
            tot += (2 * degree + 1) * spmv + 2 * degree * 24.0 * n + 2 * (12.0 * p + 20.0 * n)
        n0, a0, _ = info[0]
        return tot + 12.0 * a0 + 20.0 * n0 + 56.0 * n0

    def pcg(self, b_host, maxiter=1000, rtol=1e-12, atol=0.0, gather=False):
        """kalchev_pcg, x0 = 0.  Returns (x, iters, brr); x: full-length host vector holding the
        rank's rows (the whole solution with gather=True)."""
        b = np.ascontiguousarray(b_host, dtype=np.float64)
        assert len(b) == self.n
        x = np.zeros(self.n)
        hist = np.zeros(maxiter + 2)
        it, hl = ctypes.c_int(), ctypes.c_int()
        secs = ctypes.c_double()
        rc = self.g.sa_gpu_dist_pcg(self.d, b.ctypes.data_as(_dp), x.ctypes.data_as(_dp), maxiter, rtol, atol,
                                    1 if gather else 0, ctypes.byref(it), hist.ctypes.data_as(_dp), len(hist),
                                    ctypes.byref(hl), ctypes.byref(secs))
        assert rc == 0, self.g.sa_gpu_last_error()
        self.solve_seconds = secs.value
        return x, it.value, hist[: hl.value].tolist()

    def gather_solution(self, x):
        """Full solution on every rank from the per-rank pieces returned by pcg()."""
        import torch

        r0, r1, _, _ = self.stats()
        full = torch.zeros(self.n, dtype=torch.float64, device="cuda")
        full[r0:r1] = torch.as_tensor(x[r0:r1]).to("cuda")
        if self.world > 1:
            self.dist.all_reduce(full, group=self.group)
        return full.cpu().numpy()

    def close(self):
        if self.d:
            self.g.sa_gpu_dist_solver_destroy(self.d)
            self.d = None
        if self.comm:
            self.g.sa_gpu_comm_destroy(self.comm)
            self.comm = None
