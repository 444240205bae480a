"""Row-partitioned multi-GPU V-cycle / PCG solve.

Every rank holds the hierarchy (built with the AE loop sharded over the ranks) and owns a
contiguous row range of every level's vectors.  Each SpMV / smoother step / restriction /
prolongation runs the library's CSR kernel on the rank's rows only
(sa_gpu_dev_spmv on device pointers) after a halo exchange that brings in exactly the
column range the local rows reference (grouped NCCL send/recv between the ranks whose row
ranges overlap it; on a slab-partitioned structured grid that is the two neighbouring
slabs).  Dots are all-reduced.  The algorithm is kalchev_pcg + tg_cycle_atb unchanged
(amg/src/mfem_addons.cpp:106-248, amg/src/tg.cpp:91-132), so iteration counts match the
single-GPU solve.  torch / torch.distributed are plumbing: device vectors, stream, NCCL."""
import ctypes

import numpy as np
import torch

from . import gpu_lib, host_lib

_ip = ctypes.POINTER(ctypes.c_int)
_dp = ctypes.POINTER(ctypes.c_double)


def _ranges(n, world):
    b = [(n * q) // world for q in range(world + 1)]
    return b


def halo_plan(allneed, part, me):
    """Contiguous pieces to send / receive so that rank `me` gets the column range it needs.
    allneed[q] = (lo, hi) needed by rank q; part[q]:part[q+1] = range owned by rank q.
    Returns (sends, recvs) as lists of (peer, lo, hi)."""
    world = len(part) - 1
    sends, recvs = [], []
    my_lo, my_hi = part[me], part[me + 1]
    for q in range(world):
        if q == me:
            continue
        lo, hi = max(allneed[q][0], my_lo), min(allneed[q][1], my_hi)
        if hi > lo:
            sends.append((q, lo, hi))
        lo, hi = max(allneed[me][0], part[q]), min(allneed[me][1], part[q + 1])
        if hi > lo:
            recvs.append((q, lo, hi))
    return sends, recvs


class _Mat:
    """A level matrix on the device + this rank's row range + halo plan."""

    def __init__(self, solver, level_handle, which, row_part, col_part):
        g = solver.g
        I, J, A = _ip(), _ip(), _dp()
        rows, cols, nnz = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = g.sa_gpu_level_dev_csr(level_handle, which, ctypes.byref(I), ctypes.byref(J), ctypes.byref(A),
                                    ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(nnz))
        assert rc == 0, g.sa_gpu_last_error()
        self.I, self.J, self.A = I, J, A
        self.rows, self.cols, self.nnz = rows.value, cols.value, nnz.value
        self.row_part, self.col_part = row_part, col_part
        r = solver.rank
        self.r0, self.r1 = row_part[r], row_part[r + 1]
        # column range referenced by the local rows (host copy of the pattern, setup only)
        hI = np.zeros(self.rows + 1, dtype=np.int32)
        hJ = np.zeros(max(self.nnz, 1), dtype=np.int32)
        rc = g.sa_gpu_get_csr(level_handle, which, hI.ctypes.data_as(_ip), hJ.ctypes.data_as(_ip), None)
        assert rc == 0, g.sa_gpu_last_error()
        seg = hJ[hI[self.r0]:hI[self.r1]]
        self.local_nnz = int(hI[self.r1] - hI[self.r0])
        lo, hi = (int(seg.min()), int(seg.max()) + 1) if len(seg) else (0, 0)
        self.need = (lo, hi)
        base = ctypes.cast(I, ctypes.c_void_p).value or 0
        self.I_row0 = ctypes.c_void_p(base + 4 * self.r0)
        self.avg = self.local_nnz / max(1, self.r1 - self.r0)
        self.plan = None  # filled by DistSolver._make_plan

    def ptr_I(self):
        return self.I_row0


class DistSolver:
    def __init__(self, hier, dist, group=None):
        self.g = gpu_lib()
        self.h = host_lib()
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        h = self.h
        h.sa_drv_ml_gpu_level.restype = ctypes.c_void_p
        h.sa_drv_ml_gpu_level.argtypes = [ctypes.c_void_p, ctypes.c_int]
        h.sa_drv_ml_gpu_solver.restype = ctypes.c_void_p
        h.sa_drv_ml_gpu_solver.argtypes = [ctypes.c_void_p]
        h.sa_drv_ctx.restype = ctypes.c_void_p
        self.ctx = ctypes.c_void_p(h.sa_drv_ctx())
        self.g.sa_gpu_ctx_stream.restype = ctypes.c_void_p
        self.g.sa_gpu_ctx_stream.argtypes = [ctypes.c_void_p]
        self.stream = torch.cuda.ExternalStream(self.g.sa_gpu_ctx_stream(self.ctx))
        self.solver = ctypes.c_void_p(h.sa_drv_ml_gpu_solver(hier.handle))
        nlev = int(hier.scalar("num_rels"))
        self.nlev = nlev
        deg, nc = ctypes.c_int(), ctypes.c_int()
        roots = np.zeros(64)
        rc = self.g.sa_gpu_solver_info(self.solver, ctypes.byref(deg), roots.ctypes.data_as(_dp), 64, ctypes.byref(nc))
        assert rc == 0
        self.degree, self.roots, self.nc = deg.value, roots[: deg.value].copy(), nc.value
        self.levels = []
        sizes = [int(hier.scalar("ND", l)) for l in range(nlev)] + [self.nc]
        self.parts = [_ranges(n, self.world) for n in sizes]
        dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.stream(self.stream):
            for l in range(nlev):
                lh = ctypes.c_void_p(h.sa_drv_ml_gpu_level(hier.handle, l))
                L = {}
                L["A"] = _Mat(self, lh, 0, self.parts[l], self.parts[l])
                L["P"] = _Mat(self, lh, 2, self.parts[l], self.parts[l + 1])
                L["R"] = _Mat(self, lh, 3, self.parts[l + 1], self.parts[l])
                dinv = _dp()
                rc = self.g.sa_gpu_level_dev_dinv(lh, ctypes.byref(dinv))
                assert rc == 0
                L["dinv"] = dinv
                n = sizes[l]
                for nm in ("b", "xa", "xb", "r"):
                    L[nm] = torch.zeros(n, dtype=torch.float64, device=dev)
                self.levels.append(L)
            self.bc = torch.zeros(max(self.nc, 1), dtype=torch.float64, device=dev)
            self.xc = torch.zeros(max(self.nc, 1), dtype=torch.float64, device=dev)
        for l in range(nlev):
            for nm in ("A", "P", "R"):
                self._make_plan(self.levels[l][nm])
        self.halo_calls = 0

    # -- halo plans: who needs which contiguous piece of whose range
    def _make_plan(self, M):
        need = torch.tensor([M.need[0], M.need[1]], dtype=torch.int64, device="cuda")
        allneed = [torch.zeros_like(need) for _ in range(self.world)]
        self.dist.all_gather(allneed, need, group=self.group)
        allneed = [tuple(int(v) for v in t.tolist()) for t in allneed]
        M.plan = halo_plan(allneed, M.col_part, self.rank)

    def _halo(self, vec, M):
        sends, recvs = M.plan
        if not sends and not recvs:
            return
        ops = []
        for q, lo, hi in sends:
            ops.append(self.dist.P2POp(self.dist.isend, vec[lo:hi], q, self.group))
        for q, lo, hi in recvs:
            ops.append(self.dist.P2POp(self.dist.irecv, vec[lo:hi], q, self.group))
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()
        self.halo_calls += 1

    def _spmv(self, M, mode, x, y, xrow=None, b=None, dinv=None, mult=0.0):
        n = M.r1 - M.r0
        if n <= 0:
            return
        off = 8 * M.r0
        xr = ctypes.c_void_p((xrow.data_ptr() if xrow is not None else x.data_ptr()) + off)
        bp = ctypes.c_void_p(b.data_ptr() + off) if b is not None else None
        dv = ctypes.c_void_p((ctypes.cast(dinv, ctypes.c_void_p).value or 0) + off) if dinv is not None else None
        rc = self.g.sa_gpu_dev_spmv(self.ctx, mode, n, ctypes.c_double(M.avg), M.ptr_I(), M.J, M.A,
                                    ctypes.c_void_p(x.data_ptr()), xr, bp, dv, ctypes.c_double(mult),
                                    ctypes.c_void_p(y.data_ptr() + off))
        assert rc == 0, self.g.sa_gpu_last_error()

    def _smooth(self, L, b, xcur, xalt, x_is_zero):
        A = L["A"]
        for i in range(self.degree):
            mult = 1.0 / self.roots[i]
            if x_is_zero and i == 0:
                self._spmv(A, 4, xcur, xalt, xrow=xcur, b=b, dinv=L["dinv"], mult=mult)
            else:
                self._halo(xcur, A)
                self._spmv(A, 3, xcur, xalt, xrow=xcur, b=b, dinv=L["dinv"], mult=mult)
            xcur, xalt = xalt, xcur
        return xcur, xalt

    def vcycle(self, l, b):
        """tg_cycle_atb on level l; b valid on the rank's rows; returns x (own rows valid)."""
        L = self.levels[l]
        xcur, xalt = self._smooth(L, b, L["xa"], L["xb"], True)
        A, P, R = L["A"], L["P"], L["R"]
        self._halo(xcur, A)
        self._spmv(A, 1, xcur, L["r"], b=b)
        self._halo(L["r"], R)
        if l + 1 < self.nlev:
            bc = self.levels[l + 1]["b"]
            self._spmv(R, 0, L["r"], bc)
            xc = self.vcycle(l + 1, bc)
            self._halo(xc, P)
        else:
            self.bc.zero_()
            self._spmv(R, 0, L["r"], self.bc)
            if self.world > 1:
                self.dist.all_reduce(self.bc, group=self.group)
            rc = self.g.sa_gpu_solver_dev_coarse(self.solver, ctypes.c_void_p(self.bc.data_ptr()),
                                                 ctypes.c_void_p(self.xc.data_ptr()))
            assert rc == 0
            xc = self.xc
        self._spmv(P, 2, xc, xcur)
        xcur, xalt = self._smooth(L, b, xcur, xalt, False)
        return xcur

    def _dot(self, a, b):
        r0, r1 = self.parts[0][self.rank], self.parts[0][self.rank + 1]
        t = torch.dot(a[r0:r1], b[r0:r1]).reshape(1)
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return float(t.item())

    def pcg(self, b_host, maxiter=1000, rtol=1e-12, atol=0.0):
        """kalchev_pcg (amg/src/mfem_addons.cpp:106-248), x0 = 0.  Returns (x, iters, brr)."""
        with torch.cuda.stream(self.stream):
            n = len(b_host)
            dev = self.levels[0]["b"].device
            r0, r1 = self.parts[0][self.rank], self.parts[0][self.rank + 1]
            b = torch.as_tensor(b_host, dtype=torch.float64).to(dev)
            x = torch.zeros(n, dtype=torch.float64, device=dev)
            r = torch.zeros_like(x)
            d = torch.zeros_like(x)
            z = torch.zeros_like(x)
            A = self.levels[0]["A"]
            sl = slice(r0, r1)
            r[sl] = b[sl]  # x0 = 0
            self.levels[0]["b"][sl] = r[sl]
            z[sl] = self.vcycle(0, self.levels[0]["b"])[sl]
            d[sl] = z[sl]
            nom = self._dot(z, r)
            brr = [nom]
            r0tol = max(nom * rtol, atol)
            if nom < r0tol:
                return x, -1, brr
            self._halo(d, A)
            self._spmv(A, 0, d, z)
            den = self._dot(z, d)
            if den == 0.0:
                return x, -1, brr
            iters = 0
            i = 1
            while i <= maxiter:
                alpha = nom / den
                x[sl] += alpha * d[sl]
                r[sl] -= alpha * z[sl]
                self.levels[0]["b"][sl] = r[sl]
                z[sl] = self.vcycle(0, self.levels[0]["b"])[sl]
                betanom = self._dot(r, z)
                brr.append(betanom)
                if betanom < 0.0:
                    iters = -i
                    break
                if betanom < r0tol:
                    iters = i
                    break
                beta = betanom / nom
                d[sl] = z[sl] + beta * d[sl]
                self._halo(d, A)
                self._spmv(A, 0, d, z)
                den = self._dot(d, z)
                nom = betanom
                i += 1
            if i > maxiter:
                iters = -(i - 1)
            torch.cuda.current_stream().synchronize()
            return x, iters, brr

    def gather_solution(self, x):
        """Full solution on every rank (own rows summed over ranks)."""
        r0, r1 = self.parts[0][self.rank], self.parts[0][self.rank + 1]
        full = torch.zeros_like(x)
        full[r0:r1] = x[r0:r1]
        if self.world > 1:
            self.dist.all_reduce(full, group=self.group)
        return full.cpu().numpy()
