#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 spectral-AMGe hot path.

Metric (BASELINE.json): agglomerate eigensolves/sec (+ setup & PCG solve time) on the
128^3 hex diffusion problem with a synthetic SPE10-like lognormal coefficient.

A "step" is one pass of the local spectral stage (SURVEY.md section 8a rows a2-a7:
assemble every AE matrix, weighted-l1 D, eigenpairs of A z = lambda D z with
lambda <= theta) over all ~40k agglomerates of the finest level.
  value : AEs / s with all inputs resident in HBM (CUDA events on the library's stream)
  e2e   : the same stage through the C ABI from HOST buffers: H2D of every input
          (element blocks, operator, tables) + compute + D2H of m / lambda / vectors
Extra keys: roofline (dominant kernel), roofline_spmv, cpu_baseline (the CPU oracle on
the box's host cores, bounded sample), full-hierarchy setup and PCG solve times.

Multi-GPU (torchrun, one rank per GPU): STRONG scaling of ONE 128^3 problem -- the AEs of the
level are dealt to the ranks (contiguous ranges balanced on n^3) and the per-AE results are
exchanged device to device inside the timed region; times are max-reduced over ranks.  The
`hierarchy` block times the whole multilevel setup (owner-sharded: sa_gpu_dist_* stages) and the
row-partitioned PCG on the N ranks, beside the CPU path.

--impl reference : the reference's CPU path (the oracle port: same algorithm, same
LAPACK dsygvx calls) on all host cores, bounded sample per step; rank 0 only.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (n, levels, first_elems_per_agg, elems_per_agg, METIS tile)
    "c3_128": dict(n=128, levels=4, fepa=52, epa=64, tile=32,
                   desc="3D diffusion 128^3 hex Q1, lognormal coefficient 1e6 contrast, ~40k METIS AEs (BASELINE configs[2])"),
    "c2_64": dict(n=64, levels=3, fepa=52, epa=64, tile=32,
                  desc="3D diffusion 64^3 hex Q1, lognormal coefficient 1e6 contrast, 3-level (BASELINE configs[1])"),
    "c5_256": dict(n=256, levels=4, fepa=52, epa=64, tile=32,
                   desc="3D diffusion 256^3 hex Q1, lognormal coefficient 1e6 contrast, 4-level, ~322k METIS AEs (BASELINE configs[4]; needs N >= 2: the owner-sharded RAP keeps nnz(A P) / N per rank)"),
    "small_32": dict(n=32, levels=3, fepa=52, epa=64, tile=32, desc="32^3 smoke workload"),
}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def merge(self, other):
        self.samples += other.samples

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for nm, v in zip(names, s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def make_problem(sab, wl, seed):
    w = WORKLOADS[wl]
    p = sab.default_params(num_levels=w["levels"], first_elems_per_agg=w["fepa"], elems_per_agg=w["epa"],
                           partition_kind=2, block=(w["tile"],) * 3)
    pr = sab.Problem(3, w["n"], coef_kind=1, contrast=1e6, seed=seed)
    nae = pr.partition(p)
    return pr, p, nae


def bench_config(args, pr, nae):
    """`config` of the JSON line -- the same dict in both arms (b200 / reference)."""
    w = WORKLOADS[args.workload]
    gb = (12.0 * pr.scalar("nnz") + 8.0 * pr.scalar("NE") * pr.scalar("ne") ** 2) / 1e9
    return {"workload": args.workload, "description": w["desc"], "n_AE_total": int(nae), "theta": 0.003,
            "partition": "METIS k-way on a %d^3 tile, replicated" % w["tile"],
            "l2": "inputs larger than L2 (element blocks + operator = %.2f GB per pass over the level)" % gb}


def cpu_sample(ou, pr, p, nae, target_s, procs):
    """Times the oracle's local spectral stage on a bounded AE sample (~target_s seconds) with
    `procs` single-threaded PROCESSES, each on its own contiguous slice of the sample -- the
    analogue of the reference's `mpirun -n P` (AEs are independent per rank,
    amg/src/interp.cpp:387).  Threads inside one process would understate the CPU: the LAPACK
    in this image (scipy's OpenBLAS) serialises concurrent calls from one process (measured:
    1412 dsygvx/s on 1 thread, 990/s on 8 threads).  Returns (AEs, wall seconds, processes).
    The children only run CPU code (fork happens after the problem is built; no CUDA calls)."""
    o = ou.oracle(1)
    n0 = min(nae, 64)
    t0 = o.sa_orc_time_local_spectral(pr.handle, ctypes.byref(p), 0, n0)
    rate1 = n0 / max(t0, 1e-9)
    if procs <= 1:
        ns = int(min(nae, max(n0, rate1 * target_s)))
        t = o.sa_orc_time_local_spectral(pr.handle, ctypes.byref(p), 0, ns)
        return ns, t, 1
    ns = int(min(nae, max(procs * 8, rate1 * procs * target_s)))
    bounds = [ns * i // procs for i in range(procs + 1)]
    sys.stdout.flush()
    sys.stderr.flush()
    # fork cost (page tables of a multi-GB parent) stays outside the timed region: the workers
    # report ready, then are released together
    ready_r, ready_w = os.pipe()
    go_r, go_w = os.pipe()
    pids = []
    for i in range(procs):
        pid = os.fork()
        if pid == 0:
            rc = 0
            try:
                os.write(ready_w, b"r")
                os.read(go_r, 1)
                o.sa_orc_time_local_spectral(pr.handle, ctypes.byref(p), bounds[i], bounds[i + 1])
            except BaseException:
                rc = 1
            os._exit(rc)
        pids.append(pid)
    got = 0
    while got < procs:
        got += len(os.read(ready_r, procs - got))
    t_start = time.time()
    os.write(go_w, b"g" * procs)
    bad = 0
    left = set(pids)
    deadline = t_start + max(60.0, 20.0 * target_s)
    while left:
        for pid in list(left):
            done, st = os.waitpid(pid, os.WNOHANG)
            if done:
                left.discard(pid)
                bad += 1 if st != 0 else 0
        if left:
            if time.time() > deadline:  # a worker hung (fork in a threaded process): give up
                for pid in left:
                    try:
                        os.kill(pid, 9)
                        os.waitpid(pid, 0)
                    except OSError:
                        pass
                bad += len(left)
                left = set()
            else:
                time.sleep(0.002)
    wall = time.time() - t_start
    for fd in (ready_r, ready_w, go_r, go_w):
        os.close(fd)
    if bad:
        raise RuntimeError("cpu_sample: %d worker processes failed" % bad)
    return ns, wall, procs


def _fork_map(fn, nworkers):
    """Runs fn(i) for i < nworkers in forked single-threaded processes that start together;
    returns the list of float results (the children only run CPU code)."""
    import multiprocessing as mp

    res = mp.RawArray(ctypes.c_double, nworkers)
    ready_r, ready_w = os.pipe()
    go_r, go_w = os.pipe()
    sys.stdout.flush()
    sys.stderr.flush()
    pids = []
    for i in range(nworkers):
        pid = os.fork()
        if pid == 0:
            rc = 0
            try:
                os.write(ready_w, b"r")
                os.read(go_r, 1)
                res[i] = float(fn(i))
            except BaseException:
                rc = 1
            os._exit(rc)
        pids.append(pid)
    got = 0
    while got < nworkers:
        got += len(os.read(ready_r, nworkers - got))
    os.write(go_w, b"g" * nworkers)
    bad = 0
    for pid in pids:
        _, st = os.waitpid(pid, 0)
        bad += 1 if st != 0 else 0
    for fd in (ready_r, ready_w, go_r, go_w):
        os.close(fd)
    if bad:
        raise RuntimeError("%d CPU worker processes failed" % bad)
    return list(res)


def cpu_hierarchy_baseline(ou, sab, H, pr, p, nae, level0_rate, procs, its_gpu):
    """CPU (oracle port) time of what the north star targets -- setup local spectral stage of
    EVERY level + PCG solve -- on the box's host cores, from bounded samples:
      level 0   : nae / (AE/s of the level-0 sample already measured: assemble + D + dsygvx)
      level l>0 : `procs` AE matrices of the level, downloaded from the GPU hierarchy
                  (sa_gpu_build_AE_stiff), D + dsygvx on one single-threaded process each, all
                  at once; core-seconds scaled by sum(n^3) of the level / sum(n^3) of the sample
                  and divided by the cores.  Levels whose AEs are too large to sample in seconds
                  are extrapolated with the previous level's measured seconds per n^3.
      PCG       : the oracle's kalchev_pcg + V-cycle (OpenMP over rows) on the DOWNLOADED
                  operators of the same hierarchy for 2 iterations, scaled to the GPU's count.
    The setup figure omits tentative P / RAP / topology on the CPU: it is a lower bound."""
    import numpy as np

    o = ou.oracle(1)  # the per-AE workers are single-threaded processes
    g = sab.gpu_lib()
    h = sab.host_lib()
    h.sa_drv_ml_gpu_level.restype = ctypes.c_void_p
    h.sa_drv_ml_gpu_level.argtypes = [ctypes.c_void_p, ctypes.c_int]
    o.sa_orc_time_dense_AE.restype = ctypes.c_double
    o.sa_orc_time_dense_AE.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_double,
                                       ctypes.POINTER(ctypes.c_int)]
    o.sa_orc_time_pcg_on_hierarchy.restype = ctypes.c_double
    o.sa_orc_time_pcg_on_hierarchy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                               ctypes.POINTER(ctypes.c_int)]
    ncoarsen = p.num_levels - 1
    detail = [{"level": 0, "cpu_s": nae / level0_rate, "how": "level-0 sample rate x %d AEs" % nae}]
    sec_per_n3 = None
    for l in range(1, ncoarsen):
        sizes = np.diff(H.get("AE_to_dof.I", l)).astype(np.int64)
        n3 = float((sizes.astype(float) ** 3).sum())
        nparts = len(sizes)
        if np.median(sizes) <= 2600:
            order = np.argsort(sizes)
            k = min(procs, nparts)
            pick = order[np.linspace(0, nparts - 1, k).astype(int)]
            lev = ctypes.c_void_p(h.sa_drv_ml_gpu_level(H.handle, l))
            mats = []
            for part in pick:
                n = int(sizes[part])
                A = np.zeros(n * n)
                rc = g.sa_gpu_build_AE_stiff(lev, int(part), A.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
                if rc != 0:
                    raise RuntimeError(g.sa_gpu_last_error().decode())
                mats.append((n, A))

            def work(i):
                n, A = mats[i]
                return o.sa_orc_time_dense_AE(n, A.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), p.theta, None)

            ts = _fork_map(work, k)
            n3s = float(sum(float(n) ** 3 for n, _ in mats))
            sec_per_n3 = sum(ts) / n3s
            cpu_s = sec_per_n3 * n3 / min(procs, nparts)
            detail.append({"level": l, "cpu_s": cpu_s, "how": "%d of %d AEs (n = %d..%d) on %d cores at once, "
                           "%.1f core-seconds, scaled by sum n^3" % (k, nparts, min(n for n, _ in mats),
                                                                    max(n for n, _ in mats), k, sum(ts))})
        elif sec_per_n3 is not None:
            cpu_s = sec_per_n3 * n3 / min(procs, nparts)
            detail.append({"level": l, "cpu_s": cpu_s, "how": "%d AEs (median n = %d): extrapolated with level %d's "
                           "measured seconds per n^3, %d cores busy" % (nparts, int(np.median(sizes)), l - 1,
                                                                        min(procs, nparts))})
    cpu_setup = sum(d["cpu_s"] for d in detail)
    sab.ml_download(H)
    o = ou.oracle(procs)  # the solve uses every core (OpenMP over rows)
    run_iters = 2
    it = ctypes.c_int()
    t = o.sa_orc_time_pcg_on_hierarchy(H.handle, run_iters, 1e-12, 0.0, ctypes.byref(it))
    cpu_pcg = t * (its_gpu + 1) / (run_iters + 1)
    return {"cpu_setup_s": cpu_setup, "cpu_setup_covers": "local spectral stage (assemble + D + dsygvx) of every "
            "level; tentative P / RAP / topology not included (lower bound)", "cpu_setup_detail": detail,
            "cpu_pcg_s": cpu_pcg, "cpu_pcg_how": "oracle kalchev_pcg + V-cycle on the downloaded operators, %d "
            "V-cycles in %.2f s on %d OpenMP threads, scaled to %d iterations" % (run_iters + 1, t,
                                                                                 o.sa_orc_num_threads(), its_gpu),
            "cores": procs}


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores; rank 0 only."""
    if rank != 0:
        return
    import oracle_util as ou
    import saamge_b200 as sab

    pr, p, nae = make_problem(sab, args.workload, 12345)
    procs = os.cpu_count() or 1
    per_step_s = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    ns = cores = 0
    for _ in range(max(1, args.warmup)):  # calibration + warm-up
        ns, t, cores = cpu_sample(ou, pr, p, nae, per_step_s, procs)
    tot = 0.0
    tot_ae = 0
    for _ in range(args.steps):
        ns, t, cores = cpu_sample(ou, pr, p, nae, per_step_s, procs)
        tot += t
        tot_ae += ns
    value = tot_ae / tot
    w = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "agglomerate eigensolves/sec", "value": value, "unit": "AE/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, pr, nae),
        "sample_AEs_per_step": ns,
        "cpu_baseline": {"value": value, "unit": "AE/s", "cores": cores, "kind": "port",
                         "sample": "first %d of %d AEs per step (assemble + D + dsygvx), one single-threaded process per core on contiguous AE slices (mpirun -n P analogue)" % (ns, nae)},
        "e2e": {"value": value, "unit": "AE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_128", choices=sorted(WORKLOADS))
    ap.add_argument("--no-hierarchy", action="store_true", help="skip the full setup + PCG extras")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import saamge_b200 as sab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import numpy as np

    from saamge_b200 import sharding

    h = sab.host_lib()
    gl = sab.gpu_lib()
    h.sa_drv_gpu_profile.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    h.sa_drv_bench_level.restype = ctypes.c_void_p
    h.sa_drv_bench_level.argtypes = [ctypes.c_void_p]
    h.sa_drv_ctx.restype = ctypes.c_void_p
    gl.sa_gpu_local_spectral.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    w = WORKLOADS[args.workload]
    t0 = time.time()
    # ONE problem: with several ranks every rank builds the same inputs and the level's AEs are
    # dealt to the ranks (strong scaling, BASELINE configs[2]: "setup sharded over agglomerates")
    pr, p, nae = make_problem(sab, args.workload, 12345)
    t_inputs = time.time() - t0
    B = h.sa_drv_bench_create(pr.handle, ctypes.byref(p), local_rank)
    theta = p.first_theta
    sizes = np.diff(pr.get("AE_to_dof.I")).astype(np.int64)
    a, b = (0, nae) if world == 1 else sharding.shard_ranges(sizes, world)[rank]
    share = float((sizes[a:b].astype(float) ** 3).sum() / (sizes.astype(float) ** 3).sum())
    ctxp = ctypes.c_void_p(h.sa_drv_ctx())
    lev = ctypes.c_void_p(h.sa_drv_bench_level(B))
    exchange = sab.make_exchange(dist) if world > 1 else None

    def resident_step():
        """device ms of one pass over the level: this rank's AE range + (N > 1) the device-side
        exchange that leaves every rank with the results of all AEs"""
        if world == 1:
            return h.sa_drv_bench_step(B, 0, 0, nae)
        gl.sa_gpu_ctx_timer(ctxp, 1)
        rc = gl.sa_gpu_local_spectral(lev, theta, a, b, 0)
        if rc != 0:
            raise RuntimeError(gl.sa_gpu_last_error().decode())
        exchange(lev.value, a, b, nae)
        return gl.sa_gpu_ctx_timer(ctxp, 0)

    # ---- device-resident steps ("value")
    for _ in range(args.warmup):
        resident_step()
    launches0 = h.sa_drv_bench_scalar(B, b"launches")
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ms_local = 0.0
    for _ in range(args.steps):
        ms_local += resident_step()
    barrier()
    sampler.stop_flag = True
    launches = int(sum_over_ranks(h.sa_drv_bench_scalar(B, b"launches") - launches0))
    ms_total = max_over_ranks(ms_local)
    value = nae * args.steps / (ms_total * 1e-3)

    # ---- end to end from host buffers (pinned): upload of the inputs the rank's AEs read,
    #      compute, read-back of their m / lambda / vectors
    e2e_mode = 1 if world == 1 else 2
    for _ in range(args.warmup):
        h.sa_drv_bench_step(B, e2e_mode, a, b)
    sampler2 = ClockSampler(local_rank)
    sampler2.start()
    barrier()
    ms_e2e = 0.0
    for _ in range(args.steps):
        ms_e2e += h.sa_drv_bench_step(B, e2e_mode, a, b)
    barrier()
    sampler2.stop_flag = True
    sampler.merge(sampler2)
    ms_e2e = max_over_ranks(ms_e2e)
    e2e_value = nae * args.steps / (ms_e2e * 1e-3)
    h2d_full = h.sa_drv_bench_scalar(B, b"h2d_bytes")
    h2d = h2d_full if world == 1 else sum_over_ranks(h.sa_drv_bench_scalar(B, b"h2d_bytes_last"))
    d2h = sum_over_ranks(h.sa_drv_bench_scalar(B, b"d2h_bytes"))

    # ---- N > 1: the replicated run of round 1 as an extra (every rank the whole level, no exchange)
    weak = None
    if world > 1:
        barrier()
        ms_w = 0.0
        for _ in range(2):
            ms_w += h.sa_drv_bench_step(B, 0, 0, nae)
        barrier()
        ms_w = max_over_ranks(ms_w)
        weak = {"value": world * nae * 2 / (ms_w * 1e-3), "unit": "AE/s", "what": "every rank processes all "
                "%d AEs of the level (N replicated problems, no exchange): the weak-scaling number of round 1" % nae}

    # ---- roofline of the dominant kernel (assemble + tridiagonalise), profiled steps
    h.sa_drv_gpu_profile(1, None, 0)
    prof_steps = 2
    for _ in range(prof_steps):
        h.sa_drv_bench_step(B, 0, a, b)
    buf = ctypes.create_string_buffer(8192)
    h.sa_drv_gpu_profile(0, buf, 8192)
    prof = {}
    for ln in buf.value.decode().splitlines():
        k, v = ln.split()
        prof[k] = float(v) / prof_steps
    flops = h.sa_drv_bench_scalar(B, b"flops") * share
    abytes = h.sa_drv_bench_scalar(B, b"bytes") * share
    fp64_peak = gl.sa_gpu_bench_fp64_peak(ctxp)
    kern_ms = prof.get("eig.assemble_tridiag", float("nan"))
    traffic = None  # DRAM bytes per step of the dominant kernel, from the committed ncu capture
    ncu_note = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tj.get("workload") == args.workload and world == 1:
            traffic = tj["eigen_stage"]["dram_bytes_per_step"]
            ncu_note = tj["eigen_stage"].get("ncu")
    except Exception:
        traffic = None
    achieved = flops / (kern_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "k_tridiag_reg<S> (register-resident Householder tridiagonalisation, one launch per size "
                  "class and upload piece) fed by k_at_packed (assembly + weighted-l1 scaling); classes side by "
                  "side on three streams, timed together",
        "bound": "fp64", "bound_detail": "FP64 FMA pipe; at n ~ 125 the ~n dependent Householder steps are latency / "
                                         "issue bound (no tensor instruction: BLAS-2 shaped work)",
        "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": (achieved / fp64_peak) if fp64_peak else None,
        "peak_source": "measured in this run: dependent-free DFMA loop (sa_gpu_bench_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
        "algorithmic_flops_per_step": flops, "algorithmic_bytes_per_step": abytes,
        "kernel_ms_per_step": kern_ms, "stage_ms": prof, "traffic": traffic,
        "traffic_source": "profiles/r02_traffic.json (ncu dram bytes of the k_at_packed + k_tridiag_reg launches of one step)" if traffic else None,
        "ncu": ncu_note,  # pipe utilisation of the committed capture (not measured in this run)
    }

    line = {
        "metric": "agglomerate eigensolves/sec", "value": value, "unit": "AE/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": bench_config(args, pr, nae),
        "sharding": {"n_AE_this_rank": int(b - a), "exchange": None if world == 1 else
                     "NCCL on the level's device arrays: 1 all-reduce of the counts + one broadcast per rank "
                     "and array (eigenvalues, eigenvectors, D) over NVLink, inside the timed region"},
        "host_inputs_s": t_inputs,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "AE/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "pinned": bool(h.sa_drv_bench_scalar(B, b"pinned")),
                "what": "all inputs of the level" if world == 1 else
                "every rank uploads the inputs its own AEs read (+ the index tables) and reads its own results back"},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if weak:
        line["weak_replicated"] = weak
    h.sa_drv_bench_destroy(B)

    # ---- the target metric: full hierarchy setup + PCG solve, beside the CPU path.
    # With several ranks the AE loop of every level is sharded over the GPUs (device-side
    # exchange) and the solve is row-partitioned with NCCL halo exchange.
    peaks, peak_src = measured_peaks()
    its = None
    H = None
    if not args.no_hierarchy:
        if world > 1:
            sab.enable_sharding(dist)
            barrier()
        # host input side: page-lock the problem's arrays so the finest level goes up through the
        # pipelined path (the eigen stage starts while the operator / element blocks are in flight)
        h.sa_drv_problem_pin.restype = ctypes.c_double
        h.sa_drv_problem_pin.argtypes = [ctypes.c_void_p, ctypes.c_int]
        pin_s = h.sa_drv_problem_pin(pr.handle, local_rank)
        barrier()
        t0 = time.time()
        H = sab.ml_build(pr, p, local_rank)
        barrier()
        setup_s = max_over_ranks(time.time() - t0)
        hier = {"levels": w["levels"], "setup_s": setup_s, "setup_sharded_over_gpus": world,
                "host_pin_s": pin_s}
        if world > 1:
            # bytes this rank moved in the owner-sharded stages (MIS blocks to their owners, gathered
            # MIS bases, rows of the row-partitioned products received): the collectives of the setup
            hier["setup_exchange_bytes_rank0"] = sab.sharding_stats()
        from saamge_b200.dist_solve import DistSolver

        # PCG of the library's row-partitioned solver (saamge_b200/csrc/dist.cu) on all N ranks;
        # N = 1: the same loop without any exchange
        S = DistSolver(H, dist if world > 1 else None)
        bvec = pr.get("b")
        S.pcg(bvec, maxiter=2)  # warm-up (NCCL channels)
        barrier()
        t0 = time.time()
        _x, its_d, _brr = S.pcg(bvec)
        barrier()
        hier["pcg_s"] = max_over_ranks(time.time() - t0)
        hier["pcg_device_s"] = max_over_ranks(S.solve_seconds)
        hier["pcg_gpus"] = world
        hier["pcg_iterations"] = int(its_d)
        its = int(its_d)
        bpi = S.bytes_per_iteration()
        hbm = peaks.get("hbm_gbs") if peaks else None
        ach = bpi * abs(its) / hier["pcg_device_s"] / 1e9 if its else None
        hier["roofline_pcg"] = {
            "bound": "hbm", "achieved": ach, "peak": (hbm * world) if hbm else None, "unit": "GB/s",
            "frac": (ach / (hbm * world)) if (ach and hbm) else None, "traffic": None,
            "ms_per_iteration": 1e3 * hier["pcg_device_s"] / max(1, abs(its)),
            "algorithmic_bytes_per_iteration": bpi,
            "what": "per level (2 deg + 1) A-SpMVs + fused smoother vectors + 2 P-SpMVs, one A-SpMV and 56 B/row "
                    "of vector traffic on the finest level (SURVEY section 8d); peak = N x measured copy bandwidth",
            "levels_rows_nnzA_nnzP": S.level_info()}
        halo = S.stats()
        hier["halo"] = {"exchanges": int(halo[2]), "doubles_sent_by_rank0": int(halo[3])} if world > 1 else None
        if world == 1:
            # a second, profiled build: per-kernel times of the large-AE eigensolver (levels >= 1)
            try:
                h.sa_drv_gpu_profile(1, None, 0)
                H2 = sab.ml_build(pr, p, local_rank)
                pbuf = ctypes.create_string_buffer(8192)
                h.sa_drv_gpu_profile(0, pbuf, 8192)
                pm = {ln.split()[0]: float(ln.split()[1]) for ln in pbuf.value.decode().splitlines() if ln.strip()}
                n3, cnt_large = 0.0, 0
                for l in range(1, w["levels"] - 1):
                    nn = np.diff(H2.get("AE_to_dof.I", l)).astype(np.float64)
                    nn = nn[nn > 208]
                    n3 += float((nn ** 3).sum())
                    cnt_large += int(len(nn))
                H2.close()
                chol_traffic = None
                try:
                    tj2 = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
                    if tj2.get("workload") == args.workload:
                        chol_traffic = tj2["k_cs_chol"]["dram_bytes_per_launch"]
                except Exception:
                    chol_traffic = None
                chol_ms = pm.get("eig.cs_chol")
                if chol_ms and n3 > 0:
                    eig_ms = sum(v for k, v in pm.items() if k in ("eig.cs_chol", "eig.cs_iterate", "eig.large_assemble"))
                    hier["large_AE_eigensolver"] = {
                        "what": "levels >= 1 (n > 208): Cholesky of A^ - sigma I (DMMA) + shift-invert subspace iteration "
                                "(cholsi.cu) instead of a tridiagonalisation",
                        "matrices": cnt_large, "sum_n3": n3,
                        "ms": {k: v for k, v in pm.items() if k.startswith("eig.cs_") or k == "eig.large_assemble"},
                        "roofline": {"kernel": "k_cs_chol", "bound": "tensor",
                                     "bound_detail": "FP64 tensor pipe (DMMA m8n8k4): n^3/3 flops per matrix executed",
                                     "achieved": n3 / 3.0 / (chol_ms * 1e-3) / 1e12, "peak": fp64_peak,
                                     "unit": "TFLOP/s", "frac": n3 / 3.0 / (chol_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                     "peak_source": "FP64 FMA peak measured in this run (DMMA measured at 1.07x of it, tools/dmma_bench.cu)",
                                     "traffic": chol_traffic,
                                     "traffic_source": "profiles/r02_traffic.json: ncu DRAM bytes of the level-1 launch (630 of the %d matrices)" % cnt_large if chol_traffic else None},
                        "algorithmic_tflops_4_3_n3": (4.0 / 3.0) * n3 / (eig_ms * 1e-3) / 1e12,
                    }
            except Exception as ex:
                hier["large_AE_eigensolver"] = {"error": repr(ex)}
        if rank == 0:
            t0 = time.time()
            its1 = sab.ml_pcg(H, 1000, 1e-12, 0.0)
            t1 = time.time() - t0
            hier.update({"pcg_s_host_api_one_gpu": t1, "pcg_iterations_one_gpu": its1})
            hier.update({"final_residual": H.scalar("pcg.final_res_norm"),
                         "stage_s": {k: round(v, 4) for k, v in H.times().items()},
                         "dofs": [int(H.scalar("ND", l)) for l in range(w["levels"] - 1)]})
            line["hierarchy"] = hier
    if world > 1:
        dist.destroy_process_group()
        if rank != 0:
            return  # rank 0 alone measures the CPU baselines (no idle ranks spinning beside it)
    if H is not None and hasattr(h, "sa_drv_ml_spmv_bench"):
        h.sa_drv_ml_spmv_bench.restype = ctypes.c_double
        h.sa_drv_ml_spmv_bench.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        nnz = pr.scalar("nnz")
        nd = pr.scalar("ND")
        ms = h.sa_drv_ml_spmv_bench(H.handle, 0, 50)
        ms2 = h.sa_drv_ml_spmv_bench(H.handle, 1, 50)
        b_spmv = 12.0 * nnz + 20.0 * nd
        b_sm = b_spmv + 24.0 * nd
        pk = peaks["hbm_gbs"]
        spmv_traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("workload") == args.workload:
                spmv_traffic = tj["k_spmv"]["dram_bytes_per_launch"]
        except Exception:
            spmv_traffic = None
        line["roofline_spmv"] = {"bound": "hbm", "achieved": b_spmv / (ms * 1e-3) / 1e9, "peak": pk,
                                 "unit": "GB/s", "frac": b_spmv / (ms * 1e-3) / 1e9 / pk,
                                 "traffic": spmv_traffic,
                                 "ms": ms, "peak_source": peak_src, "algorithmic_bytes": b_spmv}
        line["roofline_smoother"] = {"bound": "hbm", "achieved": b_sm / (ms2 * 1e-3) / 1e9, "peak": pk,
                                     "unit": "GB/s", "frac": b_sm / (ms2 * 1e-3) / 1e9 / pk,
                                     "traffic": None, "ms": ms2, "algorithmic_bytes": b_sm}
    if not args.no_cpu:
        import oracle_util as ou

        procs = os.cpu_count() or 1
        try:
            ns, t, cores = cpu_sample(ou, pr, p, nae, args.cpu_seconds, procs)
            if ns == nae and t < 0.5 * args.cpu_seconds:
                # the whole level is a short sample on this box: repeat it to ~cpu_seconds of work
                reps = min(8, int(args.cpu_seconds / max(t, 1e-3)))
                for _ in range(reps):
                    ns2, t2, _c = cpu_sample(ou, pr, p, nae, args.cpu_seconds, procs)
                    ns += ns2
                    t += t2
            line["cpu_baseline"] = {"value": ns / t, "unit": "AE/s", "cores": cores, "kind": "port",
                                    "sample": "%d AEs (passes over the first AEs of the %d of the level; assemble + D + LAPACK dsygvx), %.1f s, one single-threaded process per core on contiguous AE slices (mpirun -n P analogue)" % (ns, nae, t)}
            # what ONE reference MPI rank does (SURVEY 8d asks for both numbers)
            ns1, t1, c1 = cpu_sample(ou, pr, p, nae, min(5.0, args.cpu_seconds), 1)
            line["cpu_baseline_1core"] = {"value": ns1 / t1, "unit": "AE/s", "cores": c1, "kind": "port",
                                          "sample": "first %d AEs, %.1f s" % (ns1, t1)}
            if H is not None and its:
                cb = cpu_hierarchy_baseline(ou, sab, H, pr, p, nae, ns / t, procs, its)
                hier = line["hierarchy"]
                hier.update(cb)
                gpu_total = hier["setup_s"] + hier["pcg_s"]
                cpu_total = cb["cpu_setup_s"] + cb["cpu_pcg_s"]
                hier["target_metric"] = {
                    "what": "north star: (setup local spectral stage + PCG solve) on %d B200 vs the CPU path on "
                            "the box's %d host cores; the GPU figure is the WHOLE setup (all stages, host topology "
                            "included), the CPU figure covers the local spectral stages and the solve only" % (world, procs),
                    "gpu_setup_plus_pcg_s": gpu_total, "cpu_setup_plus_pcg_s": cpu_total,
                    "ratio": cpu_total / gpu_total}
        except Exception as ex:  # the GPU numbers above must not be lost to a CPU-side failure
            line.setdefault("cpu_baseline", {"value": None, "unit": "AE/s", "cores": 0, "kind": "port",
                                             "sample": "failed: %r (see bench.py --impl reference)" % (ex,)})
            line["cpu_baseline_error"] = repr(ex)
    if H is not None:
        H.close()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
