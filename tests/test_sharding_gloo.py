"""world_size-2 gloo test of the multi-GPU host logic: AE range sharding + all-gather of
the per-AE spectral results reproduces the single-process arrays (CPU, oracle data)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_util as ou
import saamge_b200 as sab
from saamge_b200 import sharding


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    sizes = rng.integers(80, 200, size=1000)
    for world in (1, 2, 3, 8):
        r = sharding.shard_ranges(sizes, world)
        assert r[0][0] == 0 and r[-1][1] == len(sizes)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        cost = [float(np.sum(sizes[a:b].astype(float) ** 3)) for a, b in r]
        assert max(cost) <= 1.05 * (sum(cost) / world) + float(sizes.max()) ** 3
    assert sharding.shard_ranges([], 2) == [(0, 0), (0, 0)]
    assert sharding.shard_ranges([5], 4)[-1][1] == 1


def _worker(rank, world, port, ref, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes, m, ev, Z, D = ref
    a, b = sharding.shard_ranges(sizes, world)[rank]
    local = sharding.pack_range(sizes, m, ev, Z, D, a, b)
    got = sharding.allgather_spectral(local, dist)
    ok = all(np.array_equal(g, r) for g, r in zip(got, (m, ev, Z, D)))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_allgather_spectral_world2():
    p = sab.default_params(num_levels=2, first_elems_per_agg=27, partition_kind=1, block=(3, 3, 3))
    pr = sab.Problem(3, 6, coef_kind=1)
    pr.partition(p)
    H = ou.orc_build(pr, p)
    I = H.get("AE_to_dof.I", 0)
    ref = (I[1:] - I[:-1], H.get("ae_m", 0), H.get("evals", 0), H.get("evects", 0), H.get("ae_D", 0))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ref, q)) for r in range(2)]
    for pp in procs:
        pp.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for pp in procs:
        pp.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_halo_plan_is_consistent():
    """sa_gpu_halo_plan (host side of the row-partitioned solve, saamge_b200/csrc/dist.cu): what a
    rank plans to receive from q is exactly what q plans to send to it, lies in q's column range,
    and together with the own range covers every column the rank's rows reference."""
    import scipy.sparse as sp

    from saamge_b200.dist_solve import _ranges, halo_plan_lists

    rng = np.random.default_rng(1)
    for world in (2, 3, 8):
        for rows, cols in ((1000, 1000), (1000, 317), (40, 1000)):
            M = sp.random(rows, cols, density=0.01, random_state=int(rng.integers(1 << 30)), format="csr")
            M = (M + sp.eye(rows, cols, k=0) + sp.eye(rows, cols, k=1)).tocsr()
            M.sort_indices()
            rp, cp = _ranges(rows, world), _ranges(cols, world)
            plans = [halo_plan_lists(M.indptr, M.indices, world, q, rp, cp) for q in range(world)]
            for me in range(world):
                send, recv = plans[me]
                ref = np.unique(M.indices[M.indptr[rp[me]]:M.indptr[rp[me + 1]]])
                outside = ref[(ref < cp[me]) | (ref >= cp[me + 1])]
                got = np.concatenate(recv) if world else np.zeros(0, dtype=np.int32)
                assert np.array_equal(np.sort(got), outside)
                assert len(recv[me]) == 0 and len(send[me]) == 0
                for q in range(world):
                    assert np.all((recv[q] >= cp[q]) & (recv[q] < cp[q + 1]))
                    assert np.array_equal(recv[q], plans[q][0][me])  # q sends what I expect
                    assert np.all(np.diff(recv[q]) > 0)


def test_mis_exchange_plan_is_consistent():
    """sa_gpu_mis_exchange_plan (host side of sa_gpu_dist_tentative_P, the reduce-to-owner exchange
    of amg/src/contrib.cpp:492-549): every MIS has one owner -- the rank of the lowest-numbered AE
    containing it (amg/src/aggregates.cpp:583-593) --, what rank r plans to send to q is what q
    plans to receive from r, nothing is sent to oneself, and the volume is exactly the
    MIS-restricted blocks (s x m_AE doubles) of the pairs whose AE is not on the owner."""
    import ctypes

    g = sab.gpu_lib()
    ip = ctypes.POINTER(ctypes.c_int)
    lp = ctypes.POINTER(ctypes.c_int64)
    p = sab.default_params(num_levels=2, first_elems_per_agg=27, partition_kind=0)
    pr = sab.Problem(3, 12, coef_kind=1)
    nparts = pr.partition(p)
    MI, MJ = pr.get("mis_to_AE.I"), pr.get("mis_to_AE.J")
    DI = pr.get("mis_to_dof.I")
    nmis = len(MI) - 1
    sizes = np.diff(pr.get("AE_to_dof.I"))
    rng = np.random.default_rng(3)
    ae_m = np.ascontiguousarray(rng.integers(0, 4, nparts), dtype=np.int32)
    MI, MJ, DI = (np.ascontiguousarray(a, dtype=np.int32) for a in (MI, MJ, DI))
    for world in (2, 3, 8):
        part = np.ascontiguousarray([r[0] for r in sharding.shard_ranges(sizes, world)] + [nparts], dtype=np.int32)
        rank_of = lambda ae: int(np.searchsorted(part, ae, side="right") - 1)
        owners, S, R = [], [], []
        for me in range(world):
            owner = np.zeros(nmis, dtype=np.int32)
            sd, rd = np.zeros(world, dtype=np.int64), np.zeros(world, dtype=np.int64)
            rc = g.sa_gpu_mis_exchange_plan(nmis, MI.ctypes.data_as(ip), MJ.ctypes.data_as(ip), DI.ctypes.data_as(ip),
                                            nparts, ae_m.ctypes.data_as(ip), world, me, part.ctypes.data_as(ip),
                                            owner.ctypes.data_as(ip), sd.ctypes.data_as(lp), rd.ctypes.data_as(lp))
            assert rc == 0, g.sa_gpu_last_error()
            owners.append(owner)
            S.append(sd)
            R.append(rd)
        for me in range(1, world):
            assert np.array_equal(owners[me], owners[0])
        expect = np.zeros((world, world), dtype=np.int64)
        for mis in range(nmis):
            aes = MJ[MI[mis]:MI[mis + 1]]
            own = rank_of(aes.min())
            assert owners[0][mis] == own
            s = DI[mis + 1] - DI[mis]
            if s == 1:
                continue
            for ae in aes:
                if rank_of(ae) != own:
                    expect[rank_of(ae), own] += s * ae_m[ae]
        for r in range(world):
            assert S[r][r] == 0 and R[r][r] == 0
            for q in range(world):
                assert S[r][q] == R[q][r] == expect[r, q], (world, r, q)
        assert expect.sum() > 0
    pr.close()
