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
    """Every piece a rank plans to receive is planned as a send by its owner and lies in the
    owner's range; together the pieces cover the needed range outside the own rows."""
    from saamge_b200.dist_solve import _ranges, halo_plan

    rng = np.random.default_rng(1)
    for world in (2, 3, 8):
        n = 1000
        part = _ranges(n, world)
        need = []
        for q in range(world):
            lo = max(0, part[q] - int(rng.integers(0, 300)))
            hi = min(n, part[q + 1] + int(rng.integers(0, 300)))
            need.append((lo, hi))
        plans = [halo_plan(need, part, q) for q in range(world)]
        for me in range(world):
            sends, recvs = plans[me]
            covered = np.zeros(n, dtype=bool)
            covered[part[me]:part[me + 1]] = True
            for q, lo, hi in recvs:
                assert part[q] <= lo < hi <= part[q + 1]
                assert (me, lo, hi) in plans[q][0]
                covered[lo:hi] = True
            assert covered[need[me][0]:need[me][1]].all()
            for q, lo, hi in sends:
                assert (me, lo, hi) in plans[q][1]
