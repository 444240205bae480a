"""Run under torchrun on N GPUs: row-partitioned V-cycle PCG (halo exchange over NCCL)
reproduces the single-GPU PCG: same iteration count, same (B r, r) history, same solution."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saamge_b200 as sab  # noqa: E402
from saamge_b200.dist_solve import DistSolver  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    lrank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    tile = 32 if n % 32 == 0 else n
    p = sab.default_params(num_levels=levels, first_elems_per_agg=52, elems_per_agg=64,
                           partition_kind=2, block=(tile, tile, tile))
    pr = sab.Problem(3, n, coef_kind=1)
    pr.partition(p)
    sab.enable_sharding(dist)
    H = sab.ml_build(pr, p, lrank)
    t0 = time.time()
    it0 = sab.ml_pcg(H)
    t_single = time.time() - t0
    brr0 = H.get("pcg.brr")
    x0 = H.get("pcg.x")
    S = DistSolver(H, dist)
    b = pr.get("b")
    S.pcg(b, maxiter=2)  # warm-up (NCCL channels)
    dist.barrier()
    t0 = time.time()
    x, it1, brr1 = S.pcg(b)
    torch.cuda.synchronize()
    dist.barrier()
    t_dist = time.time() - t0
    xs = S.gather_solution(x)
    k = min(len(brr0), len(brr1))
    ok = (it0 == it1) and np.allclose(brr0[:k], np.array(brr1[:k]), rtol=1e-6) and \
        np.linalg.norm(xs - x0) <= 1e-8 * np.linalg.norm(x0)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print("rank %d iters single %d dist %d  time single %.3fs dist %.3fs halo exchanges %d  |dx|/|x| %.2e" % (
        rank, it0, it1, t_single, t_dist, S.halo_calls, np.linalg.norm(xs - x0) / np.linalg.norm(x0)), flush=True)
    if rank == 0:
        print("MGPU_SOLVE", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
