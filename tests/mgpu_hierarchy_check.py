"""Run under torchrun on N GPUs: full hierarchy build with the AE loop of every level
sharded over the ranks gives the same PCG iteration count / residual history as the
unsharded build (rank 0 prints PASS/FAIL and timings)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saamge_b200 as sab  # noqa: E402


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
    t0 = time.time()
    H0 = sab.ml_build(pr, p, lrank)
    t_single = time.time() - t0
    it0 = sab.ml_pcg(H0)
    brr0 = H0.get("pcg.brr")
    H0.close()
    sab.enable_sharding(dist)
    dist.barrier()
    t0 = time.time()
    H1 = sab.ml_build(pr, p, lrank)
    torch.cuda.synchronize()
    dist.barrier()
    t_shard = time.time() - t0
    it1 = sab.ml_pcg(H1)
    brr1 = H1.get("pcg.brr")
    k = min(len(brr0), len(brr1), 4)
    ok = (it0 == it1) and np.allclose(brr0[:k], brr1[:k], rtol=1e-6)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print("rank %d iters %d/%d setup single %.2fs sharded %.2fs stages %s" % (
        rank, it0, it1, t_single, t_shard,
        {k: round(v, 2) for k, v in H1.times().items() if "local_spectral" in k}), flush=True)
    if rank == 0:
        print("MGPU_HIERARCHY", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
