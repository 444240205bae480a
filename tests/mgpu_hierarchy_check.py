"""Run under torchrun on N GPUs: full hierarchy build sharded over the ranks -- mode "replicate"
(AE loop sharded, results all-gathered) and mode "owner" (MIS-owner tentative P after one
all-to-all-v, coarse element matrices by AE range, row-partitioned smoothing / RAP: the
sa_gpu_dist_* stages) -- gives the same prolongators, coarse operators, PCG iteration count and
residual history as the unsharded build (rank 0 prints PASS/FAIL and timings)."""
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
    sab.ml_download(H0)
    ref = [(H0.csr("interp", l), H0.csr("Ac", l)) for l in range(levels - 1)]
    H0.close()
    ok = True
    for mode in ("replicate", "owner"):
        sab.enable_sharding(dist, mode=mode)
        dist.barrier()
        t0 = time.time()
        H1 = sab.ml_build(pr, p, lrank)
        torch.cuda.synchronize()
        dist.barrier()
        t_shard = time.time() - t0
        it1 = sab.ml_pcg(H1)
        brr1 = H1.get("pcg.brr")
        k = min(len(brr0), len(brr1), 4)
        good = (it0 == it1) and np.allclose(brr0[:k], brr1[:k], rtol=1e-6)
        sab.ml_download(H1)
        worst = 0.0
        for l in range(levels - 1):
            P1, A1 = H1.csr("interp", l), H1.csr("Ac", l)
            P0, A0 = ref[l]
            if P1.shape != P0.shape or A1.shape != A0.shape:
                good = False
                continue
            eP = abs(P1 - P0).max() / max(abs(P0).max(), 1e-300)
            eA = abs(A1 - A0).max() / max(abs(A0).max(), 1e-300)
            worst = max(worst, eP, eA)
        good = good and worst <= 1e-10
        ok = ok and good
        print("rank %d mode %s iters %d/%d max rel. difference of P / Ac %.1e setup single %.2fs sharded %.2fs "
              "stages %s moved %s" % (rank, mode, it0, it1, worst, t_single, t_shard,
                                    {k: round(v, 3) for k, v in H1.times().items()
                                     if k.split(".")[-1] in ("local_spectral", "tentative", "rap", "coarse_elmats")},
                                    sab.sharding_stats() if mode == "owner" else None), flush=True)
        H1.close()
        sab.disable_sharding()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_HIERARCHY", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
