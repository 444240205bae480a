"""Run under torchrun on N GPUs: the AE-sharded local spectral stage + all-gather gives
every rank the same (m, lambda, eigenspaces) as the unsharded stage, and the tentative
prolongator built from it matches.  Prints one line per rank; rank 0 prints PASS/FAIL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saamge_b200 as sab  # noqa: E402
from saamge_b200 import cabi  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    lrank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    p = sab.default_params(num_levels=2, first_elems_per_agg=52, partition_kind=0)
    pr = sab.Problem(3, n, coef_kind=1)
    pr.partition(p)
    ctx = cabi.Context(lrank)
    ref = cabi.Level(ctx, pr)
    ref.local_spectral(0.003)
    m0, ev0, Z0, D0 = ref.spectral()
    ncd0, NDc0 = ref.tentative_P(1)
    lev = cabi.Level(ctx, pr)
    a, b = lev.local_spectral_sharded(0.003, dist)
    m1, ev1, Z1, D1 = lev.spectral()
    ncd1, NDc1 = lev.tentative_P(1)
    ok = (np.array_equal(m0, m1) and np.allclose(ev0, ev1, rtol=0, atol=1e-12) and np.allclose(D0, D1, rtol=1e-13)
          and np.array_equal(ncd0, ncd1) and NDc0 == NDc1)
    # eigenvectors: same up to sign per column (deterministic kernels -> usually identical)
    ok = ok and np.allclose(np.abs(Z0), np.abs(Z1), atol=1e-9)
    P0, P1 = ref.csr(1), lev.csr(1)
    ok = ok and abs(abs(P0) - abs(P1)).max() < 1e-8
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print("rank %d range [%d,%d) ok=%s" % (rank, a, b, ok), flush=True)
    if rank == 0:
        print("MGPU_SHARDED", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
