"""Sharding of the local spectral stage over ranks (one process per GPU).

AEs are independent (the reference distributes them over MPI ranks the same way,
amg/src/interp.cpp:387), so every rank computes a contiguous AE range balanced on the
eigensolve cost ~ n^3 and the per-AE results (m, lambda, vectors, D) are exchanged with
one all-gather so that every rank holds the complete set for the tentative-P stage.
torch.distributed is only plumbing here (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_ranges(ae_sizes, world):
    """Contiguous [begin, end) per rank, balanced on sum n^3; every AE in exactly one."""
    n = np.asarray(ae_sizes, dtype=np.float64)
    cost = np.cumsum(n ** 3)
    total = cost[-1] if len(cost) else 0.0
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cost, target, side="left")) + 1
        b = max(bounds[-1], min(b, len(n)))
        bounds.append(b)
    bounds.append(len(n))
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def pack_range(ae_sizes, m, evals, evects, D, begin, end):
    """Slices of the flat per-AE arrays that belong to AEs [begin, end)."""
    n = np.asarray(ae_sizes, dtype=np.int64)
    m = np.asarray(m, dtype=np.int64)
    eo = np.concatenate([[0], np.cumsum(m)])
    zo = np.concatenate([[0], np.cumsum(m * n)])
    do = np.concatenate([[0], np.cumsum(n)])
    return (np.asarray(m[begin:end], dtype=np.int32), evals[eo[begin]:eo[end]],
            evects[zo[begin]:zo[end]], D[do[begin]:do[end]])


def allgather_spectral(local, dist=None, group=None):
    """local = (m, evals, evects, D) of this rank's AE range; returns the concatenation
    over ranks in rank order.  Works with any torch.distributed backend."""
    import torch

    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    out = []
    for a, dt in zip(local, (torch.int32, torch.float64, torch.float64, torch.float64)):
        t = torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
        cnt = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt, group=group)
        sizes = [int(c.item()) for c in cnts]
        mx = max(sizes + [1])
        pad = torch.zeros(mx, dtype=dt, device=dev)
        pad[: t.numel()] = t
        bufs = [torch.zeros(mx, dtype=dt, device=dev) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out.append(torch.cat([b[:s] for b, s in zip(bufs, sizes)]).cpu().numpy())
    return tuple(out)
