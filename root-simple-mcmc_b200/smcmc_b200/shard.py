"""Chain sharding across the GPUs of one node.

Chains are independent, so the ensemble is cut into contiguous blocks, one
per rank, and there is NO collective on the data path: every rank steps its
own engine.  The only cross-rank traffic is (a) the barrier / max-time of the
benchmark and (b) optional pooled diagnostics, an all-reduce of the sufficient
statistics (count, sum x, sum x x^T) -- the formulas of MakeCovariance.C:63-89.

A chain's random draws are addressed by its GLOBAL index (chain_offset + local
index, include/smcmc_rng.h), so the chains do not depend on how many GPUs the
ensemble is spread over.
"""
import numpy as np


def chain_shard(total_chains, world_size, rank):
    """(offset, count) of the contiguous block of chains owned by `rank`;
    the first total % world ranks get one extra chain."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(total_chains, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def pooled_moments(points, dist=None, device=None):
    """Ensemble mean and covariance of `points` (local_chains x dim, numpy)
    pooled over every rank: all-reduce of (count, sum x, sum x x^T)."""
    import torch
    x = torch.as_tensor(np.ascontiguousarray(points), dtype=torch.float64)
    n, d = x.shape
    stats = torch.cat([torch.tensor([float(n)], dtype=torch.float64), x.sum(0), (x.T @ x).reshape(-1)])
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            stats = stats.to(device)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        stats = stats.cpu()
    count = float(stats[0])
    mean = stats[1:1 + d] / count
    second = stats[1 + d:].reshape(d, d) / count
    cov = second - torch.outer(mean, mean)
    return count, mean.numpy(), cov.numpy()


def gather_chains(local, dist=None):
    """Concatenate per-rank arrays (chains first axis) in rank order on every
    rank (uneven shards allowed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.asarray(local))
    return np.concatenate(parts, axis=0)
