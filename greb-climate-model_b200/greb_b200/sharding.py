"""Ensemble sharding across GPUs.

Members never interact (the reference runs one process per member, src/greb.f90:1030-1068), so the
path shards by members with NO data-path collective: rank r owns a contiguous block of members and
its own replica of the shared forcing.  The only exchange is the all-reduce of the ensemble
statistics of the per-member annual global-mean Tsurf — a few doubles per simulated year.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the members of `rank`: contiguous blocks, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("shard_range: bad arguments")
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def owner_of(member: int, n_total: int, world: int) -> int:
    """rank that owns global member index `member` under shard_range"""
    base, extra = divmod(n_total, world)
    if member < (base + 1) * extra:
        return member // (base + 1)
    return extra + (member - (base + 1) * extra) // max(base, 1)


def local_moments(values) -> "np.ndarray":
    """[count, sum, sum of squares] in float64 of a rank's per-member diagnostic values
    (numpy array or torch tensor on any device)."""
    try:
        import torch
        if isinstance(values, torch.Tensor):
            v = values.double().flatten()
            return torch.stack([torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device), v.sum(),
                                (v * v).sum()])
    except ImportError:  # pragma: no cover
        pass
    v = np.asarray(values, dtype=np.float64).ravel()
    return np.array([v.size, v.sum(), (v * v).sum()], dtype=np.float64)


def allreduce_moments(moments, group=None):
    """Sum the moments over all ranks (torch.distributed, NCCL on GPUs / gloo on CPU).  Without an
    initialised process group the input is returned unchanged (single-rank job)."""
    import torch
    import torch.distributed as dist
    t = moments if isinstance(moments, torch.Tensor) else torch.as_tensor(np.asarray(moments, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def ensemble_mean_std(moments) -> Tuple[float, float]:
    """ensemble mean and (population) standard deviation from reduced [count, sum, sumsq]"""
    n, s, ss = (float(x) for x in moments)
    if n <= 0:
        return float("nan"), float("nan")
    mean = s / n
    var = max(ss / n - mean * mean, 0.0)
    return mean, var ** 0.5
