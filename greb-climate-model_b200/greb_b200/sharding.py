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


def reduce_field_moments(ens, n_local: int, dst: int | None = 0, group=None):
    """Ensemble mean and variance FIELDS of the last completed year's monthly means over all ranks.

    Each rank's library reduces its own members on the device (greb_b200_ensemble_moments_device: float64 sum and
    sum of squares per element of [12][5][48][96]); the two vectors and the member count are then summed
    across ranks with ONE NCCL reduce to rank `dst` (dst=None: all-reduce) straight from the library's device
    buffers — 4.4 MB per rank instead of 1.1 MB per member.  Returns (mean, variance, n_total) as float64
    arrays on the destination rank(s), (None, None, n) elsewhere."""
    import torch
    import torch.distributed as dist
    ps, pq, n = ens.ensemble_moments_device()

    class _View:
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    dev = torch.device("cuda", torch.cuda.current_device())
    buf = torch.empty(2 * n + 1, dtype=torch.float64, device=dev)
    buf[:n] = torch.as_tensor(_View(ps), device=dev)
    buf[n:2 * n] = torch.as_tensor(_View(pq), device=dev)
    buf[2 * n] = float(n_local)
    return finish_field_moments(buf, n, dst, group)


def finish_field_moments(buf, n: int, dst: int | None = 0, group=None):
    """the collective + the mean/variance arithmetic of reduce_field_moments on a [sum | sumsq | count]
    float64 tensor (any device; gloo on CPU in the tests)"""
    import torch.distributed as dist
    mine = True
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dst is None:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
            mine = dist.get_rank(group) == dst
    cnt = float(buf[2 * n])
    if not mine:
        return None, None, cnt
    mean = buf[:n] / cnt
    var = (buf[n:2 * n] / cnt - mean * mean).clamp_(min=0.0)
    return mean.cpu().numpy().reshape(12, 5, 48, 96), var.cpu().numpy().reshape(12, 5, 48, 96), cnt
