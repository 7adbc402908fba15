"""Multi-GPU path without GPUs: member -> rank partitioning and the all-reduce of the ensemble
diagnostics with world_size 2 over gloo on the CPU."""
import os
import socket

import numpy as np
import pytest

from greb_b200 import sharding


def test_shard_ranges_partition_the_ensemble():
    for n_total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(n_total, world, r)
                assert 0 <= a <= b <= n_total
                seen.extend(range(a, b))
                for m in (a, b - 1):
                    if a < b:
                        assert sharding.owner_of(m, n_total, world) == r
            assert seen == list(range(n_total))
            sizes = [sharding.shard_range(n_total, world, r) for r in range(world)]
            assert max(b - a for a, b in sizes) - min(b - a for a, b in sizes) <= 1


def test_config4_shape():
    # BASELINE.json configs[3]: 65,536 members on 8 GPUs = 8,192 per GPU
    assert [sharding.shard_range(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 65536)]


def _worker(rank, world, port, n_total, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        gmean_all = rng.normal(15.0, 1.5, n_total).astype(np.float32)   # per-member annual global means
        a, b = sharding.shard_range(n_total, world, rank)
        local = torch.from_numpy(gmean_all[a:b])
        red = sharding.allreduce_moments(sharding.local_moments(local))
        ret[rank] = [float(x) for x in red]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allreduce_of_ensemble_moments_world2():
    import torch.multiprocessing as mp
    n_total, world = 1000, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    g = rng.normal(15.0, 1.5, n_total).astype(np.float32).astype(np.float64)
    want = np.array([n_total, g.sum(), (g * g).sum()])
    for r in range(world):
        assert np.allclose(ret[r], want, rtol=1e-12), (r, ret[r], want)      # every rank holds the full sums
    mean, std = sharding.ensemble_mean_std(ret[0])
    assert abs(mean - g.mean()) < 1e-9 and abs(std - g.std()) < 1e-9


def test_single_rank_is_identity():
    m = sharding.local_moments(np.array([1.0, 2.0, 3.0], dtype=np.float32))
    out = sharding.allreduce_moments(m)
    assert [float(x) for x in out] == [3.0, 6.0, 14.0]


def _field_worker(rank, world, port, n_total, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        n = 12 * 5 * 48 * 96
        fields = rng.normal(280.0, 3.0, (n_total, 200))                     # 200 elements stand in for n
        a, b = sharding.shard_range(n_total, world, rank)
        loc = fields[a:b]
        buf = torch.zeros(2 * n + 1, dtype=torch.float64)
        buf[:200] = torch.from_numpy(loc.sum(0))
        buf[n:n + 200] = torch.from_numpy((loc * loc).sum(0))
        buf[2 * n] = b - a
        mean, var, cnt = sharding.finish_field_moments(buf, n, dst=0)
        ret[rank] = None if mean is None else (mean.ravel()[:200].copy(), var.ravel()[:200].copy(), cnt)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reduce_of_ensemble_field_moments_world2():
    """the collective half of sharding.reduce_field_moments (ncclReduce on the GPUs, gloo here): rank 0 ends up
    with the ensemble mean and variance fields of ALL members, rank 1 with nothing"""
    import torch.multiprocessing as mp
    n_total, world = 37, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_field_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    fields = rng.normal(280.0, 3.0, (n_total, 200))
    assert ret[1] is None
    mean, var, cnt = ret[0]
    assert cnt == n_total
    assert np.allclose(mean, fields.mean(0), rtol=1e-13) and np.allclose(var, fields.var(0), rtol=1e-8, atol=1e-9)
