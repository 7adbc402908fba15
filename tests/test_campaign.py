"""Sharded, batched ensemble driver (greb_b200/campaign.py; BASELINE.json configs[3]).

CPU part: batch planning and the rank/batch bookkeeping + final all-reduce under gloo world_size 2,
with a stand-in for the device handle (the product default is the CUDA library; the stand-in only
records which members it was given).  GPU part: a batched run equals one big handle bit for bit.
"""
import os
import socket

import numpy as np
import pytest

from greb_b200 import campaign, sharding


def test_plan_batches_covers_the_shard():
    for n in (0, 1, 5, 2048, 8192, 8193):
        for batch in (1, 7, 2048, 4000):
            bs = campaign.plan_batches(n, batch)
            assert [m for a, b in bs for m in range(a, b)] == list(range(n))
            assert all(0 < b - a <= batch for a, b in bs)
            if bs:
                sizes = [b - a for a, b in bs]
                assert max(sizes) - min(sizes) <= 1
    assert campaign.plan_batches(8192, 2048) == [(0, 2048), (2048, 4096), (4096, 6144), (6144, 8192)]
    with pytest.raises(ValueError):
        campaign.plan_batches(4, 0)


def test_auto_batch_fits_the_device():
    per = campaign.BYTES_PER_MEMBER
    assert 40.0e6 < per < 43.5e6                                   # 40.4 MB of corrections + state + outputs
    b = campaign.auto_batch(180 << 30)
    assert b % 148 == 0 and 3800 <= b <= 4400                     # "about 4,000 perturbed members fit one B200"
    assert b * per + campaign.SHARED_BYTES + (4 << 30) <= (180 << 30)
    assert campaign.auto_batch(1 << 30) == 1                      # never zero: init reports the shortage itself
    small = campaign.auto_batch(5 << 30)                          # less than one wave fits: no rounding down to 0
    assert 1 <= small < 148


def test_perturbed_member_draws_are_reproducible_and_in_range():
    p0, c0 = campaign.perturbed_member(0)
    p1, c1 = campaign.perturbed_member(0)
    assert c0 == c1 and p0.kappa == p1.kappa and p0.a_cloud == p1.a_cloud
    for g in (0, 1, 1023, 65535):
        p, co2 = campaign.perturbed_member(g)
        assert 280.0 <= co2 <= 1120.0 and 6e5 <= p.kappa <= 1e6
        # same stream as SURVEY.md 8d: default_rng(1000 + g), co2 first, kappa second
        rng = np.random.default_rng(1000 + g)
        assert co2 == float(rng.uniform(280.0, 1120.0))
        assert p.kappa == np.float32(rng.uniform(6e5, 1e6))


class _FakeHandle:
    """records the members of one batch; 'annual means' are a deterministic function of the member's co2"""
    created = []

    def __init__(self, n, device=0):
        self.n, self.co2 = n, np.zeros(n)
        self.years = 0
        _FakeHandle.created.append(self)

    def set_arithmetic(self, mode): pass
    def set_forcing(self, f): pass
    def set_member(self, m, p, co2, year0=1940): self.co2[m] = co2[0]
    def init(self): pass
    def spinup(self, years): self.spun = years
    def reset_scenario(self): pass
    def last_kernel_ms(self): return 1.0, 1
    def flags(self): return np.zeros(self.n, dtype=np.int32)
    def close(self): self.closed = True

    def run(self, years, want_output=True, out_members=None, out=None):
        gm = np.outer(self.co2, 1.0 + 0.01 * np.arange(years)).astype(np.float32)
        n_out = self.n if out_members is None else len(out_members)
        o = np.zeros((n_out, years, 12, 5, 48, 96), dtype=np.float32) if want_output else None
        if o is not None:
            for i, m in enumerate(out_members if out_members is not None else range(self.n)):
                o[i] = self.co2[m]
        return o, gm, gm * 2


def _member(g):
    return None, [100.0 + g]


def test_run_sharded_single_rank_bookkeeping(monkeypatch):
    monkeypatch.setattr(campaign._lib, "pad_co2", lambda co2, n: np.full(n, co2[0], dtype=np.float32))
    _FakeHandle.created.clear()
    r = campaign.run_sharded(10, _member, None, 3, 4, batch=4, out_stride=3, ensemble_cls=_FakeHandle)
    assert [h.n for h in _FakeHandle.created] == [4, 3, 3] and all(h.closed and h.spun == 3 for h in _FakeHandle.created)
    assert r["gmean"].shape == (10, 4) and np.allclose(r["gmean"][:, 0], 100.0 + np.arange(10))
    assert sorted(r["monthly"]) == [0, 3, 6, 9]
    assert all(np.all(r["monthly"][g] == 100.0 + g) for g in r["monthly"])
    assert r["launches"] == 6 and r["kernel_ms_spinup"] == 3.0 and r["kernel_ms_scenario"] == 3.0
    want = 2 * (100.0 + np.arange(10))
    assert np.allclose(r["moments"][0], [10, want.sum(), (want * want).sum()])
    mean, std = campaign.ensemble_mean_std(r["moments"])
    assert abs(mean[0] - want.mean()) < 1e-9 and abs(std[0] - want.std()) < 1e-6


def _worker(rank, world, port, n_total, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        campaign._lib.pad_co2 = lambda co2, n: np.full(n, co2[0], dtype=np.float32)
        r = campaign.run_sharded(n_total, _member, None, 1, 3, rank=rank, world=world, batch=5, out_stride=8,
                                 ensemble_cls=_FakeHandle)
        ret[rank] = dict(first=r["first"], last=r["last"], moments=r["moments"].tolist(), kept=sorted(r["monthly"]),
                         gm0=r["gmean"][:, 0].tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_run_sharded_world2_gloo():
    import torch.multiprocessing as mp
    n_total, world = 23, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    assert (ret[0]["first"], ret[0]["last"], ret[1]["first"], ret[1]["last"]) == (0, 12, 12, 23)
    assert ret[0]["gm0"] + ret[1]["gm0"] == [100.0 + g for g in range(n_total)]
    assert ret[0]["kept"] == [0, 8] and ret[1]["kept"] == [16]
    for y in range(3):
        v = 2 * (100.0 + np.arange(n_total)) * np.float32(1.0 + 0.01 * y)
        want = [n_total, v.sum(), (v * v).sum()]
        for r in range(world):                          # every rank holds the sums over ALL members
            assert np.allclose(ret[r]["moments"][y], want, rtol=1e-6)
    assert sharding.owner_of(16, n_total, world) == 1


@pytest.mark.gpu
def test_batched_run_equals_one_handle(forcing):
    import greb_b200
    n, years = 37, 2
    ens = greb_b200.Ensemble(n)
    ens.set_forcing(forcing)
    for m in range(n):
        p, co2 = campaign.perturbed_member(m)
        ens.set_member(m, p, np.full(years, co2, dtype=np.float32))
    ens.init()
    ens.spinup(1)
    ens.reset_scenario()
    out, gm, gc = ens.run(years, out_members=[0, 16, 32])
    ens.close()
    r = campaign.run_sharded(n, campaign.perturbed_member, forcing, 1, years, batch=10, out_stride=16, arith="exact")
    assert [b - a for a, b in r["batches"]] == [10, 9, 9, 9]
    assert np.array_equal(r["gmean"], gm) and np.array_equal(r["gmean_coslat"], gc)
    assert sorted(r["monthly"]) == [0, 16, 32]
    for i, g in enumerate((0, 16, 32)):
        assert np.array_equal(r["monthly"][g], out[i])
    assert r["flags"].sum() == 0 and r["launches"] == 4 * (1 + years)
    # a second "rank" of a 2-rank job computes exactly the members [19, 37)
    r1 = campaign.run_sharded(n, campaign.perturbed_member, forcing, 1, years, rank=1, world=2, batch=10,
                              out_stride=16, arith="exact", reduce=False)
    assert (r1["first"], r1["last"]) == (19, 37) and np.array_equal(r1["gmean"], gm[19:])
