"""Big-grid circulation path (BASELINE.json configs[4]): oracle pinning, band decomposition with
communication-avoiding halos (gloo, world_size 2 and 3 on CPU), and the CUDA kernel through the
C ABI of include/greb_grid.h (bit-exact against oracle/grid_oracle.c)."""
import os
import socket

import numpy as np
import pytest

from greb_b200 import bigrid
from oracle import grid as og


def fields(nx, ny, seed=3):
    """smooth synthetic Ta-like field, wz in (0.6, 1.3), winds of both signs (also on the polar rows)"""
    rng = np.random.default_rng(seed)
    lat = (np.arange(ny) + 0.5) / ny * np.pi - np.pi / 2
    lon = np.arange(nx) / nx * 2 * np.pi
    X = (288 - 40 * np.sin(lat)[:, None] ** 2 + 3 * np.cos(3 * lon)[None, :] * np.cos(lat)[:, None]
         + rng.normal(0, 0.3, (ny, nx))).astype(np.float32)
    wz = (0.95 + 0.3 * np.sin(2 * lon)[None, :] * np.cos(lat)[:, None] ** 2 + rng.uniform(-0.05, 0.05, (ny, nx))).astype(np.float32)
    u = (8 * np.cos(3 * lat)[:, None] + rng.normal(0, 2, (ny, nx))).astype(np.float32)
    v = (2 * np.sin(6 * lat)[:, None] * np.cos(lat)[:, None] + rng.normal(0, 1, (ny, nx))).astype(np.float32)
    return X, wz, u, v


def test_rules_vanish_on_the_reference_grid(oracle_mod, forcing):
    """at 96x48 the generalised oracle IS the pinned reference restatement, bit for bit"""
    g = og.Geometry(96, 48)
    gr = oracle_mod.geometry()
    assert g.nsub == 24 and g.dt_crcl == 1800.0
    for name in ("dxlat", "ccx_diff", "ccx_adv", "ccx2_diff", "ccx2_adv", "polar", "time2_diff", "time2_adv"):
        assert np.array_equal(np.array(getattr(gr, name)[:]), getattr(g, name)), name
    o = oracle_mod.Oracle(forcing)
    for ityr, kappa in ((1, 8e5), (400, 8e5)):
        for X, wz in ((forcing.tclim[ityr - 1], o.derived("wz_air")), (forcing.qclim[ityr - 1], o.derived("wz_vapor"))):
            want = o.circulation(X, wz, ityr)
            got = og.substeps(g, X, wz, forcing.uclim[ityr - 1], forcing.vclim[ityr - 1], g.nsub) - X
            assert np.array_equal(want, got)


def test_quarter_degree_geometry():
    g = og.Geometry(1440, 720)
    assert g.nsub == 5400 and g.dt_crcl == 8.0 and abs(g.ccy_diff - 0.0082819) < 1e-6
    assert g.polar.all() and g.time2_diff.max() == 8 and g.time2_adv.max() == 1
    assert (g.time2_diff[:7] == 8).all() and (g.time2_diff[100:620] == 1).all()     # only the rows next to the poles iterate
    assert g.ccx2_diff.max() < 1.5                                                    # like the reference's own rows 2/47


def test_band_decomposition_in_process():
    """two and three OracleBands with manual halo copies == the undivided domain, bit for bit"""
    from grid_band import OracleBand
    nx, ny, s, n = 48, 40, 3, 8
    X, wz, u, v = fields(nx, ny)
    g = og.Geometry(nx, ny)
    want = og.substeps(g, X, wz, u, v, n)
    for world in (2, 3):
        bands = []
        for r in range(world):
            k0, k1 = bigrid.band_range(ny, world, r)
            b = OracleBand(nx, ny, k0, k1, s)
            b.set_fields(X, wz, u, v)
            bands.append(b)
        done = 0
        while done < n:
            m = min(s, n - done)
            for r in range(world - 1):                        # exchange 2*s rows across every inner boundary
                lo, hi = bands[r], bands[r + 1]
                hi.rows(hi.k0 - 2 * s, hi.k0)[:] = lo.rows(lo.k1 - 2 * s, lo.k1)
                lo.rows(lo.k1, lo.k1 + 2 * s)[:] = hi.rows(hi.k0, hi.k0 + 2 * s)
            for b in bands:
                b.halo_refreshed()
                b.substeps(m)
            done += m
        got = np.concatenate([b.get() for b in bands])
        assert np.array_equal(got, want), world


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from grid_band import OracleBand
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, s, n = 48, 40, 2, 7
        X, wz, u, v = fields(nx, ny)
        k0, k1 = bigrid.band_range(ny, world, rank)
        b = OracleBand(nx, ny, k0, k1, s)
        b.set_fields(X, wz, u, v)
        ex = bigrid.advance(b, n, rank, world)
        # two fields with overlapped exchanges (the second one circulates wz itself, like q with wz_vapor)
        pair = [OracleBand(nx, ny, k0, k1, s), OracleBand(nx, ny, k0, k1, s)]
        pair[0].set_fields(X, wz, u, v)
        pair[1].set_fields(wz, wz, u, v)
        ex2 = bigrid.advance_overlapped(pair, n, rank, world)
        ret[rank] = (k0, k1, ex, b.get(), ex2, pair[0].get(), pair[1].get())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_advance_with_halo_exchange_gloo(world):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    nx, ny, n = 48, 40, 7
    X, wz, u, v = fields(nx, ny)
    want = og.substeps(og.Geometry(nx, ny), X, wz, u, v, n)
    got = np.concatenate([ret[r][3] for r in range(world)])
    assert [ret[r][:2] for r in range(world)] == [bigrid.band_range(ny, world, r) for r in range(world)]
    assert all(ret[r][2] == 4 for r in range(world))          # ceil(7 / 2) exchanges
    assert np.array_equal(got, want)
    assert all(ret[r][4] == 8 for r in range(world))
    assert np.array_equal(np.concatenate([ret[r][5] for r in range(world)]), want)
    want2 = og.substeps(og.Geometry(nx, ny), wz, wz, u, v, n)
    assert np.array_equal(np.concatenate([ret[r][6] for r in range(world)]), want2)


def test_thin_band_is_refused():
    from grid_band import OracleBand
    b = OracleBand(48, 40, 0, 5, 4)
    with pytest.raises(ValueError):
        bigrid.advance(b, 4, 0, 2)


def test_grid_library_exports_the_abi():
    """include/greb_grid.h <-> libgreb_grid.so (loads without a GPU; no compute calls here)"""
    import re
    L = bigrid.load_grid_library()
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "greb_grid.h")).read()
    declared = sorted(set(re.findall(r"\b(greb_grid_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(bigrid.GRID_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,n", [(96, 48, 24), (192, 96, 10), (1440, 64, 3), (1536, 24, 2), (40, 720, 2),
                                     (1440, 720, 4)])
def test_device_band_is_bit_exact(nx, ny, n):
    X, wz, u, v = fields(nx, ny, seed=nx + ny)
    g = og.Geometry(nx, ny)
    b = bigrid.DeviceBand(nx, ny, 0, ny, 4)
    assert (b.nsub, b.dt_crcl) == (g.nsub, g.dt_crcl)
    b.set_fields(X, wz, u, v)
    bigrid.advance(b, n)
    want = og.substeps(g, X, wz, u, v, n)
    got = b.get()
    b.close()
    assert np.array_equal(got, want), np.abs(got - want).max()


@pytest.mark.gpu
def test_device_reference_grid_equals_the_member_kernel(forcing):
    """96x48: the big-grid kernel and the ensemble kernel's circulation entry agree bit for bit"""
    import greb_b200
    ens = greb_b200.Ensemble(1)
    ens.set_forcing(forcing)
    ens.set_member(0, greb_b200.default_physics(), [680.0])
    ens.init()
    ityr = 200
    wz = np.exp(-forcing.z_topo / np.float32(8400.0)).astype(np.float32)
    X = forcing.tclim[ityr - 1]
    want = ens.circulation(0, ityr, X, wz)
    ens.close()
    b = bigrid.DeviceBand(96, 48, 0, 48, 4)
    b.set_fields(X, wz, forcing.uclim[ityr - 1], forcing.vclim[ityr - 1])
    bigrid.advance(b, b.nsub)
    got = b.get() - X
    b.close()
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_device_bands_with_halos_equal_the_undivided_domain():
    """three DeviceBands on one GPU, halos copied by hand == one band, bit for bit"""
    nx, ny, s, n = 192, 96, 4, 10
    X, wz, u, v = fields(nx, ny)
    one = bigrid.DeviceBand(nx, ny, 0, ny, s)
    one.set_fields(X, wz, u, v)
    bigrid.advance(one, n)
    want = one.get()
    one.close()
    bands = []
    for r in range(3):
        k0, k1 = bigrid.band_range(ny, 3, r)
        b = bigrid.DeviceBand(nx, ny, k0, k1, s)
        b.set_fields(X, wz, u, v)
        bands.append(b)
    done = 0
    while done < n:
        m = min(s, n - done)
        for r in range(2):
            lo, hi = bands[r], bands[r + 1]
            hi.rows(hi.k0 - 2 * s, hi.k0).copy_(lo.rows(lo.k1 - 2 * s, lo.k1))
            lo.rows(lo.k1, lo.k1 + 2 * s).copy_(hi.rows(hi.k0, hi.k0 + 2 * s))
        import torch
        torch.cuda.synchronize()
        for b in bands:
            b.halo_refreshed()
            b.substeps(m)
        done += m
    got = np.concatenate([b.get() for b in bands])
    with pytest.raises(greb_b200_error()):
        bands[1].substeps(s + 1)                              # halo used up without an exchange
    for b in bands:
        b.close()
    assert np.array_equal(got, want)


def greb_b200_error():
    import greb_b200
    return greb_b200.GrebError


@pytest.mark.gpu
def test_overlapped_fields_equal_sequential_runs():
    nx, ny, s, n = 192, 96, 4, 11
    X, wz, u, v = fields(nx, ny)
    g = og.Geometry(nx, ny)
    pair = [bigrid.DeviceBand(nx, ny, 0, ny, s), bigrid.DeviceBand(nx, ny, 0, ny, s)]
    pair[0].set_fields(X, wz, u, v)
    pair[1].set_fields(wz, wz, u, v)
    assert bigrid.advance_overlapped(pair, n) == 0
    got = [b.get() for b in pair]
    assert pair[0].launches == n and pair[0].kernel_ms > 0
    for b in pair:
        b.close()
    assert np.array_equal(got[0], og.substeps(g, X, wz, u, v, n))
    assert np.array_equal(got[1], og.substeps(g, wz, wz, u, v, n))


@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,n", [(96, 48, 24), (192, 96, 11), (1440, 720, 6)])
def test_persistent_kernel_equals_the_launch_per_substep_path(nx, ny, n):
    """one cooperative launch (grid barrier between sub-steps, two fields as work items) == n launches, bit for bit;
    the two paths can alternate on one handle (buffer parity is kept)"""
    X, wz, u, v = fields(nx, ny, seed=nx)
    g = og.Geometry(nx, ny)
    want = [og.substeps(g, X, wz, u, v, n + 3), og.substeps(g, wz, wz, u, v, n + 3)]
    pair = [bigrid.DeviceBand(nx, ny, 0, ny, 1), bigrid.DeviceBand(nx, ny, 0, ny, 1)]
    pair[0].set_fields(X, wz, u, v)
    pair[1].set_fields(wz, wz, u, v)
    grp = bigrid.PersistentGroup(pair)
    grp.advance(n)
    assert grp.launches == 1 and grp.kernel_ms > 0
    for b in pair:                                   # 3 more sub-steps on the v1 path, then nothing breaks
        bigrid.advance(b, 3)
    got = [b.get() for b in pair]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    grp.advance(2)                                   # and back to the persistent path
    assert np.array_equal(pair[0].get(), og.substeps(g, X, wz, u, v, n + 5))
    single = bigrid.DeviceBand(nx, ny, 0, ny, 1)     # a group of one field
    single.set_fields(X, wz, u, v)
    bigrid.PersistentGroup([single]).advance(n + 5)
    assert np.array_equal(single.get(), pair[0].get())
    for b in pair + [single]:
        b.close()


@pytest.mark.gpu
def test_persistent_inner_band_without_neighbours_is_refused():
    b = bigrid.DeviceBand(192, 96, 24, 48, 1)
    X, wz, u, v = fields(192, 96)
    b.set_fields(X, wz, u, v)
    with pytest.raises(greb_b200_error()):
        bigrid.PersistentGroup([b]).advance(1)
    b.close()


def _ipc_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeBand:                                  # what connect_neighbours needs of a DeviceBand
        def __init__(self, f):
            self.f, self.imported = f, {}
        def ipc_export(self):
            return f"blob-rank{rank}-field{self.f}".encode()
        def ipc_import(self, side, blob):
            self.imported[side] = blob.decode()
    try:
        bands = [FakeBand(0), FakeBand(1)]
        bigrid.connect_neighbours(bands, rank, world)
        ret[rank] = [b.imported for b in bands]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_neighbour_handle_exchange_gloo_world3():
    """the host side of the persistent path: every rank ends up with its south and north neighbours' blobs"""
    import torch.multiprocessing as mp
    world = 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_ipc_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        for f in range(2):
            want = {}
            if r > 0:
                want[0] = f"blob-rank{r - 1}-field{f}"
            if r < world - 1:
                want[1] = f"blob-rank{r + 1}-field{f}"
            assert ret[r][f] == want, (r, f, ret[r][f])


@pytest.mark.gpu
def test_whole_step_on_tiles_equals_the_member_kernel_at_96x48(forcing):
    """BigStep = column physics on tiles (the member kernel's own device functions) + the big-grid circulations:
    on the reference grid six 12-hour steps are bit-identical to the ensemble kernel (exact mode)"""
    import greb_b200
    ens = greb_b200.Ensemble(1)
    ens.set_forcing(forcing)
    ens.set_member(0, greb_b200.default_physics(), [680.0])
    ens.init()
    names = ("Ts", "Ta", "To", "q", "cap_surf")
    state0 = {n: ens.get_state(0, n) for n in names}
    static, step_forcing = bigrid.s0_static_and_forcing(forcing, 96, 48)
    big = bigrid.BigStep(96, 48, static, state0)
    assert big.nt == 1 and big.nsub == 24
    for it in range(1, 7):
        ens.time_loop(it)
        big.step(it, step_forcing(it), 680.0)
        for n in names:
            a, b = ens.get_state(0, n), big.field(n)
            assert np.array_equal(np.where(a == 0, np.float32(0), a), np.where(b == 0, np.float32(0), b)), \
                (it, n, float(np.abs(a - b).max()))
    # the same six steps with the forcing of step it+1 staged by a worker thread while step it runs
    pre = bigrid.BigStep(96, 48, static, state0)
    pre.run(1, 6, step_forcing, 680.0, prefetch=True)
    for n in names:
        assert np.array_equal(pre.field(n), big.field(n)), n
    pre.close()
    big.close()
    ens.close()


@pytest.mark.gpu
def test_column_physics_on_a_big_grid_is_cell_local(forcing):
    """1440x720 built from 15x15 copies of the 96x48 inputs: phase A (SW, LW, sensible, hydro, deep ocean, Ts/To/
    cap update, sea ice) of every copy equals the 96x48 result bit for bit — tiles of 4,608 cells, segments
    of 96, padding and the per-segment solar values are all in the right place"""
    import greb_b200
    names = ("Ts", "Ta", "To", "q", "cap_surf")
    static, step_forcing = bigrid.s0_static_and_forcing(forcing, 96, 48)
    rng = np.random.default_rng(4)
    state0 = {"Ts": forcing.tclim[729] + rng.uniform(-3, 3, (48, 96)).astype(np.float32), "Ta": forcing.tclim[729].copy(),
              "To": forcing.tclim[729] - np.float32(1.0), "q": forcing.qclim[729].copy(),
              "cap_surf": np.where(forcing.z_topo > 0, np.float32(4.8e6), np.float32(2.1e8)).astype(np.float32)}
    f1 = step_forcing(100)
    small = bigrid.BigStep(96, 48, static, state0)
    small.step(100, f1, 560.0)
    tile = lambda a: np.tile(a, (15, 15))
    staticB = {k: tile(v) for k, v in static.items()}
    stateB = {k: tile(v) for k, v in state0.items()}
    # np.tile repeats the 48 rows 15 times: row r of the big grid = row r % 48 of the small one
    fB = {k: (np.tile(v, 15) if k == "solar" else tile(v)) for k, v in f1.items()}
    big = bigrid.BigStep(1440, 720, staticB, stateB)
    assert big.nt == 225
    # phase A only (whole steps at 1440x720 are timed by tools/run_bigrid.py): forcing as step() loads it
    big._load_forcing(fB)
    big._phase(0, 560.0)
    small2 = bigrid.BigStep(96, 48, static, state0)
    small2._load_forcing(f1)
    small2._phase(0, 560.0)
    for n in ("Ts", "To", "cap_surf"):
        assert np.array_equal(big.field(n), tile(small2.field(n))), n
    assert np.array_equal(big.stash[:, 0].reshape(-1).cpu().numpy().reshape(720, 1440),
                          tile(small2.stash[:, 0].reshape(48, 96).cpu().numpy()))
    for b in (small, small2, big):
        b.close()
