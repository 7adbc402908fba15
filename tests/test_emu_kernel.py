"""Kernel LOGIC vs the oracle, on the CPU, through the test-only lane emulator.

tests/emu/greb_emu.cpp compiles the very same warp-level source the GPU runs
(greb-climate-model_b200/csrc/greb_core.h) with 32-lane arrays instead of registers and one
pthread per warp.  Both sides use glibc's expf/logf here, so the comparison is bit-exact for the
whole step, not just the stencils.  (The GPU run itself is compared in test_gpu_parity.py.)
"""
import numpy as np
import pytest

import emu_lib

XD, YD, NT = 96, 48, 730
NAMES = ["Ts", "Ta", "To", "q", "cap_surf"]


def rand_field(rng, lo, hi):
    return rng.uniform(lo, hi, size=(YD, XD)).astype(np.float32)


def test_row_assignment_covers_grid_and_keeps_warps_homogeneous(oracle_mod):
    for kappa in (8e5, 6e5, 1e6, 1.2e6, 2e6, 1e5):
        p = oracle_mod.default_physics()
        p.kappa = kappa
        rc, rows, hslot = emu_lib.row_tables(p)
        assert rc == 0
        assert sorted(rows) == list(range(YD))
        g = oracle_mod.geometry(kappa=kappa)
        for w in range(12):  # the f:592 branch never diverges inside a warp at the default geometry
            kinds = {g.polar[k] for k in rows[4 * w:4 * w + 4]}
            assert len(kinds) == 1, (w, rows[4 * w:4 * w + 4])
        # the pole rows and the rows with several polar sub-sub-steps are the ones the helper warps own
        want = [k for k in range(YD) if k in (0, YD - 1) or (g.polar[k] and g.time2_diff[k] > 1)]
        assert [k for k in range(YD) if hslot[k] >= 0] == want
        assert sorted(hslot[k] for k in want) == list(range(len(want)))
    p = oracle_mod.default_physics()
    p.kappa = 2e7  # absurd diffusivity: more multi-step rows than helper slots -> refused, not mis-computed
    assert emu_lib.row_tables(p)[0] < 0


@pytest.mark.parametrize("kappa", [8e5, 1.2e6])
def test_emulated_circulation_is_bit_exact(oracle_mod, forcing, kappa):
    o = oracle_mod.Oracle(forcing, kappa=kappa)
    rng = np.random.default_rng(0)
    cases = [
        (forcing.tclim[10] + rand_field(rng, -1, 1), o.derived("wz_air")),
        (forcing.qclim[300] * rand_field(rng, 0.5, 1.5), o.derived("wz_vapor")),
        (rand_field(rng, -1, 1), o.derived("wz_vapor")),                 # mixed sign: polar clamps fire
        (np.full((YD, XD), 281.0, dtype=np.float32), o.derived("wz_air")),  # constant -> exact zero
    ]
    for X, wz in cases:
        for ityr in (11, 400, 730):
            ref = o.circulation(X, wz, ityr)
            got = emu_lib.circulation(o.physics, forcing.uclim[ityr - 1], forcing.vclim[ityr - 1], X, wz)
            assert np.array_equal(ref.view(np.uint32) & 0x7fffffff | ((ref != 0) * (ref.view(np.uint32) & 0x80000000)),
                                  got.view(np.uint32) & 0x7fffffff | ((got != 0) * (got.view(np.uint32) & 0x80000000))), \
                f"{np.count_nonzero(ref != got)} cells differ"
    assert not emu_lib.circulation(o.physics, forcing.uclim[0], forcing.vclim[0], cases[3][0], cases[3][1]).any()


def test_emulated_time_loop_is_bit_exact(oracle_mod, forcing):
    o = oracle_mod.Oracle(forcing)
    e = emu_lib.Emu(forcing, o.physics, np.full(2, 680.0))
    recs = []
    for it in range(1, 65):
        r = o.time_loop(it, 680.0)
        if r is not None:
            recs.append(r)
    eo = e.steps(1, 64, spinup=False, out_months=2)
    for i, n in enumerate(NAMES):
        assert np.array_equal(o.get(n), e.get(i)), n
    assert len(recs) == 1 and np.array_equal(recs[0], eo[0])


def test_emulated_spinup_and_scenario_are_bit_exact(oracle_mod, forcing):
    o = oracle_mod.Oracle(forcing, kappa=1.2e6, a_cloud=0.33, ct_sens=20.0)
    co2 = np.array([400.0], dtype=np.float32)
    e = emu_lib.Emu(forcing, o.physics, co2)
    o.spinup(1)
    e.steps(1, NT, spinup=True)
    for i, n in enumerate(NAMES):
        assert np.array_equal(o.get(n), e.get(i)), n
    c = e.corr()
    assert np.array_equal(o.fluxcorr(0), c[:, 0]) and np.array_equal(o.fluxcorr(2), c[:, 1])
    assert np.array_equal(o.fluxcorr(1), c[:, 2])
    out, gm = o.run(1, co2_ppm=co2)
    e.reset_scenario()
    eo = e.steps(1, NT, spinup=False, out_months=12)
    assert np.array_equal(out[0], eo)
    assert e.diag()[0] == gm[0]
    w = np.cos(np.deg2rad((np.arange(YD) + 0.5) * 3.75 - 90))
    want = (out[0, :, 0].astype(np.float64).mean(axis=2) * w).sum(axis=1) / w.sum()   # monthly, cos-lat
    days = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    assert abs((want * days).sum() / 365 - 273.15 - e.diag()[1]) < 2e-3
