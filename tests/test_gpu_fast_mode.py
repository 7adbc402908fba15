"""GREB_ARITH_FAST (factored stencils, FMA contraction, approximate division/log/exp in the column
physics): not bit-identical to the reference by construction, held to the tolerances BASELINE.json
states — per-cell monthly Tsurf/Tatmos/Tocean <= 0.01 K, q <= 1e-6 kg/kg, cos-lat / console global
mean <= 1e-3 K over the 50-year 2xCO2 run, identical sea-ice masks — against the oracle AND against
the golden vectors produced by the reference's own source (tests/golden)."""
import os

import numpy as np
import pytest

import greb_b200
from test_gpu_parity import TOL_GM, check_monthly, coslat_mean, product_physics, rand_field

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fast_ensemble(forcing, physics_list, co2_list):
    ens = greb_b200.Ensemble(len(physics_list))
    ens.set_arithmetic("fast")
    ens.set_forcing(forcing)
    for m, (p, c) in enumerate(zip(physics_list, co2_list)):
        ens.set_member(m, p, c)
    ens.init()
    return ens


@pytest.mark.parametrize("kappa", [8e5, 6.3e5, 1.2e6])
def test_circulation_within_rounding_of_the_oracle(oracle_mod, forcing, kappa):
    """24 sub-steps of diffusion + advection: the factored/FMA form may differ from the as-written
    evaluation by accumulated rounding only: <= 32 ulp of the field (measured: 10)."""
    o = oracle_mod.Oracle(forcing, kappa=kappa)
    ens = fast_ensemble(forcing, [product_physics(kappa=kappa)], [[680.0]])
    rng = np.random.default_rng(1)
    fields = [
        (forcing.tclim[10] + rand_field(rng, -1, 1), o.derived("wz_air")),
        (forcing.qclim[300] * rand_field(rng, 0.5, 1.5), o.derived("wz_vapor")),
        (rand_field(rng, 1e-9, 2e-2), rand_field(rng, 0.4, 1.1)),
    ]
    X = np.stack([f[0] for f in fields])
    W = np.stack([f[1] for f in fields])
    for ityr in (1, 213, 730):
        got = ens.circulation(0, ityr, X, W)
        for i in range(len(fields)):
            ref = o.circulation(X[i], W[i], ityr)
            tol = 32 * np.spacing(np.float32(np.abs(X[i]).max()))
            d = np.abs(got[i].astype(np.float64) - ref)
            assert d.max() <= tol, (kappa, ityr, i, d.max(), tol)
    ens.close()


def test_clamp_and_mixed_sign_fields_stay_close(oracle_mod, forcing):
    """fields that trigger where(d <= -T) d = -0.9*T (f:715, f:907): the clamp is kept in the fast mode"""
    o = oracle_mod.Oracle(forcing)
    ens = fast_ensemble(forcing, [product_physics()], [[680.0]])
    rng = np.random.default_rng(2)
    q = (forcing.qclim[100] * rand_field(rng, 0.5, 1.5)).astype(np.float32)
    q[rng.integers(0, 48, 60), rng.integers(0, 96, 60)] *= 1e-4
    wz = o.derived("wz_vapor")
    got = ens.circulation(0, 100, q[None], wz[None])[0]
    ref = o.circulation(q, wz, 100)
    assert np.all(q + got >= 0) and np.all(q + ref >= 0)
    assert np.abs(got - ref).max() <= 1e-7
    ens.close()


@pytest.mark.slow
def test_default_50yr_run_fast_mode_meets_the_north_star_gates(oracle_mod, forcing):
    o = oracle_mod.Oracle(forcing)
    o.spinup(3)
    out_o, gm_o = o.run(50, co2_ppm=680.0)
    ens = fast_ensemble(forcing, [product_physics()], [np.full(50, 680.0, dtype=np.float32)])
    ens.spinup(3)
    ens.reset_scenario()
    out_g, gm_g, _ = ens.run(50)
    mx = check_monthly(out_g[0], out_o, forcing.z_topo, o.physics, "fast 50yr")
    days = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    worst = 0.0
    for y in range(50):
        want = sum(coslat_mean(out_o[y, m, 0]) * days[m] for m in range(12)) / 365
        got = sum(coslat_mean(out_g[0, y, m, 0]) * days[m] for m in range(12)) / 365
        worst = max(worst, abs(want - got))
    assert worst <= TOL_GM and np.abs(gm_g[0] - gm_o).max() <= TOL_GM
    # ... and against the reference-derived golden December fields
    g = np.load(os.path.join(GOLD, "ref_config1.npz"), allow_pickle=False)
    for y in (1, 10, 50):
        d = np.abs(out_g[0, y - 1, 11].astype(np.float64) - g[f"dec_year{y}"])
        assert d[:3].max() <= 1e-2 and d[3].max() <= 1e-6 and d[4].max() <= 1e-4, (y, d.max(axis=(1, 2)))
    print(f"\nfast mode, 50-yr parity: max |dT| surf/air/ocean = {mx[0]:.2e}/{mx[1]:.2e}/{mx[2]:.2e} K, "
          f"max |dq| = {mx[3]:.2e}, cos-lat global mean {worst:.2e} K")
    ens.close()


def test_fast_and_exact_modes_can_alternate_on_one_handle(forcing):
    ens = fast_ensemble(forcing, [product_physics()], [[680.0, 680.0]])
    ens.spinup(1)
    ens.reset_scenario()
    ens.run(1, want_output=False)
    a = ens.get_state(0, "Ta")
    ens.set_arithmetic("exact")
    ens.run(1, want_output=False)
    b = ens.get_state(0, "Ta")
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b)) and int(ens.flags().sum()) == 0
    with pytest.raises(KeyError):
        ens.set_arithmetic("sloppy")
    ens.close()
