"""GPU parity tests: the CUDA path, called through the C ABI (include/greb_b200.h), against the
CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * the circulation (diffusion + advection, 24 sub-steps) is pure + - * / arithmetic and must be
    BIT-EXACT with the oracle;
  * a whole step also evaluates exp/log, where the device libm and glibc differ in the last ulp,
    so whole-run results are held to the north_star tolerances: per-cell monthly
    Tsurf/Tatmos/Tocean <= 0.01 K, q <= 1e-6 kg/kg, cos-lat annual global-mean Tsurf <= 1e-3 K,
    identical sea-ice masks (cells within tolerance of a threshold excused).
"""
import numpy as np
import pytest

import greb_b200

pytestmark = pytest.mark.gpu

XD, YD, NT = 96, 48, 730
TOL_T, TOL_Q, TOL_GM = 1e-2, 1e-6, 1e-3
NAMES = ["Ts", "Ta", "To", "q", "cap_surf"]


def rand_field(rng, lo, hi):
    return rng.uniform(lo, hi, size=(YD, XD)).astype(np.float32)


def same_bits(a, b):
    """bitwise equality, treating +0 and -0 as equal"""
    return np.array_equal(np.where(a == 0, np.float32(0), a).view(np.uint32),
                          np.where(b == 0, np.float32(0), b).view(np.uint32))


def coslat_mean(field):
    w = np.cos(np.deg2rad((np.arange(YD) + 0.5) * 3.75 - 90))
    return float((field.astype(np.float64).mean(axis=-1) * w).sum(axis=-1) / w.sum())


def make_ensemble(forcing, physics_list, co2_list):
    ens = greb_b200.Ensemble(len(physics_list))
    ens.set_forcing(forcing)
    for m, (p, c) in enumerate(zip(physics_list, co2_list)):
        ens.set_member(m, p, c)
    ens.init()
    return ens


def product_physics(**kw):
    p = greb_b200.default_physics()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def check_monthly(out_g, out_o, z_topo, phys, label=""):
    """north_star gates on monthly means [..., 12, 5, 48, 96]; returns the observed maxima."""
    d = np.abs(out_g.astype(np.float64) - out_o.astype(np.float64))
    mx = [float(d[..., v, :, :].max()) for v in range(5)]
    assert mx[0] <= TOL_T and mx[1] <= TOL_T and mx[2] <= TOL_T, f"{label} temperature diffs {mx}"
    assert mx[3] <= TOL_Q, f"{label} humidity diff {mx[3]}"
    assert mx[4] <= 1e-4, f"{label} albedo diff {mx[4]}"
    # sea-ice masks on monthly-mean Tsurf over ocean: full ice, any ice, albedo >= 0.4
    oce = (z_topo < 0)
    Tg, To_ = out_g[..., 0, :, :], out_o[..., 0, :, :]
    for thr in (phys.To_ice1, phys.To_ice2):
        near = np.abs(To_ - np.float32(thr)) <= TOL_T
        assert np.array_equal(((Tg <= thr) & oce) | near, ((To_ <= thr) & oce) | near), f"{label} ice mask {thr}"
    ag, ao = out_g[..., 4, :, :], out_o[..., 4, :, :]
    near = np.abs(ao - 0.4) <= 1e-4
    assert np.array_equal((ag >= 0.4) | near, (ao >= 0.4) | near), f"{label} albedo mask"
    return mx


# ---- kernel level: circulation -----------------------------------------------------------------

@pytest.mark.parametrize("kappa", [8e5, 6.3e5, 1.2e6])
def test_circulation_bit_exact(oracle_mod, forcing, kappa):
    o = oracle_mod.Oracle(forcing, kappa=kappa)
    ens = make_ensemble(forcing, [product_physics(kappa=kappa)], [[680.0]])
    rng = np.random.default_rng(1)
    fields = [
        (forcing.tclim[10] + rand_field(rng, -1, 1), o.derived("wz_air")),
        (forcing.tclim[500] + rand_field(rng, -3, 3), o.derived("wz_air")),
        (forcing.qclim[300] * rand_field(rng, 0.5, 1.5), o.derived("wz_vapor")),
        (rand_field(rng, -1, 1), o.derived("wz_vapor")),  # mixed sign: the -0.9*T clamps fire
        (rand_field(rng, 1e-9, 2e-2), rand_field(rng, 0.4, 1.1)),
    ]
    X = np.stack([f[0] for f in fields])
    W = np.stack([f[1] for f in fields])
    for ityr in (1, 213, 730):
        got = ens.circulation(0, ityr, X, W)
        for i in range(len(fields)):
            ref = o.circulation(X[i], W[i], ityr)
            assert same_bits(ref, got[i]), (f"kappa={kappa} ityr={ityr} field {i}: "
                                           f"{np.count_nonzero(ref != got[i])} cells differ, "
                                           f"max {np.abs(ref - got[i]).max()} rows "
                                           f"{np.unique(np.where(ref != got[i])[0])[:12]}")
    ens.close()


def test_circulation_properties_full_batch(forcing):
    """Size-independent properties on a batch as wide as the GPU: a constant field stays exactly
    constant, and identical inputs in different CTAs give identical outputs."""
    ens = make_ensemble(forcing, [product_physics()], [[680.0]])
    rng = np.random.default_rng(2)
    n = 296
    base = forcing.tclim[77] + rand_field(rng, -1, 1)
    X = np.repeat(base[None], n, axis=0)
    X[1::2] = 287.25
    W = np.repeat(rand_field(rng, 0.5, 1.05)[None], n, axis=0)
    d = ens.circulation(0, 78, X, W)
    assert not d[1::2].any()
    assert all(np.array_equal(d[0], d[i]) for i in range(2, n, 2))
    assert np.abs(d[0]).max() > 0.05
    ens.close()


# ---- step level ----------------------------------------------------------------------------------

def test_single_steps_vs_oracle(oracle_mod, forcing):
    """time_loop step by step through greb_b200_time_loop; differences can only come from the last
    ulp of exp/log (device libm vs glibc)."""
    o = oracle_mod.Oracle(forcing)
    ens = make_ensemble(forcing, [product_physics()], [[680.0]])
    recs_o, recs_g = [], []
    for it in range(1, 125):
        r = o.time_loop(it, 680.0)
        ens.time_loop(it)
        if r is not None:
            recs_o.append(r)
            recs_g.append(ens.get_monthly(0)[0])
    assert len(recs_o) == 2
    worst = {}
    for n in NAMES:
        a, b = o.get(n), ens.get_state(0, n)
        worst[n] = float(np.abs(a.astype(np.float64) - b).max() / (1.0 if n != "cap_surf" else np.abs(a).max()))
    assert worst["Ts"] < 2e-3 and worst["Ta"] < 2e-3 and worst["To"] < 2e-3 and worst["q"] < 1e-7, worst
    assert worst["cap_surf"] < 2e-4, worst  # the sea-ice ramp amplifies a 1e-4 K difference in Ts
    check_monthly(np.stack(recs_g)[None], np.stack(recs_o)[None], forcing.z_topo, o.physics, "2 months")
    ens.close()


def test_spinup_and_scenario_vs_oracle(oracle_mod, forcing):
    """1-year flux-correction spin-up + 2 scenario years, perturbed physics and a CO2 ramp."""
    kw = dict(kappa=7.1e5, a_cloud=0.33, ct_sens=20.0)
    co2 = np.array([400.0, 560.0], dtype=np.float32)
    o = oracle_mod.Oracle(forcing, **kw)
    ens = make_ensemble(forcing, [product_physics(**kw)], [co2])
    o.spinup(1)
    ens.spinup(1)
    for n in NAMES:
        a, b = o.get(n), ens.get_state(0, n)
        assert np.abs(a.astype(np.float64) - b).max() <= (1e-3 if n != "cap_surf" else 1e-5 * np.abs(a).max()), n
    # TF = T_error*cap_surf/dt [W/m2]: 1e-4 K of T_error on a deep mixed layer (cap ~8e8) is ~2 W/m2
    tol = [2.0, 1e-6, 1e-3]
    for w in range(3):  # TF [W/m2], qF [kg/kg], ToF [K]
        a, b = o.fluxcorr(w), ens.get_fluxcorr(0, w)
        assert np.abs(a.astype(np.float64) - b).max() <= tol[w], (w, np.abs(a - b).max())
    out_o, gm_o = o.run(2, co2_ppm=co2)
    ens.reset_scenario()
    out_g, gm_g, gc_g = ens.run(2)
    check_monthly(out_g[0], out_o, forcing.z_topo, o.physics, "spinup+2yr")
    assert np.abs(gm_g[0] - gm_o).max() <= TOL_GM
    days = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    for y in range(2):
        want = sum(coslat_mean(out_o[y, mth, 0]) * days[mth] for mth in range(12)) / 365 - 273.15
        assert abs(gc_g[0, y] - want) <= 2e-3
    assert not ens.flags().any()
    ens.close()


def test_ensemble_members_are_independent_and_grouped(oracle_mod, forcing):
    """A mixed ensemble: members sharing physics share one spin-up (CO2-only variants) while
    perturbed members get their own; every member must match its own single-member oracle run."""
    specs = [dict(), dict(), dict(kappa=9.4e5, ce=2.2e-3), dict(a_cloud=0.31, da_ice=0.28, co_turb=4.3), dict()]
    co2s = [[680.0], [340.0], [680.0], [1000.0], [680.0]]
    ens = make_ensemble(forcing, [product_physics(**s) for s in specs], co2s)
    ens.spinup(1)
    ens.reset_scenario()
    out_g, gm_g, _ = ens.run(1)
    for m, (s, c) in enumerate(zip(specs, co2s)):
        o = oracle_mod.Oracle(forcing, **s)
        o.spinup(1)
        out_o, gm_o = o.run(1, co2_ppm=np.array(c, dtype=np.float32))
        check_monthly(out_g[m], out_o, forcing.z_topo, o.physics, f"member {m}")
        assert abs(gm_g[m, 0] - gm_o[0]) <= TOL_GM
    # identical members are bit-identical to each other, CO2 variants are not
    assert np.array_equal(out_g[0], out_g[4])
    assert not np.array_equal(out_g[0], out_g[1])
    ens.close()


def test_output_subset_and_chained_runs(forcing):
    """greb_b200_run with an out_members subset, and run(1)+run(1) == run(2)."""
    phys = [product_physics(), product_physics(kappa=9e5), product_physics()]
    co2 = [[680.0, 700.0]] * 3
    a = make_ensemble(forcing, phys, co2)
    a.reset_scenario()
    full, gm2, _ = a.run(2)
    b = make_ensemble(forcing, phys, co2)
    b.reset_scenario()
    o1, gma, _ = b.run(1, out_members=[2, 1])
    o2, gmb, _ = b.run(1, out_members=[2, 1])
    assert np.array_equal(o1[0, 0], full[2, 0]) and np.array_equal(o1[1, 0], full[1, 0])
    assert np.array_equal(o2[0, 0], full[2, 1]) and np.array_equal(o2[1, 0], full[1, 1])
    assert np.array_equal(np.concatenate([gma, gmb], axis=1), gm2)
    assert np.array_equal(b.get_monthly(1), full[1, 1])
    a.close()
    b.close()


def test_error_behaviour(forcing):
    ens = greb_b200.Ensemble(1)
    with pytest.raises(greb_b200.GrebError, match="set_forcing"):
        ens.init()
    ens.set_forcing(forcing)
    with pytest.raises(greb_b200.GrebError, match="init"):
        ens.spinup(1)
    ens.set_member(0, product_physics(), [680.0])
    ens.init()
    with pytest.raises(greb_b200.GrebError, match="CO2 paths"):
        ens.run(2)
    with pytest.raises(greb_b200.GrebError, match="bad arguments"):
        ens.get_state(3, "Ts")
    ens.close()


# ---- the reference's own default run (BASELINE.json configs[0]) ---------------------------------

@pytest.mark.slow
def test_default_50yr_2xco2_run_vs_oracle(oracle_mod, forcing):
    """namelist defaults: 3-year spin-up at 298 ppm, 50 years at 680 ppm from 1940 (reference
    namelist:1-14), on the synthetic S0 inputs.  All north_star gates over the whole run."""
    o = oracle_mod.Oracle(forcing)
    o.spinup(3)
    out_o, gm_o = o.run(50, co2_ppm=680.0)
    ens = make_ensemble(forcing, [product_physics()], [np.full(50, 680.0, dtype=np.float32)])
    ens.spinup(3)
    ens.reset_scenario()
    out_g, gm_g, gc_g = ens.run(50)
    mx = check_monthly(out_g[0], out_o, forcing.z_topo, o.physics, "50yr")
    days = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    worst = 0.0
    for y in range(50):
        want = sum(coslat_mean(out_o[y, mth, 0]) * days[mth] for mth in range(12)) / 365
        got = sum(coslat_mean(out_g[0, y, mth, 0]) * days[mth] for mth in range(12)) / 365
        worst = max(worst, abs(want - got))
    assert worst <= TOL_GM, worst
    assert np.abs(gm_g[0] - gm_o).max() <= TOL_GM
    # the 2xCO2 signal itself (sanity of the physics, not of parity): warming of a few K
    assert 0.5 < gm_o[-1] - gm_o[0] < 8.0
    print(f"\n50-yr parity: max |dT| surf/air/ocean = {mx[0]:.2e}/{mx[1]:.2e}/{mx[2]:.2e} K, "
          f"max |dq| = {mx[3]:.2e}, max |d albedo| = {mx[4]:.2e}, cos-lat global mean {worst:.2e} K")
    ens.close()


# ---- BASELINE.json configs[2] at full size: size-independent properties -------------------------

def test_full_size_ensemble_properties(oracle_mod, forcing):
    """1,024 perturbed-physics members (the bench workload) for one spin-up + one scenario year:
    (a) members that were given identical parameters are bit-identical wherever they sit in the batch;
    (b) permuting the batch permutes the results (members never interact, f:1030-1068);
    (c) no member is flagged non-finite and every global mean is physical;
    (d) three members, picked across the batch, match the oracle within the north_star tolerances."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import member_physics
    N = 1024
    rng = np.random.default_rng(5)
    src = np.arange(N)
    dup = rng.choice(N, size=32, replace=False)
    src[dup[16:]] = dup[:16]                       # 16 members are copies of 16 others
    specs = [member_physics(int(s), greb_b200.default_physics) for s in src]
    ens = greb_b200.Ensemble(N)
    ens.set_forcing(forcing)
    for m, (p, co2) in enumerate(specs):
        ens.set_member(m, p, [co2])
    ens.init()
    ens.spinup(1)
    ens.reset_scenario()
    pick = [int(dup[0]), int(dup[16]), 7, 500, 1023]
    out, gm, gc = ens.run(1, out_members=pick)
    states = ens.get_states()
    assert int(ens.flags().sum()) == 0
    assert np.all(np.isfinite(gm)) and np.all(np.abs(gm) < 60.0)
    for a, b in zip(dup[:16], dup[16:]):
        assert np.array_equal(states[a], states[b]) and gm[a, 0] == gm[b, 0], (a, b)
    assert np.array_equal(out[0], out[1])          # dup[0] and its copy dup[16]
    ens.close()
    # (b) the same members in reversed order
    M = 64
    sub = list(range(0, N, N // M))
    e1 = greb_b200.Ensemble(M)
    e2 = greb_b200.Ensemble(M)
    for e, order in ((e1, sub), (e2, sub[::-1])):
        e.set_forcing(forcing)
        for m, s in enumerate(order):
            e.set_member(m, specs[s][0], [specs[s][1]])
        e.init()
        e.spinup(1)
        e.reset_scenario()
        e.run(1, want_output=False)
    s1, s2 = e1.get_states(), e2.get_states()
    assert np.array_equal(s1, s2[::-1])
    e1.close()
    e2.close()
    # (d) oracle check of three members
    for i, m in ((2, 7), (3, 500), (4, 1023)):
        p, co2 = specs[m]
        kw = {n: getattr(p, n) for n in ("kappa", "ct_sens", "ce", "co_turb", "a_cloud", "da_ice")}
        o = oracle_mod.Oracle(forcing, **kw)
        o.spinup(1)
        out_o, gm_o = o.run(1, co2)
        check_monthly(out[i, 0], out_o[0], forcing.z_topo, o.physics, f"member {m}")
        assert abs(float(gm[m, 0]) - float(gm_o[0])) <= TOL_GM


def test_spinup_cache_restores_a_run_bit_for_bit(forcing):
    """flux corrections + state saved from one handle let another handle skip
    qflux_correction and continue with identical results."""
    p = product_physics(kappa=9.1e5, a_cloud=0.34)
    e1 = make_ensemble(forcing, [p], [[560.0, 560.0]])
    e1.spinup(1)
    corr = [e1.get_fluxcorr(0, w) for w in range(3)]
    state = {n: e1.get_state(0, n) for n in NAMES}
    e1.reset_scenario()
    out1, gm1, _ = e1.run(2)
    e1.close()
    e2 = make_ensemble(forcing, [p], [[560.0, 560.0]])       # no spin-up here
    for w in range(3):
        e2.set_fluxcorr(0, w, corr[w])
    for n, a in state.items():
        e2.set_state(0, n, a)
    e2.reset_scenario()
    out2, gm2, _ = e2.run(2)
    assert np.array_equal(out1, out2) and np.array_equal(gm1, gm2)
    e2.close()


@pytest.mark.gpu
def test_spinup_cache_on_disk(forcing, tmp_path):
    """host.save_spinup / load_spinup: a second handle restored from the file continues bit for bit"""
    from greb_b200 import host
    p = product_physics(kappa=8.6e5)
    key = host.spinup_key(forcing, p, 1)
    path = str(tmp_path / f"spinup_{key}.npz")
    e1 = make_ensemble(forcing, [p], [[500.0]])
    e1.spinup(1)
    host.save_spinup(path, e1, 0, key)
    e1.reset_scenario()
    out1, gm1, _ = e1.run(1)
    e1.close()
    e2 = make_ensemble(forcing, [p, p], [[500.0], [500.0]])   # two members of one physics group
    assert not host.load_spinup(path, e2, [0, 1], host.spinup_key(forcing, p, 2))
    assert host.load_spinup(path, e2, [0, 1], key)
    e2.reset_scenario()
    out2, gm2, _ = e2.run(1)
    e2.close()
    assert np.array_equal(out2[0], out1[0]) and np.array_equal(out2[1], out1[0]) and gm2[1, 0] == gm1[0, 0]
