"""Full-length parity of PERTURBED members in BOTH arithmetic modes.

bench.py times 1,024 members drawn by greb_b200.campaign.perturbed_member (kappa 6e5..1e6, CO2 280..1120
ppm, da_ice, a_cloud +-0.05, ct_sens, ce, co_turb +-20 %).  The switches of the column physics
(`Ts >= To_ice2` in deep_ocean, src/greb.f90:511-514; the ice-albedo and heat-capacity ramps,
:384-392, :483-490) can amplify last-ulp differences over decades, so the gate is the whole run:
16 members — the extreme draws of every perturbed parameter among the first 4,096 members plus the
first few members — run 3 spin-up + 50 scenario years on the GPU in the exact AND the fast mode and
are compared with the CPU oracle month by month:

    per-cell monthly Tsurf / Tatmos / Tocean <= 0.01 K, q <= 1e-6 kg/kg, albedo <= 1e-4,
    console global mean (f:954) and cos-lat annual mean <= 1e-3 K, identical sea-ice masks.

Result (B200, round 2): the EXACT mode is bit-identical to the oracle for all 16 members and all 53
years (it restates glibc's expf/logf on the device); the FAST mode keeps 14 members within 1.5e-3 K but
loses the two members below 300 ppm (bistable sea-ice edge) — so the exact mode is the default
everywhere and the fast mode an opt-in whose limits this test records.

Config 2 (greb-original control + scenario) is repeated in the fast mode against the reference-derived
golden fixture.  The worst observed margins go to gpurun_out/r02_parity_margins.json (copied to
profiles/ and quoted by bench.py)."""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import greb_b200
from greb_b200 import campaign
from test_gpu_parity import TOL_GM, TOL_Q, TOL_T, check_monthly, coslat_mean, same_bits

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PERTURBED = ("kappa", "ct_sens", "ce", "co_turb", "a_cloud", "da_ice")
SPINUP, YEARS = 3, 50


def pick_members(pool: int = 4096, n: int = 16):
    """extremes of every perturbed quantity over the first `pool` members, then members 0, 1, 2, ..."""
    draws = []
    for g in range(pool):
        p, co2 = campaign.perturbed_member(g)
        draws.append([co2] + [getattr(p, k) for k in PERTURBED])
    d = np.array(draws)
    chosen = []
    for col in (0, 1, 6, 5, 2, 3, 4):                       # co2, kappa, da_ice, a_cloud, then the rest
        for g in (int(d[:, col].argmin()), int(d[:, col].argmax())):
            if g not in chosen and len(chosen) < n:
                chosen.append(g)
    g = 0
    while len(chosen) < n:
        if g not in chosen:
            chosen.append(g)
        g += 1
    return chosen


def oracle_run(oracle_mod, forcing, g):
    p, co2 = campaign.perturbed_member(g)
    o = oracle_mod.Oracle(forcing, **{k: getattr(p, k) for k in PERTURBED})
    o.spinup(SPINUP)
    out, gm = o.run(YEARS, co2_ppm=co2)
    return out, gm, o.physics


def gpu_run(forcing, members, arith):
    ens = greb_b200.Ensemble(len(members))
    ens.set_arithmetic(arith)
    ens.set_forcing(forcing)
    for m, g in enumerate(members):
        p, co2 = campaign.perturbed_member(g)
        ens.set_member(m, p, np.full(YEARS, co2, dtype=np.float32))
    ens.init()
    ens.spinup(SPINUP)
    ens.reset_scenario()
    out, gm, gc = ens.run(YEARS)
    assert int(ens.flags().sum()) == 0
    ens.close()
    return out, gm, gc


def _write_margins(key, value):
    path = os.path.join(ROOT, "gpurun_out", "r02_parity_margins.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
    data[key] = value
    with open(path, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)


@pytest.fixture(scope="module")
def oracle_runs(oracle_mod, forcing):
    members = pick_members()
    with ThreadPoolExecutor(max_workers=min(len(members), os.cpu_count() or 1)) as ex:   # ctypes drops the GIL
        res = list(ex.map(lambda g: oracle_run(oracle_mod, forcing, g), members))
    return members, res


@pytest.mark.parametrize("arith", ["exact", "fast"])
def test_16_perturbed_members_3_plus_50_years(oracle_runs, forcing, arith):
    members, ref = oracle_runs
    out, gm, gc = gpu_run(forcing, members, arith)
    days = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    worst = np.zeros(5)
    worst_gm = worst_cos = 0.0
    per_member = {}
    failures = []
    bit_identical = True
    for m, g in enumerate(members):
        out_o, gm_o, phys = ref[m]
        bit_identical = bit_identical and same_bits(out[m], out_o) and np.array_equal(gm[m], gm_o)
        try:
            mx = check_monthly(out[m], out_o, forcing.z_topo, phys, f"{arith} member {g}")
        except AssertionError as e:                     # collect every member before failing
            failures.append(str(e))
            d = np.abs(out[m].astype(np.float64) - out_o.astype(np.float64))
            mx = [float(d[..., v, :, :].max()) for v in range(5)]
        worst = np.maximum(worst, mx)
        dgm = float(np.abs(gm[m].astype(np.float64) - gm_o).max())
        dcos = 0.0
        for y in range(YEARS):
            want = sum(coslat_mean(out_o[y, k, 0]) * days[k] for k in range(12)) / 365
            got = sum(coslat_mean(out[m, y, k, 0]) * days[k] for k in range(12)) / 365
            dcos = max(dcos, abs(want - got))
        if not (dgm <= TOL_GM and dcos <= TOL_GM):
            failures.append(f"{arith} member {g}: global mean {dgm} / {dcos}")
        worst_gm, worst_cos = max(worst_gm, dgm), max(worst_cos, dcos)
        p, co2 = campaign.perturbed_member(g)
        per_member[str(g)] = {"co2": co2, "kappa": p.kappa, "da_ice": p.da_ice, "max_dT": float(max(mx[:3])),
                              "max_dq": float(mx[3]), "gmean": dgm}
    rec = {"members": members, "years": f"{SPINUP}+{YEARS}", "max_dT_surf_air_ocean_K": [float(x) for x in worst[:3]],
           "max_dq": float(worst[3]), "max_dalbedo": float(worst[4]), "max_dgmean_console_K": worst_gm,
           "max_dgmean_coslat_K": worst_cos, "gates": {"T": 1e-2, "q": 1e-6, "gmean": 1e-3, "ice_masks": "identical"},
           "bit_identical": bool(bit_identical), "members_outside_the_gates": len(failures),
           "per_member": per_member}
    _write_margins(f"perturbed_{arith}", rec)
    print(f"\n{arith}: 16 perturbed members x ({SPINUP}+{YEARS}) years vs oracle: max |dT| = {worst[:3].max():.2e} K, "
          f"|dq| = {worst[3]:.2e}, console mean {worst_gm:.2e} K, cos-lat mean {worst_cos:.2e} K, "
          f"bit-identical: {bit_identical}")
    if arith == "exact":
        # with glibc's expf/logf restated on the device (greb_simt.h) the exact mode reproduces the reference's
        # arithmetic operation for operation: 53 years of every member, bit for bit
        assert bit_identical, failures
        assert not failures, failures
        return
    # The fast mode is NOT a drop-in for every member: below ~300 ppm the sea-ice edge of this model is
    # bistable and any arithmetic that is not bit-identical ends up on another branch in a few cells (the
    # reference arithmetic with the CUDA libm instead of glibc's already moved member 22 by 0.021 K in 50
    # years; the factored stencils move members 22 and 2989 by 0.45 K / 1.7 K).  That is why the exact mode
    # is the library's and the bench's default.  What the fast mode is held to: every member with
    # CO2 >= 350 ppm inside the north_star gates, no member non-finite, every member's global mean within
    # 0.01 K; the members outside the per-cell gates are recorded in the margins file, not hidden.
    strict = [str(g) for g in members if campaign.perturbed_member(g)[1] >= 350.0]
    for g in strict:
        r = per_member[g]
        assert r["max_dT"] <= TOL_T and r["max_dq"] <= TOL_Q and r["gmean"] <= TOL_GM, (g, r)
    assert all(r["gmean"] <= 1e-2 for r in per_member.values()), per_member
    assert len(strict) >= 10


def test_config2_fast_mode_vs_reference_fixture(forcing):
    """greb.original.model.f90 (log_exp = 10) control + scenario in the FAST mode against ref_config2.npz"""
    from test_gpu_golden import check_records
    g = np.load(os.path.join(GOLD, "ref_config2.npz"), allow_pickle=False)
    ens = greb_b200.Ensemble(1)
    ens.set_arithmetic("fast")
    ens.set_forcing(forcing)
    co2 = np.concatenate([np.full(3, 340.0), np.full(50, 680.0)]).astype(np.float32)
    ens.set_member(0, greb_b200.original_physics(), co2, year0=1970)
    ens.init()
    ens.spinup(3)
    ini = {n: ens.get_state(0, n) for n in ("Ts", "Ta", "To", "q")}
    ens.reset_scenario()
    ctrl, gmc, _ = ens.run(3)
    for n, a in ini.items():
        ens.set_state(0, n, a)
    scen, gm, _ = ens.run(50)
    mx = [check_records(ctrl[0, 0, 0], g["control_first_month"], "fast control, first month")]
    for y in (1, 10, 50):
        mx.append(check_records(scen[0, y - 1, 11], g[f"dec_year{y}"], f"fast scenario december of year {y}"))
    con = g["console"]
    dgm = float(np.abs(gm[0].astype(np.float64) - con[6:, 1]).max())
    dgc = float(np.abs(gmc[0].astype(np.float64) - con[3:6, 1]).max())
    assert dgm <= TOL_GM and dgc <= TOL_GM, (dgm, dgc)
    mx = np.array(mx).max(axis=0)
    _write_margins("config2_fast", {"max_dT_K": float(mx[:3].max()), "max_dq": float(mx[3]),
                                    "max_dgmean_console_K": max(dgm, dgc)})
    ens.close()
