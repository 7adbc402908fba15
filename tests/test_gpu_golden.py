"""GPU parity against the golden vectors produced by the reference itself (tests/golden/*.npz,
written by tests/golden/make_golden.py from the machine-translated reference source).  No oracle in
the loop: CUDA path (through the C ABI) vs what the reference's own statements computed.

  * circulation(X_in, dX_crcl, h_scl, wz) — bit-exact (signs of zeros aside);
  * config 1 (reference `namelist`): December fields of years 1, 10, 50 and the yearly console value,
    within the north_star tolerances (the device libm differs from glibc in the last ulp);
  * config 2 (greb-original, log_exp = 10): control run + scenario through the ABI;
  * the batched namelist driver (greb_b200.host.run_namelists) writing reference-format files.
"""
import os

import numpy as np
import pytest

import greb_b200
from greb_b200 import host

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_T, TOL_Q, TOL_GM = 1e-2, 1e-6, 1e-3


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def same_bits(a, b):
    return np.array_equal(np.where(a == 0, np.float32(0), a).view(np.uint32),
                          np.where(b == 0, np.float32(0), b).view(np.uint32))


def check_records(got5, want5, label):
    d = np.abs(got5.astype(np.float64) - want5.astype(np.float64))
    mx = [float(d[v].max()) for v in range(5)]
    assert max(mx[:3]) <= TOL_T, (label, mx)
    assert mx[3] <= TOL_Q, (label, mx)
    assert mx[4] <= 1e-4, (label, mx)
    return mx


def test_circulation_bit_exact_vs_reference_vectors(forcing):
    g = load("ref_kernels.npz")
    assert str(g["forcing_digest"]) == forcing.digest()
    for case in range(int(g["n_cases"])):
        ityr, kappa = g[f"k{case}_meta"]
        p = greb_b200.default_physics()
        p.kappa = float(kappa)
        ens = greb_b200.Ensemble(1)
        ens.set_forcing(forcing)
        ens.set_member(0, p, [680.0])
        ens.init()
        X = np.stack([g[f"k{case}_T_in"], g[f"k{case}_q_in"]])
        W = np.stack([g[f"k{case}_T_wz"], g[f"k{case}_q_wz"]])
        got = ens.circulation(0, int(ityr), X, W)
        assert same_bits(got[0], g[f"k{case}_T_circulation"]), (case, "T")
        assert same_bits(got[1], g[f"k{case}_q_circulation"]), (case, "q")
        ens.close()


@pytest.mark.slow
def test_config1_vs_reference_run(forcing):
    g = load("ref_config1.npz")
    ens = greb_b200.Ensemble(1)
    ens.set_forcing(forcing)
    ens.set_member(0, greb_b200.default_physics(), np.full(50, 680.0, np.float32), year0=1940)
    ens.init()
    ens.spinup(3)
    ens.reset_scenario()
    out, gm, _ = ens.run(50)
    for y in (1, 10, 50):
        check_records(out[0, y - 1, 11], g[f"dec_year{y}"], f"december of year {y}")
    con = g["console"]
    scen = con[con[:, 0] >= 1940]
    assert np.abs(gm[0].astype(np.float64) - scen[:, 2]).max() <= TOL_GM          # f:954 global mean
    ens.close()


@pytest.mark.slow
def test_config2_original_control_and_scenario(forcing):
    """greb.original.model.f90:138-233 through the ABI: spin-up and control at 340 ppm, scenario at 680."""
    g = load("ref_config2.npz")
    ens = greb_b200.Ensemble(1)
    ens.set_forcing(forcing)
    p = greb_b200.original_physics()                       # cp_land = cp_ocean/4.5, co2_flux = CO2_ctrl = 340
    co2 = np.concatenate([np.full(3, 340.0), np.full(50, 680.0)]).astype(np.float32)
    ens.set_member(0, p, co2, year0=1970)
    ens.init()
    ens.spinup(3)                                          # orig:201
    tf = ens.get_fluxcorr(0, 0)                            # orig:204-206: 730 records of TF_correct
    ini = {n: ens.get_state(0, n) for n in ("Ts", "Ta", "To", "q")}
    ens.reset_scenario()
    ctrl, gmc, _ = ens.run(3)                              # orig:209-215 control run
    for n, a in ini.items():                               # orig:219: fields re-initialised, cap_surf is not
        ens.set_state(0, n, a)
    # the scenario restarts its calendar (it = 1, mon = 1, irec = 0) but continues on the same handle:
    # years 4..53 of the member's CO2 path are the 680 ppm years
    scen, gm, _ = ens.run(50)
    check_records(ctrl[0, 0, 0], g["control_first_month"], "control, first month")
    d181 = np.abs(tf[180].astype(np.float64) - g["control_rec_181"])
    assert d181.max() <= 5.0, d181.max()                   # TF_correct [W/m2]: cap_surf/dt * 1e-2 K ~ 2 W/m2
    for y in (1, 10, 50):
        check_records(scen[0, y - 1, 11], g[f"dec_year{y}"], f"scenario december of year {y}")
    con = g["console"]
    assert np.abs(gm[0].astype(np.float64) - con[6:, 1]).max() <= TOL_GM
    ens.close()


def test_run_namelists_batch_writes_reference_format(forcing, tmp_path, oracle_mod):
    forcing.write(str(tmp_path / "input"))
    nml = []
    specs = [("a", 680.0, 8e5), ("b", 400.0, 9.4e5)]
    for ens_id, co2, kappa in specs:
        path = tmp_path / f"namelist_{ens_id}"
        path.write_text(f"&PHYSICS_PAR\n kappa = {kappa}\n/\n&NUMERICS_PAR\n time_flux = 1\n time_scnr = 2\n ipx = 95\n"
                        f" ipy = 38\n/\n&DIAGNOSTICS_PAR\n output_file = \"output/scenario\"\n ens_id = \"{ens_id}\"\n/\n"
                        f"&CO2_PAR\n co2_ppm = {co2}\n/\n")
        nml.append(str(path))
    res = host.run_namelists(nml, input_dir=str(tmp_path / "input"), workdir=str(tmp_path))
    for (ens_id, co2, kappa), r in zip(specs, res):
        f = tmp_path / "output" / f"scenario_{ens_id}"
        assert os.path.getsize(f) == 2 * 12 * 5 * 96 * 48 * 4
        got = host.read_greb(str(f))["value"].reshape(2, 12, 5, 48, 96)
        o = oracle_mod.Oracle(forcing, kappa=kappa)
        o.spinup(1)
        want, gm = o.run(2, co2)
        for y in range(2):
            for m in range(12):
                check_records(got[y, m], want[y, m], f"{ens_id} y{y} m{m}")
        assert np.abs(r["gmean"] - gm).max() <= TOL_GM
        assert len(r["lines"]) == 2


def test_bistable_members_exact_mode_equals_the_reference_record_for_record(forcing):
    """The two low-CO2 members of the perturbed ensemble on which the fast arithmetic drifts (members 22 and
    2989, tests/test_gpu_long_parity.py): the default (exact) arithmetic reproduces what the translated reference
    wrote for them — every one of the 720 records of 3 + 12 years — digest for digest, no oracle in the loop."""
    import hashlib
    from greb_b200 import campaign
    g = load("ref_members.npz")
    assert str(g["forcing_digest"]) == forcing.digest()
    members = [int(m) for m in g["members"]]
    spinup, years = int(g["spinup"]), int(g["years"])
    ens = greb_b200.Ensemble(len(members))
    ens.set_arithmetic("exact")
    ens.set_forcing(forcing)
    for i, m in enumerate(members):
        p, co2 = campaign.perturbed_member(m)
        ens.set_member(i, p, np.full(years, co2, dtype=np.float32))
    ens.init()
    ens.spinup(spinup)
    ens.reset_scenario()
    out, gm, _ = ens.run(years)
    ens.close()
    for i, m in enumerate(members):
        recs = np.where(out[i] == 0, np.float32(0), out[i]).reshape(-1, 48, 96)      # signs of zeros aside
        dig = np.array([hashlib.sha1(np.ascontiguousarray(r, dtype="<f4").tobytes()).hexdigest()[:16] for r in recs])
        bad = np.nonzero(dig != g[f"m{m}_digests"])[0]
        assert bad.size == 0, (m, int(bad[0]))
        con = g[f"m{m}_console"]
        scen = con[con[:, 0] >= 1940]
        assert np.array_equal(scen[:, 2].astype(np.float32), gm[i]), m
