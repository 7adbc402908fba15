"""The `log_exp` sensitivity experiments of src/greb.original.model.f90 as
process switches of the C ABI (GREB_SW_*), against golden vectors produced by the reference itself
(tests/golden/make_golden_experiments.py: the machine-translated greb.original.model.f90).

CPU: the kernel source in the lane emulator (same libm as the reference) must reproduce the
reference's records BIT FOR BIT.  GPU: through the ABI, within the BASELINE.json tolerances
(the device's expf/logf differ from glibc by an ulp)."""
import os

import numpy as np
import pytest

import greb_b200
from greb_b200 import host

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_original_experiments.npz")
TOL_T, TOL_Q, TOL_ALB = 1e-2, 1e-6, 1e-3        # K, kg/kg, albedo (BASELINE.json north_star gates)
TOL_GM = 1e-3                                    # K, console global mean


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_experiment_table_covers_every_defined_log_exp():
    assert sorted(host.ORIGINAL_EXPERIMENTS) == list(range(1, 17))
    for L in (0, 17):
        with pytest.raises(ValueError):
            host.original_experiment(L, None, 1)
    assert host.ORIGINAL_EXPERIMENTS[10] == 0 and host.ORIGINAL_EXPERIMENTS[5] == 7
    # orig:553-555: circulation returns before assigning its result -> defined as "no circulation" (dX_crcl = 0)
    both = greb_b200.lib.SW_NO_HEAT_CIRCULATION | greb_b200.lib.SW_NO_VAPOR_CIRCULATION
    assert all(host.ORIGINAL_EXPERIMENTS[L] & both == both for L in (1, 2, 3, 4))
    assert all(host.ORIGINAL_EXPERIMENTS[L] & both == greb_b200.lib.SW_NO_VAPOR_CIRCULATION for L in (7, 16))
    assert all(host.ORIGINAL_EXPERIMENTS[L] & both == 0 for L in (5, 6, 8, 9, 10, 11, 12, 13, 14, 15))


def test_a1b_pathway_matches_orig_co2_level():
    # greb.original.model.f90:945-951
    assert [float(host.a1b_co2(y)) for y in (1950, 2000, 2050, 2100, 2101)] == [310.0, 370.0, 520.0, 700.0, 680.0]
    assert float(host.a1b_co2(1940)) == 298.0 and float(host.a1b_co2(2025)) == 445.0


def test_experiment_inputs(forcing):
    ex = host.original_experiment(9, forcing, 3)
    assert np.all(ex["forcing"].mldclim == np.float32(50.0)) and ex["forcing"].tclim is forcing.tclim
    assert ex["co2_ctrl"] == 340.0 and list(ex["co2_scenario"]) == [680.0] * 3
    ex = host.original_experiment(14, forcing, 2)
    assert ex["forcing"].mldclim is forcing.mldclim and list(ex["co2_scenario"]) == [340.0, 340.0]
    ex = host.original_experiment(12, forcing, 2)
    assert ex["co2_ctrl"] == 298.0 and list(ex["co2_scenario"]) == [298.0, np.float32(299.2)]
    ex = host.original_experiment(1, forcing, 1)           # orig:162-165
    f1 = ex["forcing"]
    assert f1.z_topo.max() == 1.0 and np.array_equal(f1.z_topo[forcing.z_topo <= 1], forcing.z_topo[forcing.z_topo <= 1])
    assert np.all(f1.cldclim == np.float32(0.7)) and np.all(f1.qclim == np.float32(0.0052)) and np.all(f1.mldclim == 50.0)
    f3 = host.original_experiment(3, forcing, 1)["forcing"]
    assert f3.z_topo is forcing.z_topo and f3.cldclim is forcing.cldclim and np.all(f3.qclim == np.float32(0.0052))


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16])
def test_emulated_kernel_source_reproduces_the_reference_bit_for_bit(L, forcing, gold):
    """40 scenario steps of the reference's own time_loop (zero flux corrections) vs the kernel source."""
    import emu_lib
    n = int(gold["nsteps"])
    ex = host.original_experiment(L, forcing, 1)
    p = greb_b200.original_physics()
    p.co2_flux = ex["co2_ctrl"]
    e = emu_lib.Emu(ex["forcing"], p, ex["co2_scenario"][:1])
    e.set_switches(ex["switches"])
    e.steps(1, n)
    want = gold[f"steps_{L}_state"]
    for i, name in enumerate(("Ts", "Ta", "To", "q", "cap_surf")):
        assert np.array_equal(e.get(i), want[i]), (L, name, np.abs(e.get(i) - want[i]).max())


def _check(got, want, what):
    for v, tol in zip(range(5), (TOL_T, TOL_T, TOL_T, TOL_Q, TOL_ALB)):
        d = np.abs(got[v].astype(np.float64) - want[v]).max()
        assert d <= tol, (what, host.VARNAMES[v], d)


@pytest.mark.gpu
@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16])
def test_experiment_through_the_abi(L, forcing, gold):
    r = host.run_original(L, forcing, time_flux=1, time_ctrl=1, time_scnr=2)
    assert r["flags"].sum() == 0
    _check(r["scenario"][1, 11], gold[f"long_{L}_scen_dec2"], f"log_exp {L} scenario december of year 2")
    # console lines of the reference's diagnostics (greb.original.model.f90 diagnostics), one per simulated year: 1 flux-correction
    # year, 1 control year, 2 scenario years; column 1 is sum(tsmn)/(xdim*ydim)-273.15
    con = gold[f"long_{L}_console"]
    assert con.shape[0] == 4
    assert abs(float(r["gmean_control"][0]) - con[1, 1]) <= TOL_GM, (L, r["gmean_control"], con[1])
    assert np.abs(r["gmean"].astype(np.float64) - con[2:4, 1]).max() <= TOL_GM, (L, r["gmean"], con[2:4, 1])


@pytest.mark.gpu
def test_switch_rules(forcing):
    ens = greb_b200.Ensemble(3)
    ens.set_forcing(forcing)
    p = greb_b200.default_physics()
    for m in range(3):
        ens.set_member(m, p, [680.0, 680.0])
    with pytest.raises(greb_b200.GrebError):
        ens.set_switches(0, 256)                           # unknown bit
    ens.set_switches(1, greb_b200.lib.SW_NO_HYDRO)
    ens.init()
    with pytest.raises(greb_b200.GrebError):
        ens.set_switches(0, greb_b200.lib.SW_NO_HYDRO)     # would change the spin-up groups
    ens.set_switches(2, greb_b200.lib.SW_SST_PLUS_1K)      # scenario-only bit may be toggled
    ens.spinup(1)
    ens.reset_scenario()
    _, gm, _ = ens.run(1, want_output=False)
    # member 0 = full model; members 1 and 2 differ from it and from each other
    assert len({float(x) for x in gm[:, 0]}) == 3
    # same physics, different switches: never one spin-up group (corrections differ)
    assert not np.array_equal(ens.get_fluxcorr(0, 0), ens.get_fluxcorr(1, 0))
    assert np.array_equal(ens.get_fluxcorr(0, 0), ens.get_fluxcorr(2, 0))
    ens.close()


def test_write_control_layout(tmp_path):
    """orig:204-215: 730 TF_correct records, the control run's monthly records written over them."""
    tf = np.arange(730, dtype=np.float32)[:, None, None] * np.ones((48, 96), np.float32) + 1000
    mon = -np.arange(2 * 12 * 5, dtype=np.float32).reshape(2, 12, 5, 1, 1) * np.ones((48, 96), np.float32)
    host.write_control(str(tmp_path / "output" / "control"), tf, mon)
    a = np.fromfile(tmp_path / "output" / "control", dtype="<f4").reshape(-1, 48, 96)
    assert a.shape[0] == 730 and a[0, 0, 0] == 0 and a[119, 5, 7] == -119 and a[120, 0, 0] == 1120 and a[729, 0, 0] == 1729
    host.write_control(str(tmp_path / "output" / "control2"), tf, np.tile(mon, (7, 1, 1, 1, 1)))   # 14 years > 730 records
    assert os.path.getsize(tmp_path / "output" / "control2") == 14 * 60 * 48 * 96 * 4


@pytest.mark.gpu
def test_namelist_original_driver(forcing, tmp_path):
    forcing.write(str(tmp_path / "input"))
    nml = tmp_path / "namelist_original"
    nml.write_text("&NUMERICS\ntime_flux = 1  ! length of flux corrections run [yrs]\ntime_ctrl = 1\n"
                   "time_scnr = 2 ! length of scenariorun [yrs\n/\n&PHYSICS\n log_exp = 9 ! no deep ocean\n/\n")
    r = host.run_original_namelist(str(nml), input_dir=str(tmp_path / "input"), workdir=str(tmp_path))
    ctrl = np.fromfile(tmp_path / "output" / "control", dtype="<f4").reshape(-1, 48, 96)
    scen = host.read_greb(str(tmp_path / "output" / "scenario"))
    assert ctrl.shape[0] == 730 and scen["value"].shape == (24, 5, 48, 96)
    assert np.array_equal(ctrl[:60].reshape(12, 5, 48, 96), r["control"][0])
    assert np.array_equal(ctrl[60:], r["tf_correct"][60:])
    assert np.array_equal(scen["value"].reshape(2, 12, 5, 48, 96), r["scenario"])
    assert r["experiment"]["switches"] == greb_b200.lib.SW_NO_DEEP_OCEAN
