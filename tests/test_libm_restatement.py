"""The exact mode's expf / logf = glibc's algorithm restated (csrc/greb_simt.h greb_expf_glibc / greb_logf_glibc).

CPU: tools/glibc_libm_check.c restates the algorithm in C and compares it with the host libm (sampled here,
exhaustive when run by hand).  GPU: the device functions, through the C ABI, against the host libm bit for bit —
and with them a WHOLE run (spin-up, scenario, monthly means, flux corrections) against the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_restatement_equals_host_libm(tmp_path):
    exe = str(tmp_path / "glibc_libm_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe, os.path.join(ROOT, "tools", "glibc_libm_check.c"),
                    "-lm", "-lpthread"], check=True)
    r = subprocess.run([exe, "61"], stdout=subprocess.PIPE, text=True)          # every 61st float of all 2^32
    assert r.returncode == 0, r.stdout
    assert "logf mismatches unfused 0 fused 0" in r.stdout and "(|x| <= 32: 0)" in r.stdout, r.stdout


def _args():
    rng = np.random.default_rng(3)
    ex = np.concatenate([rng.uniform(-32, 32, 400000), rng.uniform(-20, 6, 1200000), rng.normal(0, 1e-3, 100000),
                         np.array([0.0, -0.0, 32.0, -32.0, 1e-30, 88.0, -88.0, 100.0, -110.0, np.inf, -np.inf])])
    lg = np.concatenate([rng.uniform(0.5, 2000.0, 1200000), np.exp(rng.uniform(-80, 80, 400000)),
                         1.0 + rng.normal(0, 1e-4, 100000), np.array([1.0, 0.5, 2.0, 1e-38, 3e38, np.inf])])
    return ex.astype(np.float32), lg.astype(np.float32)


@pytest.mark.gpu
def test_device_expf_logf_equal_glibc_bit_for_bit(oracle_mod, forcing):
    import greb_b200
    ens = greb_b200.Ensemble(1)
    ex, lg = _args()
    for which, x in (("exp", ex), ("log", lg)):
        got = ens.device_libm(which, x)
        want = oracle_mod.host_libm(which, x)
        bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
        assert bad.size == 0, (which, bad.size, x[bad[:5]], got[bad[:5]], want[bad[:5]])
    ens.close()


@pytest.mark.gpu
def test_whole_run_is_bit_identical_in_the_exact_mode(oracle_mod, forcing):
    """spin-up 1 yr + 3 scenario years, a perturbed member at low CO2 (sea ice, the sensitive regime): every
    monthly-mean record, the flux corrections, the console value and the end state equal the oracle's bits"""
    import greb_b200
    from greb_b200 import campaign
    from test_gpu_parity import same_bits
    ens = greb_b200.Ensemble(2)
    ens.set_forcing(forcing)
    specs = [campaign.perturbed_member(22), (greb_b200.default_physics(), 680.0)]
    for m, (p, co2) in enumerate(specs):
        ens.set_member(m, p, np.full(3, co2, dtype=np.float32))
    ens.init()
    ens.spinup(1)
    corr = [[ens.get_fluxcorr(m, w) for w in range(3)] for m in range(2)]
    ens.reset_scenario()
    out, gm, _ = ens.run(3)
    for m, (p, co2) in enumerate(specs):
        o = oracle_mod.Oracle(forcing, **{k: getattr(p, k) for k in ("kappa", "ct_sens", "ce", "co_turb", "a_cloud", "da_ice")})
        o.spinup(1)
        for w in range(3):
            assert same_bits(corr[m][w], o.fluxcorr(w)), (m, "flux correction", w)
        out_o, gm_o = o.run(3, co2_ppm=co2)
        assert same_bits(out[m], out_o), (m, float(np.abs(out[m] - out_o).max()))
        assert np.array_equal(gm[m], gm_o), (m, gm[m], gm_o)
        for name in ("Ts", "Ta", "To", "q", "cap_surf"):
            assert same_bits(ens.get_state(m, name), o.get(name)), (m, name)
    ens.close()
