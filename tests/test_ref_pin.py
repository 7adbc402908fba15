"""The oracle against the translated reference library itself (oracle/_ref/, built by
oracle/ref.py from /root/reference/src/*.f90 when that tree is present).  Skipped where neither the
reference tree nor a prebuilt oracle/_ref/libgreb_ref.so exists; tests/test_golden.py carries the
same pin everywhere through committed fixtures."""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available("greb"), reason="no reference tree and no prebuilt oracle/_ref")


@pytest.fixture(scope="module")
def R(forcing):
    r = ref.Ref.fresh("greb")
    r.set_forcing(forcing)
    return r


def test_translator_reproduces_reference_constants(R):
    # module initialisers are evaluated in fp32 like gfortran's constant folder
    assert R.geti("nstep_yr") == 730 and R.geti("dt") == 43200 and R.geti("dt_crcl") == 1800
    assert np.float32(R.get("to_ice2")) == np.float32(np.float32(273.15) - np.float32(1.7))
    assert np.float32(R.get("cq_rain")) == np.float32(np.float32(np.float32(-0.1) / np.float32(24.0)) / np.float32(3600.0))
    assert np.float32(R.get("dlon")) == np.float32(3.75)


def test_column_physics_routines_bit_exact(R, oracle_mod, forcing):
    """SWradiation, LWradiation, hydro, seaice, deep_ocean with the reference's own argument lists."""
    o = oracle_mod.Oracle(forcing)
    rng = np.random.default_rng(3)
    R.array("dtrad", (730, 48, 96))[:] = (np.float32(-0.16) * forcing.tclim - np.float32(5.0)).astype(np.float32)
    R.array("z_ocean", (48, 96))[:] = o.derived("z_ocean")
    R.set("cap_ocean", float(np.float32(4186.0) * np.float32(999.1)))
    R.set("cap_land", float(np.float32(np.float32(926.222) * np.float32(2600.0)) * np.float32(2.0)))
    z = lambda: np.zeros((48, 96), np.float32)
    for ityr in (1, 2, 181, 730):
        Ts = (forcing.tclim[ityr - 1] + rng.uniform(-6, 6, (48, 96))).astype(np.float32)
        Ta = (Ts + rng.uniform(-3, 3, (48, 96))).astype(np.float32)
        To = (Ts - rng.uniform(0, 5, (48, 96))).astype(np.float32)
        q = (forcing.qclim[ityr - 1] * rng.uniform(0.5, 1.5, (48, 96))).astype(np.float32)
        R.seti("ityr", ityr)
        sw, alb = z(), z()
        R.call("swradiation", Ts, sw, alb)
        osw, oalb = o.SWradiation(Ts, ityr)
        assert np.array_equal(sw, osw) and np.array_equal(alb, oalb)
        lws, lwu, lwd, em = z(), z(), z(), z()
        R.call("lwradiation", Ts, Ta, q, 680.0, lws, lwu, lwd, em)
        for a, b in zip((lws, lwu, lwd, em), o.LWradiation(Ts, Ta, q, 680.0, ityr)):
            assert np.array_equal(a, b)
        ql, qla, dqe, dqr = z(), z(), z(), z()
        R.call("hydro", Ts, q, ql, qla, dqe, dqr)
        for a, b in zip((ql, qla, dqe, dqr), o.hydro(Ts, q, ityr)):
            assert np.array_equal(a, b)
        dTo_, dT = z(), z()
        R.call("deep_ocean", Ts, To, dT, dTo_)
        odT, odTo = o.deep_ocean(Ts, To, ityr)
        assert np.array_equal(dT, odT) and np.array_equal(dTo_, odTo)
        cap0 = rng.uniform(1e6, 1e8, (48, 96)).astype(np.float32)
        R.array("cap_surf", (48, 96))[:] = cap0
        o.set("cap_surf", cap0)
        R.call("seaice", Ts)
        assert np.array_equal(R.array("cap_surf", (48, 96)), o.seaice(Ts, ityr))


def test_short_default_run_bit_exact(oracle_mod, forcing):
    """greb_model (f:161-236) end to end: 1-yr flux correction + 2 yr, all 120 records + console."""
    r = ref.Ref.fresh("greb")
    r.set_forcing(forcing)
    r.set_physics(kappa=7.1e5, a_cloud=0.37)
    r.set_run(1, 2, [560.0], year0=1940)
    out = r.greb_model()
    o = oracle_mod.Oracle(forcing, kappa=7.1e5, a_cloud=0.37)
    o.spinup(1)
    oout, gm = o.run(2, 560.0)
    assert np.array_equal(out, oout.reshape(-1, 48, 96))
    con = [ln for ln in r.console() if len(ln) == 4 and ln[0] >= 1940]
    assert np.array_equal(np.array([ln[2] for ln in con], dtype=np.float32), gm)
    for name, which in (("tf_correct", 0), ("qf_correct", 1), ("tof_correct", 2)):
        assert np.array_equal(r.array(name, (730, 48, 96)), o.fluxcorr(which)), name
