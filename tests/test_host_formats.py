"""Host front-end formats: namelists, co2 padding, ens_id naming, the output
record stream and the read_greb port.  CPU only."""
import os

import numpy as np
import pytest

from greb_b200 import host, lib

# the reference's default namelist (reference `namelist:1-14`) — configuration text, restated
REFERENCE_NAMELIST = """
&PHYSICS_PAR
/
&NUMERICS_PAR
ipx = 95 ! diagnostic output point, x-coord
ipy = 38 ! diagnostic output point, y-coord
time_flux = 3  ! length of flux corrections run [yrs]
time_scnr = 50 ! length of scenariorun [yrs]
/
&DIAGNOSTICS_PAR
output_file = "output/scenario"
/
&CO2_PAR
co2_ppm = 680
/
"""


def test_reference_default_namelist():
    c = host.config_from_namelist(REFERENCE_NAMELIST)
    assert (c.ipx, c.ipy, c.time_flux, c.time_scnr, c.year0) == (95, 38, 3, 50, 1940)
    assert c.output_file == "output/scenario" and c.ens_id == "" and c.output_file_full == "output/scenario"
    assert c.co2_ppm.shape == (50,) and np.all(c.co2_ppm == 680.0)
    assert np.float32(c.physics.kappa) == np.float32(8e5) and np.float32(c.physics.co2_flux) == np.float32(298.0)


def test_defaults_match_the_library_defaults():
    try:
        want = lib.default_physics()
    except Exception as e:  # pragma: no cover - library not built
        pytest.skip(str(e))
    got = host._default_physics_struct()
    for n in lib.PHYS_FIELDS + ["co2_flux"]:
        assert np.float32(getattr(got, n)) == np.float32(getattr(want, n)), n
    assert list(got.p_emi) == list(want.p_emi)


def test_namelist_syntax_variants():
    txt = """
    &physics_par  KAPPA = 9.4E5, a_cloud=0.33 ! trailing comment
       p_emi = (/9.0721, 106.7252, 61.5562, 0.0179, 0.0028, 0.0570, 0.3462, 2.3406, 0.7032, 1.0662/)
       cq_latent = 2.257d6 /
    &NUMERICS_PAR time_flux=1 time_scnr = 5
      year0 = 2000 /
    &diagnostics_par output_file = 'out/run', ens_id = "m07" /
    &co2_par co2_flux = 300, co2_ppm = 400, 2*500, 600 /
    """
    c = host.config_from_namelist(txt)
    assert np.float32(c.physics.kappa) == np.float32(9.4e5) and np.float32(c.physics.a_cloud) == np.float32(0.33)
    assert np.float32(c.physics.cq_latent) == np.float32(2.257e6)
    assert (c.time_flux, c.time_scnr, c.year0) == (1, 5, 2000)
    assert c.output_file_full == "out/run_m07"                                   # f:1063-1068
    assert c.co2_ppm.tolist() == [400.0, 500.0, 500.0, 600.0, 600.0]             # f:1053-1061 padding
    assert np.float32(c.physics.co2_flux) == np.float32(300.0)


def test_namelist_subscripts_and_quoted_equals():
    """gfortran accepts `a(i) = v`: the subscript must be honoured, not dropped (f:1047-1061 padding then
    sees -1 in the unassigned elements); a '=' or 'x(2) =' inside a quoted string is just text."""
    txt = """
    &physics_par p_emi(3) = 5.0, 6.0  kappa = 7e5 /
    &numerics_par time_scnr = 4 /
    &diagnostics_par ens_id = "a=b, co2_ppm(2) = 9" output_file = 'o/x=y' /
    &co2_par co2_ppm(2) = 400 /
    """
    c = host.config_from_namelist(txt)
    d = host._default_physics_struct()
    want = list(d.p_emi)
    want[2], want[3] = np.float32(5.0), np.float32(6.0)
    assert [np.float32(x) for x in c.physics.p_emi] == [np.float32(x) for x in want]
    assert c.co2_ppm.tolist() == [680.0, 400.0, 400.0, 400.0]          # reference: [-1,400,-1,-1] -> padded
    assert c.ens_id == "a=b, co2_ppm(2) = 9" and c.output_file == "o/x=y"
    c = host.config_from_namelist("&numerics_par time_scnr = 3 / &co2_par co2_ppm(1) = 300 co2_ppm(3) = 500 /")
    assert c.co2_ppm.tolist() == [300.0, 300.0, 300.0]                 # f:1056-1059: the first hole ends the scan
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&physics_par kappa(2) = 3 /")       # scalar with a subscript
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&physics_par kappa = 3, 4 /")
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&physics_par p_emi(0) = 3 /")


def test_co2_padding_rules():
    assert host.pad_co2([], 3).tolist() == [680.0, 680.0, 680.0]                 # first < 0 -> 680
    assert host.pad_co2([350.0], 4).tolist() == [350.0] * 4
    assert host.pad_co2([350.0, 360.0], 2).tolist() == [350.0, 360.0]
    assert host.pad_co2([1.0, 2.0, 3.0], 0).shape == (0,)
    try:
        assert np.array_equal(host.pad_co2([400.0, 500.0], 5), lib.pad_co2([400.0, 500.0], 5))   # C ABI twin
    except lib.GrebError:  # pragma: no cover - library not built
        pass


def test_namelist_errors():
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&physics_par kapa = 1 /")
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&numerics_par time_scnr = 2 / &co2_par co2_ppm = 1, 2, 3 /")
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&numerics_par time_scnr = 2")
    with pytest.raises(host.NamelistError):
        host.config_from_namelist("&foo a = 1 /")


def test_output_stream_and_read_greb_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    monthly = rng.normal(280, 10, (2, 12, 5, 48, 96)).astype(np.float32)
    path = tmp_path / "output" / "scenario_x"
    host.write_output(str(path), monthly)
    assert os.path.getsize(path) == 2 * 12 * 5 * 96 * 48 * 4                     # R/functions.R:41
    raw = np.fromfile(path, dtype="<f4")
    # record r (1-based) = ((year*12 + month)*5 + var); lon fastest, then lat (f:978-982)
    assert raw[(((1 * 12 + 3) * 5 + 2) * 48 + 7) * 96 + 11] == monthly[1, 3, 2, 7, 11]
    g = host.read_greb(str(path))
    assert g["value"].shape == (24, 5, 48, 96) and g["variable"] == list(host.VARNAMES)
    assert np.array_equal(g["value"].reshape(2, 12, 5, 48, 96), monthly)
    assert g["lon"][0] == 1.875 and g["lon"][-1] == 358.125 and g["lat"][0] == -88.125 and g["lat"][-1] == 88.125
    one = host.read_greb(str(path), varname=["albedo"], ivar=[5], nvar=5)
    assert np.array_equal(one["value"][:, 0], monthly.reshape(24, 5, 48, 96)[:, 4])
    with pytest.raises(ValueError):
        host.read_greb(str(path), nvar=7)
    # a shorter rerun does not truncate the file (the reference opens without status='replace', f:174)
    host.write_output(str(path), monthly[:1])
    assert os.path.getsize(path) == 2 * 12 * 5 * 96 * 48 * 4


def test_read_inputs_reports_missing_files(tmp_path):
    with pytest.raises(FileNotFoundError):
        host.read_inputs(str(tmp_path))


class _CacheEns:
    """host-memory stand-in with the accessor names of greb_b200.Ensemble (the disk format is what is tested)"""

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.state = {n: rng.normal(280, 5, (48, 96)).astype(np.float32) for n in lib.STATE}
        self.corr = [rng.normal(0, 30, (730, 48, 96)).astype(np.float32) for _ in range(3)]

    def get_state(self, m, n): return self.state[n]
    def get_fluxcorr(self, m, w): return self.corr[w]
    def set_state(self, m, n, a): self.state[n] = np.array(a, copy=True)
    def set_fluxcorr(self, m, w, a): self.corr[w] = np.array(a, copy=True)


def test_spinup_cache_roundtrip_and_key(forcing, tmp_path):
    p = lib.default_physics()
    k1 = host.spinup_key(forcing, p, 3)
    assert k1 == host.spinup_key(forcing, p, 3, switches=lib.SW_SST_PLUS_1K)   # scenario-only switch
    p2 = lib.default_physics()
    p2.kappa = 9e5
    keys = {k1, host.spinup_key(forcing, p2, 3), host.spinup_key(forcing, p, 2),
            host.spinup_key(forcing, p, 3, switches=lib.SW_NO_HYDRO), host.spinup_key(forcing, p, 3, arith="fast")}
    assert len(keys) == 5
    a, b = _CacheEns(1), _CacheEns(2)
    path = str(tmp_path / "cache" / f"spinup_{k1}.npz")
    assert not host.load_spinup(path, b, [0], k1)                      # nothing there yet
    host.save_spinup(path, a, 0, k1)
    assert not host.load_spinup(path, b, [0], "another key") and not np.array_equal(b.corr[0], a.corr[0])
    assert host.load_spinup(path, b, [0, 1], k1)
    assert all(np.array_equal(a.corr[w], b.corr[w]) for w in range(3))
    assert all(np.array_equal(a.state[n], b.state[n]) for n in a.state)


class _CkptEns:
    """host.save_checkpoint / load_checkpoint only use these methods of lib.Ensemble"""
    def __init__(self, n, seed):
        rng = np.random.default_rng(seed)
        self.n = n
        self.state = rng.normal(size=(n, 5, 48, 96)).astype(np.float32)
        self.acc = rng.normal(size=(n, 6, 48, 96)).astype(np.float32)
        self.corr = rng.normal(size=(3, n, 4, 48, 96)).astype(np.float32)   # 4 steps stand in for 730
        self.it = int(rng.integers(1, 5000))
    def get_calendar(self): return self.it
    def set_calendar(self, it): self.it = it
    def get_states(self): return self.state.copy()
    def set_states(self, a): self.state = np.array(a, copy=True)
    def get_accumulators(self): return self.acc.copy()
    def set_accumulators(self, a): self.acc = np.array(a, copy=True)
    def get_fluxcorr(self, m, w): return self.corr[w, m]
    def set_fluxcorr(self, m, w, a): self.corr[w, m] = a


def test_checkpoint_roundtrip(tmp_path):
    a, b = _CkptEns(3, 1), _CkptEns(3, 2)
    path = str(tmp_path / "ck" / "run.npz")
    host.save_checkpoint(path, a)
    assert host.load_checkpoint(path, b) == a.it and b.it == a.it
    assert np.array_equal(a.state, b.state) and np.array_equal(a.acc, b.acc) and np.array_equal(a.corr, b.corr)
    c = _CkptEns(3, 3)
    keep = c.corr.copy()
    host.save_checkpoint(path, a, with_fluxcorr=False)
    host.load_checkpoint(path, c)
    assert np.array_equal(c.state, a.state) and np.array_equal(c.corr, keep)
    with pytest.raises(ValueError):
        host.load_checkpoint(path, _CkptEns(2, 4))
