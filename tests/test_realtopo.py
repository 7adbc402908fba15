"""The REAL orography.

tests/golden/ref_realtopo.npz carries the three input fields the reference mount does have —
input/topography, input/glacier.masks, input/solar.radiation — and what the reference itself (src/greb.f90,
machine-translated and compiled: tests/golden/make_golden_realtopo.py) computed on them for 1 flux-correction
year + 2 scenario years at 680 ppm, the seven missing climatologies generated around that orography.

  * CPU: the C oracle reproduces every output record, the console values and the flux corrections bit for
    bit; the kernel source in the lane emulator reproduces the oracle's steps bit for bit; where
    /root/reference exists the fixture's inputs are compared with the files themselves.
  * GPU: the CUDA path in the exact mode reproduces all 120 records, bit for bit."""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_realtopo.npz")
REF_INPUT = "/root/reference/input"


def digests(recs):
    return np.array([hashlib.sha1(np.ascontiguousarray(r, dtype="<f4").tobytes()).hexdigest()[:16] for r in recs])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=False)


@pytest.fixture(scope="module")
def forcing_real(gold):
    from greb_b200 import synth
    f = synth.make_forcing(topo="reference", reference_fields=(gold["z_topo"], gold["glacier"], gold["sw_solar"]))
    assert f.digest() == str(gold["forcing_digest"])          # the generator is bit-reproducible on this host
    return f


def test_fixture_inputs_are_the_reference_files(gold):
    if not os.path.isdir(REF_INPUT):
        pytest.skip("no reference mount here")
    for name, key, shape in (("topography", "z_topo", (48, 96)), ("glacier.masks", "glacier", (48, 96)),
                             ("solar.radiation", "sw_solar", (730, 48))):
        a = np.fromfile(os.path.join(REF_INPUT, name), dtype="<f4").reshape(shape)
        assert np.array_equal(a, gold[key]), name
    z = gold["z_topo"]
    assert (z > 0).mean() > 0.2 and (z < 0).mean() > 0.5 and z.max() > 4000     # continents, oceans, Himalaya


def test_oracle_reproduces_the_reference_on_the_real_orography(gold, forcing_real, oracle_mod):
    o = oracle_mod.Oracle(forcing_real)
    o.spinup(1)
    for w, name in enumerate(("tf_correct", "qf_correct", "tof_correct")):
        assert np.array_equal(digests(o.fluxcorr(w)[::73]), gold[name + "_digest"]), name
    out, gm = o.run(2, co2_ppm=680.0)
    assert np.array_equal(digests(out.reshape(120, 48, 96)), gold["digests"])
    assert np.array_equal(out[1, 11], gold["dec_year2"])
    con = gold["console"]
    assert np.array_equal(gm.astype(np.float64), con[1:, 2])


def test_kernel_source_in_the_lane_emulator_on_the_real_orography(forcing_real, oracle_mod):
    """40 time_loop steps of the warp-level kernel source (CPU lane emulator, glibc libm) against the oracle
    from the same initial state, all five state fields bit for bit (real coast lines, glaciers, mountains)"""
    import emu_lib
    import greb_b200
    e = emu_lib.Emu(forcing_real, greb_b200.default_physics(), [680.0])
    o = oracle_mod.Oracle(forcing_real)
    e.steps(1, 40)
    for it in range(1, 41):
        o.time_loop(it, 680.0)
    for i, name in enumerate(("Ts", "Ta", "To", "q", "cap_surf")):
        a, b = e.get(i), o.get(name)
        assert np.array_equal(np.where(a == 0, np.float32(0), a), np.where(b == 0, np.float32(0), b)), \
            (name, float(np.abs(a - b).max()))


@pytest.mark.gpu
def test_cuda_exact_mode_reproduces_the_reference_on_the_real_orography(gold, forcing_real):
    import greb_b200
    ens = greb_b200.Ensemble(1)
    ens.set_forcing(forcing_real)
    ens.set_member(0, greb_b200.default_physics(), np.full(2, 680.0, dtype=np.float32))
    ens.init()
    ens.spinup(1)
    for w, name in enumerate(("tf_correct", "qf_correct", "tof_correct")):
        assert np.array_equal(digests(ens.get_fluxcorr(0, w)[::73]), gold[name + "_digest"]), name
    ens.reset_scenario()
    out, gm, _ = ens.run(2)
    got = digests(out[0].reshape(120, 48, 96))
    bad = np.nonzero(got != gold["digests"])[0]
    assert bad.size == 0, (bad[:10], float(np.abs(out[0, 1, 11] - gold["dec_year2"]).max()))
    assert np.array_equal(gm[0].astype(np.float64), gold["console"][1:, 2])
    ens.close()
