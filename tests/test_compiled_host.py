"""The compiled host (greb-climate-model_b200/host/greb_main.cpp -> host/greb_host): PROGRAM greb_run of the
reference (src/greb.f90:996-1098) written against the C ABI only.  CPU: its namelist front-end equals the Python
twin (greb_b200/host.py) on the syntax gfortran accepts, the same inputs are errors in both, and without a GPU a
run fails loudly.  GPU: `greb_host nml_a nml_b` writes the same bytes and prints the same console lines as
host.run_namelists."""
import json
import os
import subprocess

import numpy as np
import pytest

from greb_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "greb-climate-model_b200", "host", "greb_host")

GOOD = {
    "defaults": "&physics_par\n/\n&numerics_par\n time_flux = 2, time_scnr = 3\n/\n&diagnostics_par\n/\n&co2_par\n/\n",
    "fortran_syntax": """! a comment with a 'quote
&PHYSICS_PAR
  kappa = 9.3d5, a_cloud = .33   ! trailing comment
  p_emi(3) = 5.0, 6.0
  Tl_ice1 = 262.5
/
&numerics_par
 time_flux = 1 time_scnr = 5
 ipx = 40, ipy=20, year0 = 1850
/
&diagnostics_par
 output_file = 'out/a=b', ens_id = "007"
/
&co2_par
 co2_ppm(2) = 400, 2*450
 co2_flux = 310.5
/
""",
    "array_constructor": "&physics_par\n p_emi = (/ 1., 2., 3. /)\n/\n&numerics_par\n time_scnr = 4 time_flux=0\n/\n"
                         "&diagnostics_par\n output_file = \"output/x(1) = y\"\n&end\n&co2_par\n co2_ppm = 340, 1e3\n/\n",
    "co2_default": "&numerics_par\n time_scnr = 2\n/\n",
    "repeat_all": "&numerics_par\n time_scnr = 3\n/\n&co2_par\n co2_ppm = 3*560.25\n/\n",
}
BAD = {
    "unterminated": "&physics_par\n kappa = 1e6\n",
    "unknown_variable": "&physics_par\n kapa = 1e6\n/\n",
    "unknown_group": "&physic_par\n/\n",
    "too_many_co2": "&numerics_par\n time_scnr = 1\n/\n&co2_par\n co2_ppm = 340, 350\n/\n",
    "subscripted_scalar": "&physics_par\n kappa(2) = 1e6\n/\n",
    "two_values_for_a_scalar": "&numerics_par\n time_scnr = 1, 2\n/\n",
    "group_twice": "&numerics_par\n/\n&numerics_par\n/\n",
}


@pytest.fixture(scope="module")
def binary():
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(BIN), "greb_host was not built (make -C greb-climate-model_b200/csrc)"
    return BIN


def _check(binary, path):
    return subprocess.run([binary, "--check", str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


@pytest.mark.parametrize("name", sorted(GOOD))
def test_check_mode_equals_the_python_front_end(binary, tmp_path, name):
    path = tmp_path / "namelist"
    path.write_text(GOOD[name])
    r = _check(binary, path)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    want = host.config_from_namelist(GOOD[name])
    for k in ("time_flux", "time_scnr", "year0", "ipx", "ipy"):
        assert got[k] == getattr(want, k), k
    assert got["output_file_full"] == want.output_file_full
    f32 = lambda x: np.float32(x)
    for k, v in got["physics"].items():
        if k == "p_emi":
            assert [f32(x) for x in v] == [f32(x) for x in want.physics.p_emi], k
        else:
            field = {n.lower(): n for n in host._lib.PHYS_FIELDS}.get(k, k)
            assert f32(v) == f32(getattr(want.physics, field)), (k, v, getattr(want.physics, field))
    assert [f32(x) for x in got["co2_ppm"]] == [f32(x) for x in want.co2_ppm]


@pytest.mark.parametrize("name", sorted(BAD))
def test_the_same_namelists_are_errors_in_both(binary, tmp_path, name):
    path = tmp_path / "namelist"
    path.write_text(BAD[name])
    r = _check(binary, path)
    assert r.returncode == 2 and r.stderr.startswith("greb: "), (r.returncode, r.stderr)
    with pytest.raises(host.NamelistError):
        host.config_from_namelist(BAD[name])


def test_default_namelist_name_and_missing_inputs(binary, tmp_path):
    """no argument = the file `namelist` (f:1030-1032); a missing namelist or input file is exit code 2"""
    r = subprocess.run([binary, "--check"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "namelist" in r.stderr
    (tmp_path / "namelist").write_text(GOOD["defaults"])
    r = subprocess.run([binary, "--check"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and json.loads(r.stdout)["time_scnr"] == 3
    r = subprocess.run([binary], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "input/topography" in r.stderr


def test_a_run_without_a_gpu_fails_loudly(binary, tmp_path, forcing):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    forcing.write(str(tmp_path / "input"))
    (tmp_path / "namelist").write_text(GOOD["co2_default"])
    r = subprocess.run([binary], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 3 and "greb_b200_create" in r.stderr, (r.returncode, r.stderr)
    assert not (tmp_path / "output").exists()


@pytest.mark.gpu
def test_compiled_host_writes_the_same_bytes_as_the_python_host(binary, tmp_path, forcing):
    forcing.write(str(tmp_path / "input"))
    nml = []
    for ens_id, co2, kappa in (("a", "680.0", 8e5), ("b", "400.0, 420.", 9.4e5)):
        path = tmp_path / f"namelist_{ens_id}"
        path.write_text(f"&PHYSICS_PAR\n kappa = {kappa}\n/\n&NUMERICS_PAR\n time_flux = 1\n time_scnr = 2\n ipx = 95\n"
                        f" ipy = 38\n/\n&DIAGNOSTICS_PAR\n output_file = \"output/scenario\"\n ens_id = \"{ens_id}\"\n/\n"
                        f"&CO2_PAR\n co2_ppm = {co2}\n/\n")
        nml.append(str(path))
    py_dir, c_dir = tmp_path / "py", tmp_path / "c"
    py_dir.mkdir()
    c_dir.mkdir()
    res = host.run_namelists(nml, input_dir=str(tmp_path / "input"), workdir=str(py_dir))
    r = subprocess.run([binary, "--input", str(tmp_path / "input")] + nml, cwd=c_dir, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    for ens_id in ("a", "b"):
        a = (py_dir / "output" / f"scenario_{ens_id}").read_bytes()
        b = (c_dir / "output" / f"scenario_{ens_id}").read_bytes()
        assert len(a) == 2 * 12 * 5 * 96 * 48 * 4 and a == b, ens_id
    lines = [ln for ln in r.stdout.splitlines() if not ln.lstrip().startswith("%")]
    want = [res[m]["lines"][y] for y in range(2) for m in range(2)]       # year-major, member-minor
    assert lines == want


# ---- greb-original mode (src/greb.original.model.f90 + shell) ----------------------------------------------
def _orig_namelist(tf, tc, ts, log_exp):
    return f"&NUMERICS\n time_flux = {tf}\n time_ctrl = {tc}\n time_scnr = {ts}\n/\n&PHYSICS\n log_exp = {log_exp}\n/\n"


@pytest.mark.parametrize("log_exp", list(range(1, 17)))
def test_original_mode_plans_the_same_experiment_as_the_python_host(binary, tmp_path, forcing, log_exp):
    path = tmp_path / "namelist_original"
    path.write_text(_orig_namelist(1, 2, 4, log_exp))
    r = subprocess.run([binary, "--original", "--check", str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    ex = host.original_experiment(log_exp, forcing, 4)
    assert (got["time_flux"], got["time_ctrl"], got["time_scnr"], got["log_exp"]) == (1, 2, 4, log_exp)
    assert got["switches"] == ex["switches"]
    assert np.float32(got["co2_ctrl"]) == np.float32(ex["co2_ctrl"])
    want = np.concatenate([np.full(2, ex["co2_ctrl"], dtype=np.float32), ex["co2_scenario"]])
    assert np.array_equal(np.asarray(got["co2"], dtype=np.float32), want)


def test_original_mode_rejects_an_unknown_experiment(binary, tmp_path):
    path = tmp_path / "namelist_original"
    path.write_text(_orig_namelist(1, 1, 1, 17))
    r = subprocess.run([binary, "--original", "--check", str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "log_exp" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("log_exp", [10, 5, 13])
def test_original_mode_writes_the_same_control_and_scenario_files(binary, tmp_path, forcing, log_exp):
    """full model, an experiment with modified inputs (mld = d_ocean) and the A1B pathway without hydrology"""
    forcing.write(str(tmp_path / "input"))
    path = tmp_path / "namelist_original"
    path.write_text(_orig_namelist(1, 1, 2, log_exp))
    py_dir, c_dir = tmp_path / "py", tmp_path / "c"
    py_dir.mkdir()
    c_dir.mkdir()
    host.run_original_namelist(str(path), input_dir=str(tmp_path / "input"), workdir=str(py_dir))
    r = subprocess.run([binary, "--original", "--input", str(tmp_path / "input"), str(path)], cwd=c_dir,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    for name, size in (("control", 730 * 96 * 48 * 4), ("scenario", 2 * 12 * 5 * 96 * 48 * 4)):
        a = (py_dir / "output" / name).read_bytes()
        b = (c_dir / "output" / name).read_bytes()
        assert len(a) == size and a == b, name


@pytest.mark.gpu
def test_compiled_host_shards_members_over_gpus(binary, tmp_path, forcing):
    """--gpus 2: three members in blocks over two GPUs (one handle and one host thread each) write the same files
    and print the same lines as one GPU — members never interact, there is nothing to exchange"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    forcing.write(str(tmp_path / "input"))
    nml = []
    for ens_id, co2, kappa in (("a", 680.0, 8e5), ("b", 400.0, 9.4e5), ("c", 900.0, 7e5)):
        path = tmp_path / f"namelist_{ens_id}"
        path.write_text(f"&PHYSICS_PAR\n kappa = {kappa}\n/\n&NUMERICS_PAR\n time_flux = 1\n time_scnr = 2\n/\n"
                        f"&DIAGNOSTICS_PAR\n ens_id = \"{ens_id}\"\n/\n&CO2_PAR\n co2_ppm = {co2}\n/\n")
        nml.append(str(path))
    outs = {}
    for g in (1, 2):
        d = tmp_path / f"g{g}"
        d.mkdir()
        r = subprocess.run([binary, "--gpus", str(g), "--input", str(tmp_path / "input")] + nml, cwd=d,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        outs[g] = (r.stdout, [(d / "output" / f"scenario_{e}").read_bytes() for e in "abc"])
    assert outs[1][0] == outs[2][0]
    assert outs[1][1] == outs[2][1] and all(len(b) == 2 * 12 * 5 * 96 * 48 * 4 for b in outs[2][1])
