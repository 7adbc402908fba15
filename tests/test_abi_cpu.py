"""CPU checks of the drop-in boundary: the library builds, loads, exports every symbol that
include/greb_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import greb_b200
from greb_b200 import lib as gl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "greb-climate-model_b200")


def test_header_symbols_are_exported():
    if not os.path.exists(greb_b200.library_path()):
        greb_b200.build_library()
    L = greb_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "greb_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(greb_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/greb_b200.h but not exported"
    assert sorted(gl.ABI_SYMBOLS) == declared


def test_physics_defaults_match_reference_values():
    p = greb_b200.default_physics()
    assert np.float32(p.pi) == np.float32(3.1416) and np.float32(p.kappa) == np.float32(8e5)
    assert np.float32(p.To_ice2) == np.float32(273.15) - np.float32(1.7)
    assert np.float32(p.cq_rain) == np.float32(-0.1) / np.float32(24.) / np.float32(3600.)
    assert p.co2_flux == 298.0 and abs(p.p_emi[1] - 106.7252) < 1e-4
    o = greb_b200.original_physics()
    assert np.float32(o.cp_land) == np.float32(4186.) / np.float32(4.5) and o.co2_flux == 340.0


def test_physics_struct_layout_matches_oracle(oracle_mod):
    """Same field order in the product ABI struct and in the oracle's struct."""
    assert [f[0] for f in gl.Physics._fields_] == [f[0] for f in oracle_mod.Physics._fields_]
    assert C.sizeof(gl.Physics) == C.sizeof(oracle_mod.Physics) == 39 * 4
    a, b = greb_b200.default_physics(), oracle_mod.default_physics()
    assert bytes(a) == bytes(b)


def test_co2_padding():
    # reference src/greb.f90:1047-1061
    assert list(greb_b200.pad_co2([], 3)) == [680.0, 680.0, 680.0]
    assert list(greb_b200.pad_co2([340, 350], 5)) == [340.0, 350.0, 350.0, 350.0, 350.0]
    assert list(greb_b200.pad_co2([300, 310, 320, 330], 2)) == [300.0, 310.0]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(greb_b200.GrebError, match="no usable CUDA device|no CPU fallback"):
        greb_b200.Ensemble(1)


def test_product_sources_do_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may use it."""
    bad = []
    for base in (os.path.join(ROOT, "greb-climate-model_b200"), os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            for fn in files:
                if fn.endswith((".py", ".h", ".cu", ".cpp", ".f90", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"greb_oracle|from oracle|import oracle|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_headers_are_plain_c99_and_link_from_c(tmp_path):
    """The boundary is a C ABI: both headers compile as pedantic C99, and a C program linked against the library
    calls the host-side entries (defaults, CO2 padding) and gets an error code — not a crash — from
    greb_b200_create when there is no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include "greb_b200.h"
#include "greb_grid.h"
int main(void) {
  greb_physics_par p;
  float co2[4];
  const float given[2] = {-1.0f, 400.0f};
  greb_b200_t h = 0;
  int rc;
  greb_b200_physics_defaults(&p);
  greb_b200_pad_co2(given, 2, co2, 4);
  rc = greb_b200_create(&h, 1, 0);
  printf("%g %g %g %g %g %d %s\n", (double)p.kappa, (double)co2[0], (double)co2[1], (double)co2[2], (double)co2[3], rc,
         rc == GREB_OK ? "ok" : greb_b200_last_error(0));
  if (h) greb_b200_destroy(h);
  return 0;
}
''')
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-L", PKG_DIR, "-lgreb_b200", f"-Wl,-rpath,{PKG_DIR}", "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    f = r.stdout.split(None, 6)
    assert [float(x) for x in f[:5]] == [8e5, 680.0, 400.0, 400.0, 400.0]
    import torch
    if torch.cuda.is_available():
        assert int(f[5]) == 0
    else:
        assert int(f[5]) == -2 and len(f[6].strip()) > 0          # GREB_E_NO_DEVICE with a message
