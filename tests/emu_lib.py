"""ctypes binding of the TEST-ONLY lane emulator (tests/emu/greb_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "greb-climate-model_b200", "csrc")
LIB = os.path.join(HERE, "emu", "libgreb_emu.so")
SRCS = [os.path.join(HERE, "emu", "greb_emu.cpp"), os.path.join(CSRC, "greb_setup.cpp")]
DEPS = SRCS + [os.path.join(CSRC, f) for f in ("greb_core.h", "greb_simt.h", "greb_types.h", "greb_setup.h")]
XD, YD, NT = 96, 48, 730
NC = XD * YD
fp = C.POINTER(C.c_float)


def build():
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in DEPS):
        return
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DGREB_EMU", "-I" + CSRC,
           "-I" + os.path.join(ROOT, "include")] + SRCS + ["-o", LIB, "-lpthread", "-lm"]
    subprocess.run(cmd, check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.emu_circulation.argtypes = [C.c_void_p, fp, fp, fp, fp, fp]
        L.emu_row_tables.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [fp] * 10 + [C.c_void_p, fp, C.c_int]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_steps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, fp, C.c_int]
        L.emu_reset_scenario.argtypes = [C.c_void_p]
        L.emu_set_switches.argtypes = [C.c_void_p, C.c_uint]
        L.emu_get_state.argtypes = [C.c_void_p, C.c_int, fp]
        L.emu_set_state.argtypes = [C.c_void_p, C.c_int, fp]
        L.emu_get_corr.argtypes = [C.c_void_p, fp]
        L.emu_get_diag.argtypes = [C.c_void_p, fp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(fp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def circulation(phys, u, v, X, wz):
    u, v, X, wz = _f(u), _f(v), _f(X), _f(wz)
    out = np.zeros((YD, XD), dtype=np.float32)
    lib().emu_circulation(C.byref(phys), _p(u), _p(v), _p(X), _p(wz), _p(out))
    return out


def row_tables(phys):
    """(status, row_of_group[48], hslot_of_row[48]) of the host-side row assignment"""
    rg = (C.c_int * 48)()
    hs = (C.c_int * 48)()
    rc = lib().emu_row_tables(C.byref(phys), rg, hs)
    return rc, list(rg), list(hs)


class Emu:
    def __init__(self, forcing, phys, co2):
        f = forcing
        self.co2 = _f(co2)
        self._keep = [_f(a) for a in (f.z_topo, f.glacier, f.sw_solar, f.tclim, f.qclim, f.swetclim, f.uclim,
                                      f.vclim, f.mldclim, f.cldclim)]
        self.h = lib().emu_create(*[_p(a) for a in self._keep], C.byref(phys), _p(self.co2), len(self.co2))

    def __del__(self):
        try:
            lib().emu_destroy(self.h)
        except Exception:
            pass

    def steps(self, it0, nsteps, spinup=False, out_months=0):
        out = np.zeros((max(out_months, 1), 5, YD, XD), dtype=np.float32)
        lib().emu_steps(self.h, it0, nsteps, int(spinup), _p(out) if out_months else None, out_months)
        return out[:out_months]

    def reset_scenario(self):
        lib().emu_reset_scenario(self.h)

    def set_switches(self, mask):
        lib().emu_set_switches(self.h, mask)

    def get(self, which):
        a = np.zeros((YD, XD), dtype=np.float32)
        lib().emu_get_state(self.h, which, _p(a))
        return a

    def set(self, which, a):
        a = _f(a)
        lib().emu_set_state(self.h, which, _p(a))

    def corr(self):
        a = np.zeros((NT, 3, YD, XD), dtype=np.float32)
        lib().emu_get_corr(self.h, _p(a))
        return a

    def diag(self):
        a = np.zeros(2, dtype=np.float32)
        lib().emu_get_diag(self.h, _p(a))
        return a
