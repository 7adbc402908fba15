"""ctypes binding of the CPU oracle (oracle/greb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under greb-climate-model_b200/ may import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgreb_oracle.so")
CLI = os.path.join(HERE, "greb_oracle")
XD, YD, NT = 96, 48, 730
NC = XD * YD

PHYS_FIELDS = ["pi", "sig", "rho_ocean", "rho_land", "rho_air", "cp_ocean", "cp_land", "cp_air", "eps",
               "d_ocean", "d_land", "d_air", "ct_sens", "da_ice", "a_no_ice", "a_cloud", "Tl_ice1",
               "Tl_ice2", "To_ice1", "To_ice2", "co_turb", "kappa", "ce", "cq_latent", "cq_rain",
               "z_air", "z_vapor", "r_qviwv"]


class Physics(C.Structure):
    _fields_ = [(n, C.c_float) for n in PHYS_FIELDS] + [("p_emi", C.c_float * 10), ("co2_flux", C.c_float)]


class Geometry(C.Structure):
    _fields_ = [("deg", C.c_float), ("dyy", C.c_float), ("ccy_diff", C.c_float), ("ccy_adv", C.c_float),
                ("dxlat", C.c_float * YD), ("ccx_diff", C.c_float * YD), ("ccx_adv", C.c_float * YD),
                ("ccx2_diff", C.c_float * YD), ("ccx2_adv", C.c_float * YD),
                ("polar", C.c_int * YD), ("time2_diff", C.c_int * YD), ("time2_adv", C.c_int * YD)]


def build(force: bool = False) -> None:
    """Compile the oracle in-tree with the committed Makefile (gcc only)."""
    src = [os.path.join(HERE, f) for f in ("greb_oracle.c", "greb_oracle.h", "greb_oracle_cli.c", "Makefile")]
    newest = max(os.path.getmtime(s) for s in src)
    if (not force and os.path.exists(LIB) and os.path.exists(CLI)
            and min(os.path.getmtime(LIB), os.path.getmtime(CLI)) >= newest):
        return
    subprocess.run(["make", "-C", HERE, "-s"], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        fp = C.POINTER(C.c_float)
        L.go_create.restype = C.c_void_p
        L.go_destroy.argtypes = [C.c_void_p]
        L.go_physics_defaults.argtypes = [C.POINTER(Physics)]
        L.go_physics_original.argtypes = [C.POINTER(Physics)]
        L.go_set_physics.argtypes = [C.c_void_p, C.POINTER(Physics)]
        L.go_get_physics.argtypes = [C.c_void_p, C.POINTER(Physics)]
        L.go_set_forcing.argtypes = [C.c_void_p] + [fp] * 10
        L.go_setup.argtypes = [C.c_void_p]
        L.go_qflux_correction.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.go_run_scenario.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, fp, fp, C.c_int]
        L.go_time_loop.argtypes = [C.c_void_p, C.c_int, C.c_float, fp]
        L.go_time_loop.restype = C.c_int
        L.go_get_state.argtypes = [C.c_void_p, C.c_int, fp]
        L.go_set_state.argtypes = [C.c_void_p, C.c_int, fp]
        L.go_fluxcorr.argtypes = [C.c_void_p, C.c_int]
        L.go_fluxcorr.restype = fp
        L.go_get_derived.argtypes = [C.c_void_p, C.c_int, fp]
        L.go_set_ityr.argtypes = [C.c_void_p, C.c_int]
        L.go_diffusion.argtypes = [C.c_void_p, fp, fp, fp]
        L.go_advection.argtypes = [C.c_void_p, fp, fp, fp]
        L.go_circulation.argtypes = [C.c_void_p, fp, fp, fp]
        L.go_SWradiation.argtypes = [C.c_void_p, fp, fp, fp]
        L.go_LWradiation.argtypes = [C.c_void_p, fp, fp, fp, C.c_float, fp, fp, fp, fp]
        L.go_hydro.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, fp]
        L.go_seaice.argtypes = [C.c_void_p, fp]
        L.go_deep_ocean.argtypes = [C.c_void_p, fp, fp, fp, fp]
        L.go_geometry_compute.argtypes = [C.c_float, C.c_float, C.POINTER(Geometry)]
        L.go_libm_array.argtypes = [C.c_int, fp, fp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def default_physics() -> Physics:
    p = Physics()
    lib().go_physics_defaults(C.byref(p))
    return p


def original_physics() -> Physics:
    p = Physics()
    lib().go_physics_original(C.byref(p))
    return p


def geometry(pi: float = 3.1416, kappa: float = 8e5) -> Geometry:
    g = Geometry()
    lib().go_geometry_compute(C.c_float(pi), C.c_float(kappa), C.byref(g))
    return g


def host_libm(which: str, x) -> np.ndarray:
    """glibc's expf / logf on an array (what the reference's exp / log call)"""
    x = _f(x).ravel()
    y = np.zeros_like(x)
    lib().go_libm_array({"exp": 0, "log": 1}[which], _p(x), _p(y), x.size)
    return y


STATE = {"Ts": 0, "Ta": 1, "To": 2, "q": 3, "cap_surf": 4}


class Oracle:
    """One reference model instance (the Fortran module state of one ./greb process)."""

    def __init__(self, forcing, physics: Physics | None = None, **overrides):
        self.L = lib()
        self.h = self.L.go_create()
        p = physics if physics is not None else default_physics()
        for k, v in overrides.items():
            if k == "p_emi":
                for i, x in enumerate(v):
                    p.p_emi[i] = x
            else:
                setattr(p, k, v)
        self.physics = p
        self.L.go_set_physics(self.h, C.byref(p))
        f = forcing
        self._keep = [_f(a) for a in (f.z_topo, f.glacier, f.sw_solar, f.tclim, f.qclim, f.swetclim,
                                      f.uclim, f.vclim, f.mldclim, f.cldclim)]
        self.L.go_set_forcing(self.h, *[_p(a) for a in self._keep])
        self.L.go_setup(self.h)

    def __del__(self):
        try:
            self.L.go_destroy(self.h)
        except Exception:
            pass

    # ---- model-level ----
    def spinup(self, years: int, co2: float | None = None):
        self.L.go_qflux_correction(self.h, years, C.c_float(self.physics.co2_flux if co2 is None else co2))

    def run(self, years: int, co2_ppm=680.0, year0: int = 1940, want_output: bool = True, continue_run: bool = False):
        co2 = np.full(years, co2_ppm, dtype=np.float32) if np.isscalar(co2_ppm) else _f(co2_ppm)
        assert co2.shape[0] >= years
        out = np.zeros((years, 12, 5, YD, XD), dtype=np.float32) if want_output else None
        gm = np.zeros(years, dtype=np.float32)
        self.L.go_run_scenario(self.h, years, _p(co2), year0, _p(out) if want_output else None, _p(gm),
                               int(continue_run))
        return out, gm

    def time_loop(self, it: int, co2: float):
        out5 = np.zeros((5, YD, XD), dtype=np.float32)
        wrote = self.L.go_time_loop(self.h, it, C.c_float(co2), _p(out5))
        return out5 if wrote else None

    def get(self, name: str) -> np.ndarray:
        a = np.zeros((YD, XD), dtype=np.float32)
        self.L.go_get_state(self.h, STATE[name], _p(a))
        return a

    def set(self, name: str, a) -> None:
        a = _f(a)
        self.L.go_set_state(self.h, STATE[name], _p(a))

    def fluxcorr(self, which: int) -> np.ndarray:
        ptr = self.L.go_fluxcorr(self.h, which)
        return np.ctypeslib.as_array(ptr, shape=(NT, YD, XD)).copy()

    def derived(self, name: str) -> np.ndarray:
        a = np.zeros((YD, XD), dtype=np.float32)
        self.L.go_get_derived(self.h, {"wz_air": 0, "wz_vapor": 1, "z_ocean": 2, "Toclim": 3}[name], _p(a))
        return a

    def set_ityr(self, ityr: int) -> None:
        self.L.go_set_ityr(self.h, ityr)

    # ---- kernel-level (argument meaning = the Fortran subroutines) ----
    def _out(self, n=1):
        return [np.zeros((YD, XD), dtype=np.float32) for _ in range(n)]

    def diffusion(self, T1, wz):
        T1, wz = _f(T1), _f(wz)
        (o,) = self._out()
        self.L.go_diffusion(self.h, _p(T1), _p(o), _p(wz))
        return o

    def advection(self, T1, wz, ityr: int):
        T1, wz = _f(T1), _f(wz)
        self.set_ityr(ityr)
        (o,) = self._out()
        self.L.go_advection(self.h, _p(T1), _p(o), _p(wz))
        return o

    def circulation(self, X, wz, ityr: int):
        X, wz = _f(X), _f(wz)
        self.set_ityr(ityr)
        (o,) = self._out()
        self.L.go_circulation(self.h, _p(X), _p(o), _p(wz))
        return o

    def SWradiation(self, Ts, ityr: int):
        Ts = _f(Ts)
        self.set_ityr(ityr)
        sw, alb = self._out(2)
        self.L.go_SWradiation(self.h, _p(Ts), _p(sw), _p(alb))
        return sw, alb

    def LWradiation(self, Ts, Ta, q, co2: float, ityr: int):
        Ts, Ta, q = _f(Ts), _f(Ta), _f(q)
        self.set_ityr(ityr)
        o = self._out(4)
        self.L.go_LWradiation(self.h, _p(Ts), _p(Ta), _p(q), C.c_float(co2), *[_p(x) for x in o])
        return o  # LWsurf, LWair_up, LWair_down, em

    def hydro(self, Ts, q, ityr: int):
        Ts, q = _f(Ts), _f(q)
        self.set_ityr(ityr)
        o = self._out(4)
        self.L.go_hydro(self.h, _p(Ts), _p(q), *[_p(x) for x in o])
        return o  # Qlat, Qlat_air, dq_eva, dq_rain

    def seaice(self, Ts, ityr: int):
        Ts = _f(Ts)
        self.set_ityr(ityr)
        self.L.go_seaice(self.h, _p(Ts))
        return self.get("cap_surf")

    def deep_ocean(self, Ts, To, ityr: int):
        Ts, To = _f(Ts), _f(To)
        self.set_ityr(ityr)
        o = self._out(2)
        self.L.go_deep_ocean(self.h, _p(Ts), _p(To), *[_p(x) for x in o])
        return o  # dT_ocean, dTo
