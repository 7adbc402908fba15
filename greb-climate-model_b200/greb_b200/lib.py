"""ctypes binding of the C ABI in include/greb_b200.h (the drop-in boundary).

Fails loudly when the CUDA library is missing or no B200 is usable — there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

XD, YD, NT = 96, 48, 730
NC = XD * YD
PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")

PHYS_FIELDS = ["pi", "sig", "rho_ocean", "rho_land", "rho_air", "cp_ocean", "cp_land", "cp_air", "eps",
               "d_ocean", "d_land", "d_air", "ct_sens", "da_ice", "a_no_ice", "a_cloud", "Tl_ice1",
               "Tl_ice2", "To_ice1", "To_ice2", "co_turb", "kappa", "ce", "cq_latent", "cq_rain",
               "z_air", "z_vapor", "r_qviwv"]

# every symbol include/greb_b200.h declares
ABI_SYMBOLS = ["greb_b200_physics_defaults", "greb_b200_physics_original", "greb_b200_create", "greb_b200_destroy",
               "greb_b200_last_error", "greb_b200_n_members", "greb_b200_set_arithmetic", "greb_b200_set_forcing", "greb_b200_set_member", "greb_b200_set_switches",
               "greb_b200_pad_co2", "greb_b200_init", "greb_b200_spinup", "greb_b200_reset_scenario",
               "greb_b200_run", "greb_b200_time_loop", "greb_b200_get_state", "greb_b200_set_state",
               "greb_b200_get_states", "greb_b200_set_states",
               "greb_b200_get_fluxcorr", "greb_b200_set_fluxcorr", "greb_b200_get_monthly", "greb_b200_diag_device", "greb_b200_get_flags",
               "greb_b200_circulation", "greb_b200_last_kernel_ms",
               "greb_b200_run_async", "greb_b200_wait", "greb_b200_time_steps", "greb_b200_set_states_async",
               "greb_b200_get_states_async", "greb_b200_sync_compute", "greb_b200_get_calendar",
               "greb_b200_set_calendar", "greb_b200_get_accumulators", "greb_b200_set_accumulators",
               "greb_b200_device_libm", "greb_b200_ensemble_moments", "greb_b200_ensemble_moments_device",
               "greb_b200_fetch_monthly_async", "greb_b200_tile_phase", "greb_b200_wz"]


class Physics(C.Structure):
    """struct greb_physics_par (namelist physics_par + co2_flux, reference src/greb.f90:68-104)."""
    _fields_ = [(n, C.c_float) for n in PHYS_FIELDS] + [("p_emi", C.c_float * 10), ("co2_flux", C.c_float)]

    def copy(self) -> "Physics":
        p = Physics()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Physics))
        return p


class GrebError(RuntimeError):
    pass


def library_path() -> str:
    # GREB_B200_LIB: load another build of the same library (kernel experiments); default = in-tree
    return os.environ.get("GREB_B200_LIB") or os.path.join(PKG_DIR, "libgreb_b200.so")


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise GrebError("building libgreb_b200.so failed")
    return library_path()


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise GrebError(f"{path} is missing: build it with greb_b200.build_library() / make -C {CSRC} "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    L.greb_b200_physics_defaults.argtypes = [C.POINTER(Physics)]
    L.greb_b200_physics_defaults.restype = None
    L.greb_b200_physics_original.argtypes = [C.POINTER(Physics)]
    L.greb_b200_physics_original.restype = None
    L.greb_b200_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    L.greb_b200_destroy.argtypes = [vp]
    L.greb_b200_last_error.argtypes = [vp]
    L.greb_b200_last_error.restype = C.c_char_p
    L.greb_b200_n_members.argtypes = [vp]
    L.greb_b200_set_arithmetic.argtypes = [vp, C.c_int]
    L.greb_b200_set_forcing.argtypes = [vp] + [fp] * 10
    L.greb_b200_set_member.argtypes = [vp, C.c_int, C.POINTER(Physics), fp, C.c_int, C.c_int]
    L.greb_b200_set_switches.argtypes = [vp, C.c_int, C.c_uint]
    L.greb_b200_pad_co2.argtypes = [fp, C.c_int, fp, C.c_int]
    L.greb_b200_pad_co2.restype = None
    L.greb_b200_wz.argtypes = [fp, C.c_float, fp, C.c_long]
    L.greb_b200_wz.restype = None
    L.greb_b200_init.argtypes = [vp]
    L.greb_b200_spinup.argtypes = [vp, C.c_int]
    L.greb_b200_reset_scenario.argtypes = [vp]
    L.greb_b200_run.argtypes = [vp, C.c_int, vp, ip, C.c_int, fp, fp]
    L.greb_b200_run_async.argtypes = [vp, C.c_int, vp, ip, C.c_int, fp, fp]
    L.greb_b200_wait.argtypes = [vp]
    L.greb_b200_fetch_monthly_async.argtypes = [vp, vp, ip, C.c_int]
    L.greb_b200_time_loop.argtypes = [vp, C.c_int]
    L.greb_b200_time_steps.argtypes = [vp, C.c_int, C.c_int]
    L.greb_b200_set_states_async.argtypes = [vp, vp]
    L.greb_b200_get_states_async.argtypes = [vp, vp]
    L.greb_b200_sync_compute.argtypes = [vp]
    L.greb_b200_get_calendar.argtypes = [vp, ip]
    L.greb_b200_set_calendar.argtypes = [vp, C.c_int]
    L.greb_b200_get_accumulators.argtypes = [vp, vp]
    L.greb_b200_set_accumulators.argtypes = [vp, vp]
    L.greb_b200_get_state.argtypes = [vp, C.c_int, C.c_int, fp]
    L.greb_b200_set_state.argtypes = [vp, C.c_int, C.c_int, fp]
    L.greb_b200_get_states.argtypes = [vp, vp]
    L.greb_b200_set_states.argtypes = [vp, vp]
    L.greb_b200_get_fluxcorr.argtypes = [vp, C.c_int, C.c_int, fp]
    L.greb_b200_set_fluxcorr.argtypes = [vp, C.c_int, C.c_int, fp]
    L.greb_b200_get_monthly.argtypes = [vp, C.c_int, fp]
    L.greb_b200_diag_device.argtypes = [vp, C.POINTER(vp), ip]
    L.greb_b200_get_flags.argtypes = [vp, ip]
    L.greb_b200_circulation.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, C.c_int]
    L.greb_b200_last_kernel_ms.argtypes = [vp, fp, ip]
    L.greb_b200_device_libm.argtypes = [vp, C.c_int, fp, fp, C.c_int]
    L.greb_b200_tile_phase.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Physics), C.c_float] + [vp] * 10
    L.greb_b200_ensemble_moments.argtypes = [vp, vp, vp]
    L.greb_b200_ensemble_moments_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip]
    _lib = L
    return L


def default_physics() -> Physics:
    p = Physics()
    load_library().greb_b200_physics_defaults(C.byref(p))
    return p


def original_physics() -> Physics:
    p = Physics()
    load_library().greb_b200_physics_original(C.byref(p))
    return p


def wz_field(z_topo, h_scale: float) -> np.ndarray:
    """exp(-z_topo / h_scale) with the library's host libm (reference src/greb.f90:201-202), any shape"""
    z = np.ascontiguousarray(z_topo, dtype=np.float32)
    out = np.empty_like(z)
    load_library().greb_b200_wz(_p(z), h_scale, _p(out), z.size)
    return out


def pad_co2(given, n_years: int) -> np.ndarray:
    """co2_ppm padding of reference src/greb.f90:1047-1061."""
    given = np.ascontiguousarray(np.atleast_1d(given), dtype=np.float32)
    out = np.zeros(n_years, dtype=np.float32)
    load_library().greb_b200_pad_co2(_p(given), len(given), _p(out), n_years)
    return out


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


STATE = {"Ts": 0, "Ta": 1, "To": 2, "q": 3, "cap_surf": 4}
# include/greb_b200.h GREB_SW_*: process switches (greb.original.model.f90 log_exp experiments)
SW_NO_ICE_ALBEDO, SW_NO_HYDRO, SW_NO_DEEP_OCEAN, SW_VAPOR_DIFFUSION_ONLY, SW_LINEAR_VAPOR_EMISSIVITY, \
    SW_SST_PLUS_1K, SW_NO_HEAT_CIRCULATION, SW_NO_VAPOR_CIRCULATION = 1, 2, 4, 8, 16, 32, 64, 128


class Ensemble:
    """A batch of independent GREB members on one B200 (one handle of the C ABI).

    Mirrors what one ``./greb <namelist>`` process per member does in the reference
    (src/greb.f90:1030-1038, 1063-1068), for ``n_members`` members at once."""

    def __init__(self, n_members: int, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.greb_b200_create(C.byref(self.h), n_members, device)
        if rc != 0:
            msg = self.L.greb_b200_last_error(None).decode()
            self.h = None
            raise GrebError(f"greb_b200_create failed ({rc}): {msg}")
        self.n = n_members
        self.years_run = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.greb_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise GrebError(f"{what} failed ({rc}): {self.L.greb_b200_last_error(self.h).decode()}")

    def set_arithmetic(self, mode: str):
        """'exact' (default: bit-identical circulation) or 'fast' (factored stencils + FMA, within the
        north_star tolerances); include/greb_b200.h GREB_ARITH_*."""
        self._ck(self.L.greb_b200_set_arithmetic(self.h, {"exact": 0, "fast": 1}[mode]), "greb_b200_set_arithmetic")

    def set_forcing(self, f):
        arrs = [_f(a) for a in (f.z_topo, f.glacier, f.sw_solar, f.tclim, f.qclim, f.swetclim, f.uclim, f.vclim,
                                f.mldclim, f.cldclim)]
        self._ck(self.L.greb_b200_set_forcing(self.h, *[_p(a) for a in arrs]), "greb_b200_set_forcing")

    def set_member(self, m: int, physics: Physics, co2_ppm, year0: int = 1940):
        co2 = _f(np.atleast_1d(co2_ppm))
        self._ck(self.L.greb_b200_set_member(self.h, m, C.byref(physics), _p(co2), len(co2), year0),
                 "greb_b200_set_member")

    def set_switches(self, m: int, mask: int):
        """GREB_SW_* process switches of member m (include/greb_b200.h)."""
        self._ck(self.L.greb_b200_set_switches(self.h, m, mask), "greb_b200_set_switches")

    def init(self):
        self._ck(self.L.greb_b200_init(self.h), "greb_b200_init")
        self.years_run = 0

    def spinup(self, years: int):
        self._ck(self.L.greb_b200_spinup(self.h, years), "greb_b200_spinup")

    def reset_scenario(self):
        self._ck(self.L.greb_b200_reset_scenario(self.h), "greb_b200_reset_scenario")
        self.years_run = 0

    def run(self, years: int, want_output: bool = True, out_members=None, out: np.ndarray | None = None):
        """Returns (monthly [n_out][years][12][5][48][96] or None, gmean [n][years], gmean_coslat [n][years])."""
        n_out = self.n if out_members is None else len(out_members)
        if want_output and out is None:
            out = np.zeros((n_out, years, 12, 5, YD, XD), dtype=np.float32)
        gm = np.zeros((self.n, years), dtype=np.float32)
        gc = np.zeros((self.n, years), dtype=np.float32)
        om = None
        if out_members is not None:
            om = np.ascontiguousarray(out_members, dtype=np.int32)
        rc = self.L.greb_b200_run(self.h, years, out.ctypes.data_as(C.c_void_p) if want_output else None,
                                  om.ctypes.data_as(C.POINTER(C.c_int)) if om is not None else None, n_out, _p(gm),
                                  _p(gc))
        self._ck(rc, "greb_b200_run")
        self.years_run += years
        return (out if want_output else None), gm, gc

    def run_raw(self, years: int, out_ptr=None):
        """greb_b200_run without diagnostics copies; out_ptr = host address (e.g. pinned) or None."""
        rc = self.L.greb_b200_run(self.h, years, C.c_void_p(out_ptr) if out_ptr else None, None, self.n, None, None)
        self._ck(rc, "greb_b200_run")
        self.years_run += years

    def run_async(self, years: int, out_ptr=None, out_members=None):
        """greb_b200_run_async: enqueue `years` years and return; records go to the host address `out_ptr`
        (pinned memory, valid until the next wait()).  No diagnostics arrays: read diag_device() or call run()."""
        om = None if out_members is None else np.ascontiguousarray(out_members, dtype=np.int32)
        n_out = self.n if om is None else len(om)
        rc = self.L.greb_b200_run_async(self.h, years, C.c_void_p(out_ptr) if out_ptr else None,
                                        om.ctypes.data_as(C.POINTER(C.c_int)) if om is not None else None, n_out,
                                        None, None)
        self._ck(rc, "greb_b200_run_async")
        self._keep_om = om
        self.years_run += years

    def fetch_monthly_async(self, out_ptr: int):
        """records of the last completed year of ALL members -> host address `out_ptr` (pinned), behind what
        the compute stream holds now; wait() completes it"""
        self._ck(self.L.greb_b200_fetch_monthly_async(self.h, C.c_void_p(out_ptr), None, self.n),
                 "greb_b200_fetch_monthly_async")

    def wait(self):
        self._ck(self.L.greb_b200_wait(self.h), "greb_b200_wait")

    def sync_compute(self):
        self._ck(self.L.greb_b200_sync_compute(self.h), "greb_b200_sync_compute")

    def set_states_async(self, ptr: int):
        self._ck(self.L.greb_b200_set_states_async(self.h, C.c_void_p(ptr)), "greb_b200_set_states_async")

    def get_states_async(self, ptr: int):
        self._ck(self.L.greb_b200_get_states_async(self.h, C.c_void_p(ptr)), "greb_b200_get_states_async")

    def time_loop(self, it: int):
        self._ck(self.L.greb_b200_time_loop(self.h, it), "greb_b200_time_loop")

    def time_steps(self, it0: int, nsteps: int):
        """`nsteps` (<= 730) consecutive time_loop calls it0, it0+1, ... in one launch"""
        self._ck(self.L.greb_b200_time_steps(self.h, it0, nsteps), "greb_b200_time_steps")

    # ---- checkpoint / resume (include/greb_b200.h) ----
    def get_calendar(self) -> int:
        it = C.c_int()
        self._ck(self.L.greb_b200_get_calendar(self.h, C.byref(it)), "greb_b200_get_calendar")
        return it.value

    def set_calendar(self, it_next: int):
        self._ck(self.L.greb_b200_set_calendar(self.h, int(it_next)), "greb_b200_set_calendar")

    def get_accumulators(self) -> np.ndarray:
        a = np.zeros((self.n, 6, YD, XD), dtype=np.float32)
        self._ck(self.L.greb_b200_get_accumulators(self.h, C.c_void_p(a.ctypes.data)), "greb_b200_get_accumulators")
        return a

    def set_accumulators(self, a):
        a = _f(a)
        assert a.shape == (self.n, 6, YD, XD)
        self._ck(self.L.greb_b200_set_accumulators(self.h, C.c_void_p(a.ctypes.data)), "greb_b200_set_accumulators")

    def get_state(self, m: int, name: str) -> np.ndarray:
        a = np.zeros((YD, XD), dtype=np.float32)
        self._ck(self.L.greb_b200_get_state(self.h, m, STATE[name], _p(a)), "greb_b200_get_state")
        return a

    def set_state(self, m: int, name: str, a):
        a = _f(a)
        self._ck(self.L.greb_b200_set_state(self.h, m, STATE[name], _p(a)), "greb_b200_set_state")

    def get_states(self, out: np.ndarray | None = None, ptr: int | None = None):
        """all members' state [n][5][48][96]; ptr = raw host address (e.g. pinned) instead of an array"""
        if ptr is None:
            out = np.zeros((self.n, 5, YD, XD), dtype=np.float32) if out is None else out
            ptr = out.ctypes.data
        self._ck(self.L.greb_b200_get_states(self.h, C.c_void_p(ptr)), "greb_b200_get_states")
        return out

    def set_states(self, arr: np.ndarray | None = None, ptr: int | None = None):
        if ptr is None:
            arr = _f(arr)
            assert arr.shape == (self.n, 5, YD, XD)
            ptr = arr.ctypes.data
        self._ck(self.L.greb_b200_set_states(self.h, C.c_void_p(ptr)), "greb_b200_set_states")

    def get_fluxcorr(self, m: int, which: int) -> np.ndarray:
        a = np.zeros((NT, YD, XD), dtype=np.float32)
        self._ck(self.L.greb_b200_get_fluxcorr(self.h, m, which, _p(a)), "greb_b200_get_fluxcorr")
        return a

    def set_fluxcorr(self, m: int, which: int, a) -> None:
        a = _f(a)
        assert a.shape == (NT, YD, XD)
        self._ck(self.L.greb_b200_set_fluxcorr(self.h, m, which, _p(a)), "greb_b200_set_fluxcorr")

    def get_monthly(self, m: int) -> np.ndarray:
        a = np.zeros((12, 5, YD, XD), dtype=np.float32)
        self._ck(self.L.greb_b200_get_monthly(self.h, m, _p(a)), "greb_b200_get_monthly")
        return a

    def diag_device(self):
        ptr = C.c_void_p()
        n = C.c_int()
        self._ck(self.L.greb_b200_diag_device(self.h, C.byref(ptr), C.byref(n)), "greb_b200_diag_device")
        return ptr.value, n.value

    def flags(self) -> np.ndarray:
        a = np.zeros(self.n, dtype=np.int32)
        self._ck(self.L.greb_b200_get_flags(self.h, a.ctypes.data_as(C.POINTER(C.c_int))), "greb_b200_get_flags")
        return a

    def circulation(self, member: int, ityr: int, X, wz) -> np.ndarray:
        X, wz = _f(X), _f(wz)
        n = 1 if X.ndim == 2 else X.shape[0]
        out = np.zeros_like(X)
        self._ck(self.L.greb_b200_circulation(self.h, member, ityr, _p(X), _p(wz), _p(out), n),
                 "greb_b200_circulation")
        return out

    def ensemble_moments(self):
        """(sum, sum of squares) over the members of the last completed year's monthly means,
        float64 [12][5][48][96] each, reduced on the device"""
        s = np.zeros((12, 5, YD, XD), dtype=np.float64)
        q = np.zeros_like(s)
        self._ck(self.L.greb_b200_ensemble_moments(self.h, C.c_void_p(s.ctypes.data), C.c_void_p(q.ctypes.data)),
                 "greb_b200_ensemble_moments")
        return s, q

    def ensemble_moments_device(self):
        """device addresses of the same two float64 vectors + their length (for an NCCL reduce)"""
        ps, pq, n = C.c_void_p(), C.c_void_p(), C.c_int()
        self._ck(self.L.greb_b200_ensemble_moments_device(self.h, C.byref(ps), C.byref(pq), C.byref(n)),
                 "greb_b200_ensemble_moments_device")
        return ps.value, pq.value, n.value

    def device_libm(self, which: str, x) -> np.ndarray:
        """the exact mode's expf / logf (glibc's algorithm on the device) on an array"""
        x = _f(x).ravel()
        y = np.zeros_like(x)
        self._ck(self.L.greb_b200_device_libm(self.h, {"exp": 0, "log": 1}[which], _p(x), _p(y), x.size),
                 "greb_b200_device_libm")
        return y

    def last_kernel_ms(self):
        ms = C.c_float()
        n = C.c_int()
        self.L.greb_b200_last_kernel_ms(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value
