"""ctypes harness around the reference itself, machine-translated and compiled here (oracle/_ref/).

TEST INFRASTRUCTURE ONLY.  `oracle/f90_to_cpp.py` translates the modules and subroutines of
`/root/reference/src/greb.f90` (and `src/greb.original.model.f90`) statement by statement into
C++; `build()` compiles the result into `oracle/_ref/libgreb_ref.so` / `libgreb_orig_ref.so`
(git-ignored, never committed, but it travels to the GPU box).  This module plays the role of the
reference's PROGRAM unit (greb.f90:996-1098): it fills the module variables (physics namelist,
forcing arrays, `Toclim`, padded `co2_ppm`) and calls the translated `greb_model` or any single
subroutine with the reference's own argument lists.

The translated library is one Fortran program image: module variables are process globals, so
there is ONE model instance per loaded library (use `Ref.fresh()` to load a private copy).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("GREB_REFERENCE_ROOT", "/root/reference")
SOURCES = {"greb": "src/greb.f90", "orig": "src/greb.original.model.f90"}
LIBS = {"greb": os.path.join(REFDIR, "libgreb_ref.so"), "orig": os.path.join(REFDIR, "libgreb_orig_ref.so")}
XD, YD, NT = 96, 48, 730
CXXFLAGS = ["-O3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c++17", "-w"]

PHYS_SCALARS = ["pi", "sig", "rho_ocean", "rho_land", "rho_air", "cp_ocean", "cp_land", "cp_air", "eps",
                "d_ocean", "d_land", "d_air", "ct_sens", "da_ice", "a_no_ice", "a_cloud", "tl_ice1",
                "tl_ice2", "to_ice1", "to_ice2", "co_turb", "kappa", "ce", "cq_latent", "cq_rain",
                "z_air", "z_vapor", "r_qviwv"]


def reference_present(which: str = "greb") -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, SOURCES[which]))


def available(which: str = "greb") -> bool:
    return os.path.exists(LIBS[which]) or reference_present(which)


def build(which: str = "greb", force: bool = False) -> str:
    """Translate + compile the reference source where it lies; outputs only under oracle/_ref/."""
    lib = LIBS[which]
    src = os.path.join(REFERENCE_ROOT, SOURCES[which])
    tool = os.path.join(HERE, "f90_to_cpp.py")
    if not os.path.exists(src):
        if os.path.exists(lib):
            return lib          # GPU box: the prebuilt library travelled with the snapshot
        raise FileNotFoundError(f"{src} not present and {lib} not prebuilt")
    if (not force and os.path.exists(lib)
            and os.path.getmtime(lib) >= max(os.path.getmtime(src), os.path.getmtime(tool))):
        return lib
    os.makedirs(REFDIR, exist_ok=True)
    cpp = os.path.join(REFDIR, f"{which}_ref.cpp")
    import importlib.util
    spec = importlib.util.spec_from_file_location("f90_to_cpp", tool)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(cpp, "w") as fh:
        fh.write(mod.translate(src))
    subprocess.run(["g++"] + CXXFLAGS + ["-o", lib, cpp], check=True)
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Ref:
    """The translated reference program image."""

    def __init__(self, which: str = "greb", private_copy: bool = False):
        path = build(which)
        self._tmp = None
        if private_copy:
            self._tmp = tempfile.mkdtemp(prefix="greb_ref_")
            p2 = os.path.join(self._tmp, os.path.basename(path))
            shutil.copy(path, p2)
            path = p2
        self.which = which
        self.L = C.CDLL(path)
        self._keep = {}
        self.L.f90_out_nrecs.restype = C.c_size_t
        self.L.f90_out_rec_data.restype = C.POINTER(C.c_float)
        self.L.f90_out_rec_data.argtypes = [C.c_size_t]
        self.L.f90_out_rec_info.argtypes = [C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        self.L.f90_print_nvals.restype = C.c_size_t
        self.L.f90_print_data.restype = C.POINTER(C.c_double)
        self.L.f90_print_nlines.restype = C.c_size_t
        self.L.f90_print_counts.restype = C.POINTER(C.c_int)

    @classmethod
    def fresh(cls, which: str = "greb"):
        return cls(which, private_copy=True)

    def __del__(self):
        if getattr(self, "_tmp", None):
            shutil.rmtree(self._tmp, ignore_errors=True)

    # ---- module variables -------------------------------------------------------------------
    def has(self, name: str) -> bool:
        return hasattr(self.L, "f_" + name.lower())

    def scalar(self, name: str, ctype=C.c_float):
        return ctype.in_dll(self.L, "f_" + name.lower())

    def get(self, name: str, ctype=C.c_float):
        return self.scalar(name, ctype).value

    def set(self, name: str, value, ctype=C.c_float):
        self.scalar(name, ctype).value = value

    def seti(self, name: str, value: int):
        self.set(name, int(value), C.c_int)

    def geti(self, name: str) -> int:
        return self.get(name, C.c_int)

    def array(self, name: str, shape, dtype=np.float32) -> np.ndarray:
        """numpy VIEW of a module array; `shape` in C order (= reversed Fortran dims)."""
        n = int(np.prod(shape))
        ct = C.c_float if dtype == np.float32 else C.c_int
        buf = (ct * n).in_dll(self.L, "f_" + name.lower())
        return np.ctypeslib.as_array(buf).reshape(shape)

    def set_allocatable(self, name: str, a: np.ndarray):
        a = _f32(a)
        self._keep[name] = a
        C.c_void_p.in_dll(self.L, "f_" + name.lower()).value = a.ctypes.data

    # ---- the PROGRAM unit's job (greb.f90:1042-1094) -------------------------------------------
    def set_physics(self, **overrides):
        for k, v in overrides.items():
            k = k.lower()
            if k == "p_emi":
                self.array("p_emi", (10,))[:] = np.asarray(v, dtype=np.float32)
            elif k == "co2_flux":
                self.set("co2_flux", float(v))
            elif k in PHYS_SCALARS:
                self.set(k, float(np.float32(v)))
            else:
                raise KeyError(k)

    def set_forcing(self, f):
        """greb.f90:1073-1094: the ten input fields + Toclim = max(min_t Tclim, -1.7+273.15)."""
        self.array("z_topo", (YD, XD))[:] = f.z_topo
        self.array("glacier", (YD, XD))[:] = f.glacier
        self.array("sw_solar", (NT, YD))[:] = f.sw_solar
        for name, a in (("tclim", f.tclim), ("qclim", f.qclim), ("swetclim", f.swetclim), ("uclim", f.uclim),
                        ("vclim", f.vclim), ("mldclim", f.mldclim), ("cldclim", f.cldclim)):
            self.array(name, (NT, YD, XD))[:] = a
        to = f.tclim.min(axis=0).astype(np.float32)
        floor = np.float32(np.float32(-1.7) + np.float32(273.15))
        to = np.where((to - np.float32(273.15)).astype(np.float32) < np.float32(-1.7), floor, to).astype(np.float32)
        self.array("toclim", (NT, YD, XD))[:] = to[None]

    def set_run(self, time_flux: int, time_scnr: int, co2_ppm, year0: int = 1940, ipx: int = 1, ipy: int = 1):
        """numerics_par + co2_par incl. the padding rule of greb.f90:1047-1061."""
        self.seti("time_flux", time_flux)
        self.seti("time_scnr", time_scnr)
        self.seti("year0", year0)
        self.seti("ipx", ipx)
        self.seti("ipy", ipy)
        co2 = np.full(max(time_scnr, 1), -1.0, dtype=np.float32)
        given = np.atleast_1d(np.asarray(co2_ppm, dtype=np.float32))[:time_scnr]
        co2[:len(given)] = given
        if co2[0] == -1:
            co2[0] = 680
        for i in range(1, time_scnr):
            if co2[i] < 0:
                co2[i:] = co2[i - 1]
                break
        self.set_allocatable("co2_ppm", co2)

    # ---- calls -----------------------------------------------------------------------------------
    def call(self, sub: str, *args):
        """Call a translated subroutine; numpy arrays pass by reference, Python floats/ints as
        temporaries.  Returns the list of ctypes scalars created (to read back inout values)."""
        fn = getattr(self.L, "f_" + sub.lower())
        conv, scal = [], []
        for a in args:
            if isinstance(a, np.ndarray):
                assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.float32, np.int32)
                conv.append(a.ctypes.data_as(C.c_void_p))
            elif isinstance(a, (C.c_float, C.c_int)):
                conv.append(C.byref(a))
                scal.append(a)
            elif isinstance(a, (int, np.integer)) and not isinstance(a, bool):
                s = C.c_int(int(a))
                conv.append(C.byref(s))
                scal.append(s)
            else:
                s = C.c_float(float(a))
                conv.append(C.byref(s))
                scal.append(s)
        fn.restype = None
        fn(*conv)
        return scal

    def reset_output(self):
        self.L.f90_out_reset()

    def record_output(self, on: bool):
        self.L.f90_set_record_output(int(on))

    def output_records(self):
        """[(unit, rec, array)] of every direct-access write since reset_output()."""
        n = self.L.f90_out_nrecs()
        recs = []
        for i in range(n):
            u, r, cnt = C.c_int(), C.c_int(), C.c_size_t()
            self.L.f90_out_rec_info(i, C.byref(u), C.byref(r), C.byref(cnt))
            data = np.ctypeslib.as_array(self.L.f90_out_rec_data(i), shape=(cnt.value,)).copy()
            recs.append((u.value, r.value, data))
        return recs

    def output_file(self, unit: int = 22) -> np.ndarray:
        """The direct-access file image of `unit`: records ordered by record number."""
        recs = {}
        for u, r, d in self.output_records():
            if u == unit:
                recs[r] = d
        if not recs:
            return np.zeros((0, YD, XD), dtype=np.float32)
        out = np.zeros((max(recs), YD * XD), dtype=np.float32)
        for r, d in recs.items():
            out[r - 1] = d
        return out.reshape(-1, YD, XD)

    def console(self):
        """numeric items of every `print *` since reset_output(), one list per line."""
        nl = self.L.f90_print_nlines()
        counts = np.ctypeslib.as_array(self.L.f90_print_counts(), shape=(nl,)) if nl else []
        nv = self.L.f90_print_nvals()
        vals = np.ctypeslib.as_array(self.L.f90_print_data(), shape=(nv,)) if nv else np.zeros(0)
        lines, pos = [], 0
        for c in counts:
            lines.append(vals[pos:pos + c].tolist())
            pos += c
        return lines

    def greb_model(self):
        self.reset_output()
        self.call("greb_model")
        return self.output_file(22)
