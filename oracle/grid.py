"""ctypes wrapper of oracle/grid_oracle.c — the circulation restated for arbitrary grids.

TEST INFRASTRUCTURE ONLY: parity oracle of the big-grid path (BASELINE.json configs[4]); only
tests/, __graft_entry__.smoke() and bench-type tools used as checkers may import it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgrid_oracle.so")
SRC = os.path.join(HERE, "grid_oracle.c")
CFLAGS = ["-O3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c11", "-shared"]
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or (os.path.exists(SRC) and os.path.getmtime(LIB) < os.path.getmtime(SRC)):
        subprocess.run(["gcc"] + CFLAGS + ["-o", LIB, SRC, "-lm"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.gg_geometry.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float] + [fp] * 5 + [ip] * 3 + [fp]
        L.gg_geometry.restype = C.c_int
        L.gg_substep.argtypes = [C.c_int] * 4 + [fp] * 4 + [C.c_float] * 2 + [fp] * 4 + [ip] * 3 + [fp]
        L.gg_substep.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(fp)


def _pi(a):
    return a.ctypes.data_as(ip)


class Geometry:
    """f:543, 578-582, 652-654, 749-753, 838-840 with the declared rules R1/R2 of grid_oracle.c"""

    def __init__(self, nx: int, ny: int, pi: float = 3.1416, kappa: float = 8e5):
        self.nx, self.ny = nx, ny
        f = lambda: np.zeros(ny, dtype=np.float32)
        i = lambda: np.zeros(ny, dtype=np.int32)
        self.dxlat, self.ccx_diff, self.ccx_adv, self.ccx2_diff, self.ccx2_adv = f(), f(), f(), f(), f()
        self.polar, self.time2_diff, self.time2_adv = i(), i(), i()
        sc = np.zeros(3, dtype=np.float32)
        self.nsub = lib().gg_geometry(nx, ny, pi, kappa, _p(self.dxlat), _p(self.ccx_diff), _p(self.ccx_adv),
                                      _p(self.ccx2_diff), _p(self.ccx2_adv), _pi(self.polar), _pi(self.time2_diff),
                                      _pi(self.time2_adv), _p(sc))
        self.dt_crcl, self.ccy_diff, self.ccy_adv = (float(x) for x in sc)


def substep(g: Geometry, X, wz, u, v, r0: int = 0, r1: int | None = None, out=None):
    """one circulation sub-step for the global rows [r0, r1); X, wz, u, v are full [ny][nx] fields"""
    r1 = g.ny if r1 is None else r1
    X, wz, u, v = (np.ascontiguousarray(a, dtype=np.float32) for a in (X, wz, u, v))
    assert X.shape == (g.ny, g.nx)
    out = np.array(X, copy=True, order="C") if out is None else out
    assert out.flags["C_CONTIGUOUS"] and out.dtype == np.float32 and out.shape == X.shape
    lib().gg_substep(g.nx, g.ny, r0, r1, _p(X), _p(wz), _p(u), _p(v), g.ccy_diff, g.ccy_adv, _p(g.ccx_diff),
                     _p(g.ccx_adv), _p(g.ccx2_diff), _p(g.ccx2_adv), _pi(g.polar), _pi(g.time2_diff),
                     _pi(g.time2_adv), _p(out))
    return out


def substeps(g: Geometry, X, wz, u, v, n: int):
    """n sub-steps of the whole domain (double-buffered)"""
    a = np.array(X, dtype=np.float32, copy=True, order="C")
    b = np.empty_like(a)
    for _ in range(n):
        substep(g, a, wz, u, v, out=b)
        a, b = b, a
    return a
