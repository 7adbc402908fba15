"""Algorithmic (as-written) FLOP and byte model of one GREB member-step.

Counting rule: every add/sub/mul/div/compare written in the reference formulas counts 1, a
transcendental counts 1.  Per cell and circulation sub-step (reference src/greb.f90):
  x-diffusion bracket (f:620-625)           34   (5 groups x 5, 3 weight multiplies, 4 adds, ccx*, /20.)
  y-diffusion (f:587-588)                    6
  wz*(dTx+dTy) (f:721)                       2
  x-advection, main rows (f:816-820)        16
  x-advection, polar rows (f:872-878)       26
  y-advection (f:774-778)                   16
  dTx+dTy (f:913)                            1
  X + dx_diffuse + dx_advec (f:549)          2
  polar extras per sub-sub-step (f:715-718)  3   (compare, T1h+dTxh, T1h-T1)
The kernels execute fewer instructions than this (shared edge products), which is why the
roofline line in bench.py reports the as-written rate AND the measured issue rate from ncu.
"""
from __future__ import annotations

import math

import numpy as np

XD, YD, NT, NSUB = 96, 48, 730, 24
COLUMN_FLOPS_PER_CELL = 150  # SW/LW/hydro/deep_ocean/update/seaice/accumulate incl. 9 transcendentals
F = XD * YD * 4              # bytes per field


def _nint(x):
    return int(math.floor(x + 0.5))


def row_table(pi: float = 3.1416, kappa: float = 8e5):
    """polar flag and sub-sub-step counts per latitude row (f:578-582, 652-654, 838-840), in fp32."""
    f32 = np.float32
    pi, kappa = f32(pi), f32(kappa)
    deg = f32(2.0) * pi * f32(6.371e6) / f32(360.0)
    rows = []
    for k in range(1, YD + 1):
        lat = f32(3.75) * f32(k) - f32(3.75) / f32(2.0) - f32(90.0)
        dxlat = f32(3.75) * deg * f32(math.cos(float(f32(2.0) * pi / f32(360.0) * lat)))
        polar = not (dxlat > f32(2.5e5))
        dd = max(1, _nint(float(f32(1800.0) / (dxlat * dxlat / kappa))))
        t2d = max(1, _nint(1800.0 / int(1800.0 / dd)))
        dd = max(1, _nint(float(f32(1800.0) / (dxlat / f32(10.0)))))
        t2a = max(1, _nint(1800.0 / int(1800.0 / dd)))
        rows.append((polar, t2d, t2a))
    return rows


def flops_per_member_step(pi: float = 3.1416, kappa: float = 8e5) -> float:
    per_substep = 0
    for polar, t2d, t2a in row_table(pi, kappa):
        if not polar:
            c = 34 + 6 + 2 + 16 + 16 + 1 + 2
        else:
            c = t2d * (34 + 3) + 6 + 2 + t2a * (26 + 3) + 16 + 1 + 2
        per_substep += c * XD
    return per_substep * NSUB * 2 + COLUMN_FLOPS_PER_CELL * XD * YD


def flops_per_member_year(pi: float = 3.1416, kappa: float = 8e5) -> float:
    return flops_per_member_step(pi, kappa) * NT


def bytes_per_member_year(shared_corrections: bool = False) -> dict:
    """Algorithmic HBM bytes: flux corrections read + monthly means written per
    member-year; the forcing is shared by all members of a GPU and counted once per GPU-year."""
    return {
        "fluxcorr_read": 0 if shared_corrections else 3 * NT * F,
        "monthly_written": 60 * F,
        "forcing_per_gpu_year": 10 * NT * F + NT * YD * 4,   # 10 per-step fields (greb_types.h GF_*) + sw_solar
    }


if __name__ == "__main__":
    fs = flops_per_member_step()
    print(f"as-written flop per member-step {fs:.4e}, per member-year {fs * NT:.4e}")
    print(bytes_per_member_year())
