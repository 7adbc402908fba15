#!/usr/bin/env python
"""Every kernel of both libraries once on a small workload (member kernel in both arithmetic modes
incl. the switch build, the circulation entry, the big-grid row kernel on bands with halos at the
domain edges) — the input for a memory checker where one is available (compute-sanitizer is closed
on the round-1 GPU pool; the plain run passes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))
import greb_b200  # noqa: E402
from greb_b200 import bigrid, synth  # noqa: E402

f = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))
for mode in ("exact", "fast"):
    ens = greb_b200.Ensemble(3)
    ens.set_arithmetic(mode)
    ens.set_forcing(f)
    for m in range(3):
        p = greb_b200.default_physics()
        p.kappa = [8e5, 6.5e5, 1.0e6][m]
        ens.set_member(m, p, [680.0, 700.0])
    if mode == "exact":
        ens.set_switches(2, greb_b200.lib.SW_VAPOR_DIFFUSION_ONLY | greb_b200.lib.SW_LINEAR_VAPOR_EMISSIVITY |
                         greb_b200.lib.SW_NO_ICE_ALBEDO)
    ens.init()
    ens.spinup(0)
    ens.reset_scenario()
    for it in range(1, 4):
        ens.time_loop(it)
    X = f.tclim[0]
    ens.circulation(1, 1, X, np.ones_like(X))
    assert ens.flags().sum() == 0
    ens.close()
nx, ny, s = 192, 40, 2
rng = np.random.default_rng(0)
X = (280 + rng.normal(0, 1, (ny, nx))).astype(np.float32)
wz = np.ones_like(X)
u = rng.normal(0, 3, (ny, nx)).astype(np.float32)
v = rng.normal(0, 1, (ny, nx)).astype(np.float32)
for world in (1, 3):
    for r in range(world):
        k0, k1 = bigrid.band_range(ny, world, r)
        b = bigrid.DeviceBand(nx, ny, k0, k1, s)
        b.set_fields(X, wz, u, v)
        b.substeps(s)
        b.get()
        b.close()
print("all_kernels_once: done")
