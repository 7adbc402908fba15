#!/usr/bin/env python
"""Times the kernel-level circulation entry (greb_b200_circulation: 24 sub-steps of diffusion + advection per
field, one CTA per field) on N fields: the circulation without the column physics.  GREB_B200_TILE6=<late mask>
selects the 6-cell-tile kernel (greb_core6.h)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))
import greb_b200  # noqa: E402
from greb_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
arith = sys.argv[2] if len(sys.argv) > 2 else "exact"
f = synth.cached_forcing(cache_dir="/tmp/greb_b200_cache")
ens = greb_b200.Ensemble(1)
ens.set_arithmetic(arith)
ens.set_forcing(f)
ens.set_member(0, greb_b200.default_physics(), [680.0])
ens.init()
rng = np.random.default_rng(0)
X = (f.tclim[10][None] + rng.uniform(-1, 1, (n, 48, 96))).astype(np.float32)
W = np.broadcast_to(np.exp(-f.z_topo / np.float32(8400.0)).astype(np.float32), (n, 48, 96)).copy()
best = 1e9
for rep in range(4):
    out = ens.circulation(0, 11, X, W)
    ms, _ = ens.last_kernel_ms()
    best = min(best, ms)
print(f"{arith} TILE6={os.environ.get('GREB_B200_TILE6')}: {n} fields, {best:.3f} ms -> {n * 24 / best / 1e3:.2f} M field-sub-steps/s; checksum {float(np.abs(out).sum()):.6e}")
ens.close()
