#!/usr/bin/env python
"""Debug helper: one circulation call on the GPU against the oracle; prints the differing cells."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200")); sys.path.insert(0, ROOT)
import greb_b200
from greb_b200 import synth
from oracle import oracle as om
om.build()
f = synth.cached_forcing(cache_dir="/tmp/greb_b200_cache")
o = om.Oracle(f)
ens = greb_b200.Ensemble(1)
if "--fast" in sys.argv:
    ens.set_arithmetic("fast")
ens.set_forcing(f)
ens.set_member(0, greb_b200.default_physics(), [680.0])
ens.init()
rng = np.random.default_rng(1)
X = (f.tclim[10] + rng.uniform(-1, 1, (48, 96))).astype(np.float32)
wz = o.derived("wz_air")
for ityr in (1, 213):
    got = ens.circulation(0, ityr, X[None], wz[None])[0]
    ref = o.circulation(X, wz, ityr)
    bad = np.argwhere(ref != got)
    print("ityr", ityr, "n diff", len(bad))
    for k, i in bad[:40]:
        print("  row", k, "col", i, "col%12", i % 12, "ref", ref[k, i], "got", got[k, i], "u", f.uclim[ityr-1, k, i], "v", f.vclim[ityr-1, k, i])
# one sub-step level check: diffusion+advection of a single sub-step cannot be called; use X with tiny perturbation
ens.close()
