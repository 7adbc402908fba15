#!/usr/bin/env python
"""Fast arithmetic mode against the oracle: circulation differences, 5-year drift, speed."""
import os, sys, time
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
ens.set_arithmetic("fast")
ens.set_forcing(f)
ens.set_member(0, greb_b200.default_physics(), np.full(60, 680.0, np.float32))
ens.init()
rng = np.random.default_rng(1)
X = (f.tclim[10] + rng.uniform(-1, 1, (48, 96))).astype(np.float32)
q = (f.qclim[300] * rng.uniform(0.5, 1.5, (48, 96))).astype(np.float32)
for name, fld, wz in (("T", X, o.derived("wz_air")), ("q", q, o.derived("wz_vapor"))):
    for ityr in (1, 213):
        got = ens.circulation(0, ityr, fld[None], wz[None])[0]
        ref = o.circulation(fld, wz, ityr)
        d = np.abs(got.astype(np.float64) - ref)
        print(f"{name} ityr {ityr}: max |d(dX)| {d.max():.3e}  (max |dX| {np.abs(ref).max():.3e}, ulp(X) {np.spacing(np.abs(fld).max()):.2e}) rows of max {np.unravel_index(d.argmax(), d.shape)}")
years = int(sys.argv[1]) if len(sys.argv) > 1 else 5
o.spinup(1); out_o, gm_o = o.run(years, 680.0)
ens.spinup(1); ens.reset_scenario()
out_g, gm_g, _ = ens.run(years)
d = np.abs(out_g[0].astype(np.float64) - out_o)
for y in sorted(set([0, years // 2, years - 1])):
    print(f"year {y+1}: max |d| Ts {d[y,:,0].max():.2e} Ta {d[y,:,1].max():.2e} To {d[y,:,2].max():.2e} q {d[y,:,3].max():.2e} alb {d[y,:,4].max():.2e}  gmean diff {abs(gm_g[0,y]-gm_o[y]):.2e}")
ens.close()
