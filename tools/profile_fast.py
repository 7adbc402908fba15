#!/usr/bin/env python
"""profile_run.py in the fast arithmetic mode (148 members, 1 spin-up + N scenario years)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200")); sys.path.insert(0, ROOT)
import greb_b200
from greb_b200 import synth
from bench import member_physics
years = int(sys.argv[1]) if len(sys.argv) > 1 else 2
f = synth.cached_forcing(cache_dir="/tmp/greb_b200_cache")
ens = greb_b200.Ensemble(148); ens.set_arithmetic("fast"); ens.set_forcing(f)
for m in range(148):
    p, co2 = member_physics(m, greb_b200.default_physics); ens.set_member(m, p, np.full(years + 1, co2, np.float32))
ens.init(); ens.spinup(1); ens.reset_scenario()
for y in range(years):
    ens.run_raw(1); ms, n = ens.last_kernel_ms()
    print(f"FAST scenario year {y}: {ms:.2f} ms -> {148 / (ms / 1e3):.1f} member-years/s")
print("flags", int(ens.flags().sum()))
ens.close()
