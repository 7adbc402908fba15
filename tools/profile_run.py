#!/usr/bin/env python
"""Small fixed workload for ncu: N members (default one wave = 148), 1 spin-up year + Y scenario
years; prints the CUDA-event time of each phase.  Used both plain and under ncu (same command)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))
sys.path.insert(0, ROOT)
import greb_b200  # noqa: E402
from greb_b200 import synth  # noqa: E402
from bench import member_physics  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=148)
ap.add_argument("--years", type=int, default=2)
ap.add_argument("--shared", action="store_true")
ap.add_argument("--no-spinup", action="store_true")
ap.add_argument("--arith", default="exact", choices=["exact", "fast"])
a = ap.parse_args()

f = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))
ens = greb_b200.Ensemble(a.members)
ens.set_arithmetic(a.arith)
ens.set_forcing(f)
for m in range(a.members):
    p, co2 = member_physics(m, greb_b200.default_physics)
    if a.shared:
        p = greb_b200.default_physics()
    ens.set_member(m, p, np.full(a.years + 1, co2, dtype=np.float32))
ens.init()
if not a.no_spinup:
    ens.spinup(1)
    ms, n = ens.last_kernel_ms()
    print(f"spinup: {ms:.2f} ms for {n} launch(es)")
ens.reset_scenario()
for y in range(a.years):
    ens.run_raw(1)
    ms, n = ens.last_kernel_ms()
    print(f"scenario year {y}: {ms:.2f} ms  -> {a.members / (ms / 1e3):.1f} member-years/s")
print("flags", int(ens.flags().sum()))
ens.close()
