#!/usr/bin/env python
"""Golden vectors for the two perturbed-physics members on which the two arithmetic modes part company
(tests/test_gpu_long_parity.py: members 22 and 2989 of greb_b200.campaign.perturbed_member, CO2 < 300 ppm, a
bistable sea-ice edge): the TRANSLATED REFERENCE (oracle/_ref, built from /root/reference/src/greb.f90) runs
their 3-year flux correction + 12 scenario years; per-record digests and the console lines are committed as
tests/golden/ref_members.npz.  The oracle must reproduce them bit for bit (tests/test_golden.py), the GPU's exact
mode reproduces the oracle bit for bit over 3 + 50 years (tests/test_gpu_long_parity.py) — which ties the exact
mode to the reference itself exactly where last-ulp differences matter most.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_members.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))

from greb_b200 import campaign, synth  # noqa: E402
from oracle import ref  # noqa: E402
from make_golden import pack_run  # noqa: E402

MEMBERS = (22, 2989)
PERTURBED = ("kappa", "ct_sens", "ce", "co_turb", "a_cloud", "da_ice")
SPINUP, YEARS = 3, 12


def main():
    f = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))
    d = {"forcing_digest": np.array(f.digest()), "members": np.array(MEMBERS), "spinup": np.array(SPINUP),
         "years": np.array(YEARS)}
    for g in MEMBERS:
        p, co2 = campaign.perturbed_member(g)
        R = ref.Ref.fresh("greb")
        R.set_forcing(f)
        R.set_physics(**{k: getattr(p, k) for k in PERTURBED})
        R.set_run(SPINUP, YEARS, [co2], year0=1940)
        out = R.greb_model()
        r = pack_run(out, [ln for ln in R.console() if len(ln) == 4], YEARS)
        for k, v in r.items():
            d[f"m{g}_{k}"] = v
        d[f"m{g}_co2"] = np.array(co2)
        print(g, co2, out.shape, r["console"][-1])
    np.savez_compressed(os.path.join(HERE, "ref_members.npz"), **d)


if __name__ == "__main__":
    main()
