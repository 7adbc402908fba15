#!/usr/bin/env python
"""Golden vectors on the REAL orography (run in the build container only; needs /root/reference).

The mount carries three of the reference's ten input files: input/topography, input/glacier.masks and
input/solar.radiation.  This script takes them as they are, generates the seven missing climatologies
around that orography (greb_b200.synth.make_forcing(topo="reference"), bit-reproducible), runs the
reference itself — src/greb.f90 machine-translated and compiled, oracle/ref.py — for 1 flux-correction
year + 2 scenario years at 680 ppm, and stores

  * the three input fields (so that the GPU box, where /root/reference does not exist, runs the CUDA path
    on exactly these inputs),
  * a SHA-1 digest of every one of the 120 output records, the console values, the December fields of
    year 2 and digests of the three flux-correction fields.

    python tests/golden/make_golden_realtopo.py        # writes tests/golden/ref_realtopo.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))
sys.path.insert(0, HERE)

from greb_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402
from make_golden import rec_digests  # noqa: E402

INPUT = os.path.join(ref.REFERENCE_ROOT, "input")


def main():
    f = synth.make_forcing(topo="reference", reference_input=INPUT)
    R = ref.Ref.fresh("greb")
    R.set_forcing(f)
    R.set_run(1, 2, [680.0], year0=1940, ipx=46, ipy=32)
    out = R.greb_model()                                    # [120][48][96]
    d = {"z_topo": f.z_topo, "glacier": f.glacier, "sw_solar": f.sw_solar, "forcing_digest": np.array(f.digest()),
         "digests": rec_digests(out), "dec_year2": out[(12 + 11) * 5:(12 + 11) * 5 + 5].copy(),
         "console": np.array([ln for ln in R.console() if len(ln) == 4], dtype=np.float64)}
    for name in ("tf_correct", "qf_correct", "tof_correct"):
        d[name + "_digest"] = rec_digests(R.array(name, (730, 48, 96))[::73])
    np.savez_compressed(os.path.join(HERE, "ref_realtopo.npz"), **d)
    print("ref_realtopo.npz", out.shape, d["console"], os.path.getsize(os.path.join(HERE, "ref_realtopo.npz")))


if __name__ == "__main__":
    main()
