#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference itself (run in the build container only).

The reference source under /root/reference is machine-translated by oracle/f90_to_cpp.py and
compiled into oracle/_ref/ (oracle/ref.py); this script drives that library exactly like the
reference's PROGRAM units do and stores what it produced on the synthetic forcing set S0
(greb_b200/synth.py, seed 20110101 — the generator is bit-reproducible, its digest is stored).
The fixtures are small: per-record SHA-1 digests of every output record, the yearly console
values, and the full fields of a few months.  tests/test_golden.py checks the hand-written
oracle (and, on the GPU box, the CUDA path) against them without needing /root/reference.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))

from greb_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

KEEP_YEARS = (1, 10, 50)   # December of these scenario years is stored in full


def rec_digests(recs: np.ndarray) -> np.ndarray:
    return np.array([hashlib.sha1(np.ascontiguousarray(r, dtype="<f4").tobytes()).hexdigest()[:16] for r in recs])


def pack_run(out: np.ndarray, console, years: int):
    """out: [records][48][96] of one direct-access file holding `years` x 12 x 5 records."""
    d = {"digests": rec_digests(out)}
    for y in KEEP_YEARS:
        if y <= years:
            r0 = ((y - 1) * 12 + 11) * 5
            d[f"dec_year{y}"] = out[r0:r0 + 5].copy()
    d["console"] = np.array(console, dtype=np.float64)
    return d


def kernel_kats(f, R, seed=7):
    """single-subroutine known-answer vectors on seeded inputs (reference argument lists)."""
    rng = np.random.default_rng(seed)
    u, v = f.uclim, f.vclim
    R.array("uclim_m", (730, 48, 96))[:] = np.where(u >= 0, u, 0)
    R.array("uclim_p", (730, 48, 96))[:] = np.where(u >= 0, 0, u)
    R.array("vclim_m", (730, 48, 96))[:] = np.where(v >= 0, v, 0)
    R.array("vclim_p", (730, 48, 96))[:] = np.where(v >= 0, 0, v)
    R.array("dtrad", (730, 48, 96))[:] = (np.float32(-0.16) * f.tclim - np.float32(5.0)).astype(np.float32)
    d = {}
    z = lambda: np.zeros((48, 96), np.float32)
    for case, (ityr, kappa) in enumerate([(11, 8e5), (400, 8e5), (730, 1.2e6), (1, 6.3e5)]):
        R.set_physics(kappa=kappa)
        R.seti("ityr", ityr)
        T = (f.tclim[ityr - 1] + rng.uniform(-2, 2, (48, 96))).astype(np.float32)
        q = (f.qclim[ityr - 1] * rng.uniform(0.5, 1.5, (48, 96))).astype(np.float32)
        q[rng.integers(0, 48, 40), rng.integers(0, 96, 40)] *= 1e-4     # triggers the -0.9*q clamp near the poles
        wza = np.exp(-f.z_topo / np.float32(8400.0)).astype(np.float32)
        wzv = np.exp(-f.z_topo / np.float32(5000.0)).astype(np.float32)
        for nm, X, wz in (("T", T, wza), ("q", q, wzv)):
            dd, aa, cc = z(), z(), z()
            R.call("diffusion", X, dd, 8400.0, wz)
            R.call("advection", X, aa, 8400.0, wz)
            R.call("circulation", X, cc, 8400.0, wz)
            d[f"k{case}_{nm}_in"] = X
            d[f"k{case}_{nm}_wz"] = wz
            d[f"k{case}_{nm}_diffusion"] = dd
            d[f"k{case}_{nm}_advection"] = aa
            d[f"k{case}_{nm}_circulation"] = cc
        d[f"k{case}_meta"] = np.array([ityr, kappa], dtype=np.float64)
    # case 4: mixed-sign anomaly fields — the only inputs on which where(d <= -T) d = -0.9*T (f:715, f:907)
    # fires (on positive fields the stable sub-step never removes more than a cell holds)
    case, ityr, kappa = 4, 213, 8e5
    R.set_physics(kappa=kappa)
    R.seti("ityr", ityr)
    T = rng.normal(0.0, 1.0, (48, 96)).astype(np.float32)
    q = (rng.normal(0.0, 1.0, (48, 96)) * 1e-3).astype(np.float32)
    wza = np.exp(-f.z_topo / np.float32(8400.0)).astype(np.float32)
    wzv = np.exp(-f.z_topo / np.float32(5000.0)).astype(np.float32)
    for nm, X, wz in (("T", T, wza), ("q", q, wzv)):
        dd, aa, cc = z(), z(), z()
        R.call("diffusion", X, dd, 8400.0, wz)
        R.call("advection", X, aa, 8400.0, wz)
        R.call("circulation", X, cc, 8400.0, wz)
        d[f"k{case}_{nm}_in"], d[f"k{case}_{nm}_wz"] = X, wz
        d[f"k{case}_{nm}_diffusion"], d[f"k{case}_{nm}_advection"], d[f"k{case}_{nm}_circulation"] = dd, aa, cc
    d[f"k{case}_meta"] = np.array([ityr, kappa], dtype=np.float64)
    d["n_cases"] = np.array(5)
    R.set_physics(kappa=8e5)
    return d


def main():
    f = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))
    digest = f.digest()

    # ---- kernel-level known answers -----------------------------------------------------------
    R = ref.Ref.fresh("greb")
    R.set_forcing(f)
    kd = kernel_kats(f, R)
    kd["forcing_digest"] = np.array(digest)
    np.savez_compressed(os.path.join(HERE, "ref_kernels.npz"), **kd)
    print("ref_kernels.npz", len(kd))
    if sys.argv[1:] == ["kernels"]:
        return

    # ---- config 1: default namelist (3 yr flux correction + 50 yr at 680 ppm from 1940) ---------
    R = ref.Ref.fresh("greb")
    R.set_forcing(f)
    R.set_run(3, 50, [680.0], year0=1940, ipx=95, ipy=38)
    out = R.greb_model()
    d = pack_run(out, [ln for ln in R.console() if len(ln) == 4], 50)
    d["forcing_digest"] = np.array(digest)
    d["tf_correct_digest"] = rec_digests(R.array("tf_correct", (730, 48, 96))[::73])
    np.savez_compressed(os.path.join(HERE, "ref_config1.npz"), **d)
    print("ref_config1.npz", out.shape, d["console"][:2], d["console"][-1])

    # ---- a perturbed member with a CO2 ramp (2 + 4 yr) ----------------------------------------
    R = ref.Ref.fresh("greb")
    R.set_forcing(f)
    pert = dict(kappa=9.4e5, ct_sens=20.0, a_cloud=0.33, da_ice=0.28, ce=2.2e-3, co_turb=4.5)
    R.set_physics(**pert)
    R.set_run(2, 4, [400.0, 500.0, 600.0], year0=2000, ipx=10, ipy=20)
    out = R.greb_model()
    d = pack_run(out, [ln for ln in R.console() if len(ln) == 4], 4)
    d["dec_year4"] = out[((4 - 1) * 12 + 11) * 5:((4 - 1) * 12 + 11) * 5 + 5].copy()
    d["forcing_digest"] = np.array(digest)
    d["physics"] = np.array(sorted(pert.items()), dtype=object).astype(str)
    np.savez_compressed(os.path.join(HERE, "ref_perturbed.npz"), **d)
    print("ref_perturbed.npz", out.shape, d["console"])

    # ---- config 2: greb-original, log_exp = 10 (3 yr flux + 3 yr control + 50 yr scenario) ----------
    R = ref.Ref.fresh("orig")
    R.set_forcing(f)
    R.seti("time_flux", 3)
    R.seti("time_ctrl", 3)
    R.seti("time_scnr", 50)
    R.seti("log_exp", 10)
    R.seti("ipx", 46)
    R.seti("ipy", 32)
    R.reset_output()
    R.call("greb_model")
    ctrl, scen = R.output_file(21), R.output_file(22)
    d = pack_run(scen, [ln for ln in R.console() if len(ln) == 4], 50)
    d["control_digests"] = rec_digests(ctrl)
    d["control_first_month"] = ctrl[:5].copy()
    d["control_rec_181"] = ctrl[180].copy()      # first TF_correct record that survives the control run
    d["forcing_digest"] = np.array(digest)
    np.savez_compressed(os.path.join(HERE, "ref_config2.npz"), **d)
    print("ref_config2.npz", ctrl.shape, scen.shape, d["console"][:4], d["console"][-1])


if __name__ == "__main__":
    main()
