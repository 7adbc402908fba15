#!/usr/bin/env python
"""Golden vectors for the `log_exp` sensitivity experiments of src/greb.original.model.f90.

Same method as make_golden.py: the reference source is machine-translated (oracle/f90_to_cpp.py)
and compiled into oracle/_ref/, and `greb_model` of that library is run on the synthetic forcing
S0 for every experiment whose result the reference defines (log_exp 5, 6, 8-15; for log_exp <= 4,
7 and 16 `circulation` returns without assigning its result, orig:553-555).

  long_L   time_flux = time_ctrl = 1, time_scnr = 2: December of scenario year 2 (5 records),
           December of the control year, the console values           -> GPU tests
  steps_L  the reference's `time_loop` called directly for NSTEPS steps after the `greb_model`
           preamble (time_flux = time_ctrl = time_scnr = 0: setup only, zero flux corrections),
           the two scenario-loop lines orig:225-226 restated here for log_exp 14/15 with the
           module variable ityr = 730 as a control run leaves it (orig:226 reads Tclim(:,:,ityr)
           BEFORE time_loop updates ityr); final Ts, Ta, To, q, cap_surf -> CPU tests (emulator)

    python tests/golden/make_golden_experiments.py      # writes tests/golden/ref_original_experiments.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))

from greb_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

import ctypes as C  # noqa: E402

LONG = (5, 6, 8, 9, 11, 12, 13, 14, 15)
# log_exp <= 4, 7, 16: `circulation` returns without assigning its intent(out) result (orig:553-555).  In the
# translated reference the locals dTa_crcl / dq_crcl of time_loop and qflux_correction are zero-initialised
# static arrays that nothing ever writes in these experiments, i.e. it computes "dX_crcl = 0" — the definition
# the C ABI adopts (GREB_SW_NO_HEAT_CIRCULATION / GREB_SW_NO_VAPOR_CIRCULATION).
NOCRCL = (1, 2, 3, 4, 7, 16)
NSTEPS = 40


def run(f, log_exp, tf, tc, ts):
    R = ref.Ref.fresh("orig")
    R.set_forcing(f)
    R.seti("time_flux", tf)
    R.seti("time_ctrl", tc)
    R.seti("time_scnr", ts)
    R.seti("log_exp", log_exp)
    R.seti("ipx", 46)
    R.seti("ipy", 32)
    R.reset_output()
    R.call("greb_model")
    return R.output_file(21), R.output_file(22), [ln for ln in R.console() if len(ln) == 4]


def steps(f, log_exp):
    R = ref.Ref.fresh("orig")
    R.set_forcing(f)
    for n in ("time_flux", "time_ctrl", "time_scnr"):
        R.seti(n, 0)
    R.seti("log_exp", log_exp)
    R.seti("ipx", 46)
    R.seti("ipy", 32)
    R.reset_output()
    R.call("greb_model")                                   # preamble only (orig:138-199)
    R.record_output(False)
    # orig:172-175 (Ts_ini ... are locals of greb_model): step 730 of the (possibly modified) climatologies
    Ts1 = R.array("tclim", (730, 48, 96))[729].copy()
    Ta1 = Ts1.copy()
    To1 = R.array("toclim", (730, 48, 96))[729].copy()
    q1 = R.array("qclim", (730, 48, 96))[729].copy()
    Ts0, Ta0, To0, q0 = (np.zeros((48, 96), np.float32) for _ in range(4))
    ocean = R.array("z_topo", (48, 96)) < 0
    tclim = R.array("tclim", (730, 48, 96))
    irec, mon = C.c_int(0), C.c_int(1)
    R.seti("ityr", 730)
    co2_ctrl = 298.0 if log_exp in (12, 13) else 340.0
    year = C.c_float(1940.0)
    for it in range(1, NSTEPS + 1):
        co2 = C.c_float(0.0)
        R.call("co2_level", it, year, co2)                 # orig:222
        if 14 <= log_exp <= 16:                            # orig:225-226
            co2 = C.c_float(co2_ctrl)
            Ts1[ocean] = tclim[R.geti("ityr") - 1][ocean] + np.float32(1.0)
        R.call("time_loop", it, 0, year, co2, irec, mon, 22, Ts1, Ta1, q1, To1, Ts0, Ta0, q0, To0)
        Ts1[:], Ta1[:], q1[:], To1[:] = Ts0, Ta0, q0, To0
    return np.stack([Ts1, Ta1, To1, q1, R.array("cap_surf", (48, 96)).copy()])


def main():
    f = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))
    d = {}
    path = os.path.join(HERE, "ref_original_experiments.npz")
    todo = LONG + NOCRCL
    if sys.argv[1:] == ["nocrcl"] and os.path.exists(path):      # add the new experiments, keep the rest
        with np.load(path) as z:
            d = {k: z[k] for k in z.files}
        todo = NOCRCL
    for L in todo:
        ctrl, scen, con = run(f, L, 1, 1, 2)
        d[f"long_{L}_scen_dec2"] = scen[(12 + 11) * 5:(12 + 11) * 5 + 5].copy()
        d[f"long_{L}_console"] = np.array(con, dtype=np.float64)
        print("long", L, scen.shape, con[-1])
    for L in todo:
        d[f"steps_{L}_state"] = steps(f, L)
        print("steps", L, float(d[f"steps_{L}_state"][0].mean()))
    d["nsteps"] = np.array(NSTEPS)
    np.savez_compressed(os.path.join(HERE, "ref_original_experiments.npz"), **d)


if __name__ == "__main__":
    main()
