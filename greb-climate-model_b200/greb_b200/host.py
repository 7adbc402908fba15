"""Host front-end of the reference, restated for batched runs.

What `PROGRAM greb_run` does for ONE member (reference src/greb.f90:996-1098) this module does
for a list of namelists at once: read the four namelist groups in the reference's order, pad
`co2_ppm` (f:1047-1061), build `output_file // '_' // ens_id` (f:1063-1068), read the ten
direct-access input files (f:1018-1027, 1073-1085), run all members as ONE ensemble on the GPU
through the C ABI, and write each member's `output/scenario[_ens_id]` byte-compatible with the
reference (5 records per month, f:978-982) plus the yearly console line (f:954).

`read_greb` is a port of the reference's reader R/functions.R:34-81, the normative description
of the output layout.

Nothing here computes physics: the stepping runs on the B200 (`Ensemble`); without the CUDA
library / a GPU `run_namelists` raises.
"""
from __future__ import annotations

import dataclasses
import os
import re
from typing import Dict, List, Sequence

import numpy as np

from . import lib as _lib
from . import synth

XD, YD, NT = 96, 48, 730
VARNAMES = ("tsurf", "tair", "tocean", "vapor", "albedo")   # record order of f:978-982 / R/functions.R:23-32


# ------------------------------------------------------------------------------------------------
# namelists
# ------------------------------------------------------------------------------------------------
class NamelistError(ValueError):
    pass


def _strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def _scalar(tok: str):
    t = tok.strip()
    if not t:
        raise NamelistError("empty value")
    if t[0] in "'\"":
        if len(t) < 2 or t[-1] != t[0]:
            raise NamelistError(f"unterminated string {t!r}")
        return t[1:-1]
    tl = t.lower()
    if tl in (".true.", "t", ".t."):
        return True
    if tl in (".false.", "f", ".f."):
        return False
    if re.fullmatch(r"[+-]?\d+", t):
        return int(t)
    try:
        return float(tl.replace("d", "e"))
    except ValueError:
        raise NamelistError(f"cannot parse value {t!r}") from None


def _values(text: str):
    text = text.strip().rstrip(",").strip()
    if text.startswith("(/") and text.endswith("/)"):
        text = text[2:-2]
    parts, cur, q = [], [], None
    for ch in text:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "," or (ch.isspace() and cur and "".join(cur).strip()):
            if "".join(cur).strip():
                parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur))
    vals = []
    for p in parts:
        m = re.fullmatch(r"\s*(\d+)\*(.+)", p)          # repeat count r*c
        if m:
            vals.extend([_scalar(m.group(2))] * int(m.group(1)))
        else:
            vals.append(_scalar(p))
    return vals


def parse_namelist(text: str) -> Dict[str, Dict[str, list]]:
    """{group: {name: [values]}} — group and variable names lower-cased, `! comments`, `&G ... /`
    (or `&end`), `a = 1, 2, 3`, `(/ .. /)`, repeat counts, quoted strings (which may contain `=`),
    empty groups, subscripted assignments `a(3) = 5, 6` (elements 3 and 4; unassigned positions of the
    returned list are None)."""
    body = "\n".join(_strip_comment(ln) for ln in text.splitlines())
    groups: Dict[str, Dict[str, list]] = {}
    pos = 0
    while True:
        m = re.search(r"&\s*([A-Za-z_]\w*)", body[pos:])
        if not m:
            break
        name = m.group(1).lower()
        start = pos + m.end()
        end = None
        q = None
        for i in range(start, len(body)):
            ch = body[i]
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "/" and not (body[i - 1:i + 1] == "(/" or body[i:i + 2] == "/)"):
                end = i
                break
            elif body[i:i + 4].lower() == "&end":
                end = i
                break
        if end is None:
            raise NamelistError(f"namelist group &{name} is not terminated by '/'")
        content = body[start:end]
        # text inside quotes can hold anything ("a = b", "x(1)"): blank it for the pair search only
        masked, q = [], None
        for ch in content:
            if q:
                masked.append(q if ch == q else "_")
                if ch == q:
                    q = None
            else:
                if ch in "'\"":
                    q = ch
                masked.append(ch)
        masked = "".join(masked)
        # split "name = values" / "name(i) = values" pairs: a new pair starts at an identifier followed by '='
        pairs = list(re.finditer(r"([A-Za-z_]\w*)\s*(?:\(\s*(\d+)\s*\))?\s*=", masked))
        slots: Dict[str, Dict[int, object]] = {}
        for j, pm in enumerate(pairs):
            v_end = pairs[j + 1].start() if j + 1 < len(pairs) else len(content)
            vals = _values(content[pm.end():v_end])
            first = int(pm.group(2)) - 1 if pm.group(2) else 0         # name(i) = v1, v2 fills i, i+1, ...
            if first < 0:
                raise NamelistError(f"{pm.group(1)}({pm.group(2)}): subscripts start at 1")
            d = slots.setdefault(pm.group(1).lower(), {})
            for k, v in enumerate(vals):
                d[first + k] = v
        # positions never assigned stay None (the variable keeps its previous value there)
        items: Dict[str, list] = {n: [d.get(i) for i in range(max(d) + 1)] if d else [] for n, d in slots.items()}
        if name in groups:
            raise NamelistError(f"namelist group &{name} appears twice")
        groups[name] = items
        pos = end + 1
    return groups


@dataclasses.dataclass
class RunConfig:
    """One `./greb <namelist>` invocation (f:1042-1068)."""
    physics: "_lib.Physics"
    time_flux: int = 0
    time_scnr: int = 0
    year0: int = 1940
    ipx: int = 1
    ipy: int = 1
    output_file: str = "output/scenario"
    ens_id: str = ""
    co2_ppm: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, dtype=np.float32))

    @property
    def output_file_full(self) -> str:                                  # f:1063-1068
        return self.output_file if not self.ens_id.strip() else f"{self.output_file.strip()}_{self.ens_id.strip()}"


_PHYS = {n.lower(): n for n in _lib.PHYS_FIELDS}


def pad_co2(given: Sequence[float], n_years: int) -> np.ndarray:
    """f:1047-1061: allocate(time_scnr) = -1; first < 0 -> 680; the tail repeats the last value."""
    co2 = np.full(n_years, -1.0, dtype=np.float32)
    g = np.asarray(list(given), dtype=np.float32)[:n_years]
    co2[:len(g)] = g
    if n_years > 0 and co2[0] == -1:
        co2[0] = 680.0
    for i in range(1, n_years):
        if co2[i] < 0:
            co2[i:] = co2[i - 1]
            break
    return co2


def _only(name: str, v: list):
    """the value of a scalar namelist variable (a subscript or several values are errors in Fortran too)"""
    if len(v) != 1 or v[0] is None:
        raise NamelistError(f"{name} is a scalar: exactly one value, no subscript")
    return v[0]


def config_from_namelist(text: str, defaults: "_lib.Physics | None" = None) -> RunConfig:
    """The four groups of greb.f90 (`physics_par`, `numerics_par`, `diagnostics_par`, `co2_par`); a
    missing group keeps its defaults (the reference requires all four to be present, f:1042-1050)."""
    g = parse_namelist(text)
    known = {"physics_par", "numerics_par", "diagnostics_par", "co2_par"}
    extra = set(g) - known
    if extra:
        raise NamelistError(f"unknown namelist group(s): {sorted(extra)}")
    p = defaults.copy() if defaults is not None else _default_physics_struct()
    for k, v in g.get("physics_par", {}).items():
        if k == "p_emi":
            if len(v) > 10:
                raise NamelistError("p_emi has 10 elements")
            for i, x in enumerate(v):
                if x is not None:
                    p.p_emi[i] = float(x)
        elif k in _PHYS:
            setattr(p, _PHYS[k], float(_only(k, v)))
        else:
            raise NamelistError(f"physics_par: unknown variable {k}")
    cfg = RunConfig(physics=p)
    for k, v in g.get("numerics_par", {}).items():
        if k not in ("ipx", "ipy", "time_flux", "time_scnr", "year0"):
            raise NamelistError(f"numerics_par: unknown variable {k}")
        setattr(cfg, k, int(_only(k, v)))
    for k, v in g.get("diagnostics_par", {}).items():
        if k not in ("output_file", "ens_id"):
            raise NamelistError(f"diagnostics_par: unknown variable {k}")
        setattr(cfg, k, str(_only(k, v)))
    given = []
    for k, v in g.get("co2_par", {}).items():
        if k == "co2_flux":
            p.co2_flux = float(_only(k, v))
        elif k == "co2_ppm":
            given = [-1.0 if x is None else float(x) for x in v]   # co2_ppm(:) = -1 before the read, f:1047
        else:
            raise NamelistError(f"co2_par: unknown variable {k}")
    if len(given) > max(cfg.time_scnr, 0):
        # gfortran aborts when more values than allocated elements are supplied
        raise NamelistError(f"co2_ppm has {len(given)} values but time_scnr = {cfg.time_scnr}")
    cfg.co2_ppm = pad_co2(given, cfg.time_scnr)
    return cfg


def _default_physics_struct() -> "_lib.Physics":
    """Reference defaults f:68-104 without touching the CUDA library (so that namelists parse on any
    host); tests/test_host_formats.py checks them against greb_b200_physics_defaults."""
    p = _lib.Physics()
    f32 = np.float32
    vals = dict(pi=3.1416, sig=5.6704e-8, rho_ocean=999.1, rho_land=2600., rho_air=1.2, cp_ocean=4186.,
                cp_land=926.222, cp_air=1005., eps=1., d_ocean=50., d_land=2., d_air=5000., ct_sens=22.5,
                da_ice=0.25, a_no_ice=0.1, a_cloud=0.35, Tl_ice1=float(f32(273.15) - f32(10.)), Tl_ice2=273.15,
                To_ice1=float(f32(273.15) - f32(7.)), To_ice2=float(f32(273.15) - f32(1.7)), co_turb=5.0, kappa=8e5,
                ce=2e-3, cq_latent=2.257e6, cq_rain=float(f32(f32(-0.1) / f32(24.)) / f32(3600.)), z_air=8400.,
                z_vapor=5000., r_qviwv=2.6736e3)
    for k, v in vals.items():
        setattr(p, k, v)
    for i, x in enumerate((9.0721, 106.7252, 61.5562, 0.0179, 0.0028, 0.0570, 0.3462, 2.3406, 0.7032, 1.0662)):
        p.p_emi[i] = x
    p.co2_flux = 298.
    return p


# ------------------------------------------------------------------------------------------------
# files
# ------------------------------------------------------------------------------------------------
def read_inputs(directory: str = "input") -> "synth.Forcing":
    """The ten direct-access files of f:1018-1027 (RECL = 4*xdim*ydim bytes, little-endian fp32)."""
    missing = [n for n in synth.INPUT_FILES if not os.path.exists(os.path.join(directory, n))]
    if missing:
        raise FileNotFoundError(f"{directory}: missing input file(s) {missing}")
    return synth.Forcing.read(directory)


def write_output(path: str, monthly: np.ndarray) -> None:
    """`monthly` [years][12][5][48][96] -> the reference's record stream (f:978-982).  Like the
    reference (opened without status='replace', f:174) a longer existing file is not truncated."""
    a = np.ascontiguousarray(monthly, dtype="<f4")
    if a.shape[-3:] != (5, YD, XD):
        raise ValueError("monthly must end in [5][48][96]")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    mode = "r+b" if os.path.exists(path) else "wb"
    with open(path, mode) as fh:
        fh.seek(0)
        fh.write(a.tobytes())


def read_greb(file: str, tstamps=None, varname: Sequence[str] = VARNAMES, ivar: Sequence[int] | None = None,
              nvar: int | None = None, nlon: int = XD, nlat: int = YD, nbyte: int = 4):
    """Port of R/functions.R:34-81.  Returns dict(time, variable, lon, lat, value) where value has
    shape [ntime][len(ivar)][nlat][nlon] (lon fastest in the file, R/functions.R:45-51)."""
    varname = list(varname)
    ivar = list(range(1, len(varname) + 1)) if ivar is None else list(ivar)
    nvar = max(ivar) if nvar is None else nvar
    if len(ivar) != len(varname):
        raise ValueError("ivar and varname must have the same length")
    size = os.path.getsize(file)
    ngrid = nlon * nlat
    ntime, rem = divmod(size, ngrid * nvar * nbyte)
    if rem != 0:                                                    # R/functions.R:41
        raise ValueError(f"{file}: size {size} is not a multiple of nlon*nlat*nvar*nbyte = {ngrid * nvar * nbyte}")
    tstamps = list(range(1, ntime + 1)) if tstamps is None else list(tstamps)
    if len(tstamps) != ntime:
        raise ValueError("length(tstamps) != ntime")
    dlon, dlat = 360.0 / nlon, 180.0 / nlat
    lon = dlon / 2 + dlon * np.arange(nlon)                         # 1.875 ... 358.125
    lat = -90.0 + dlat / 2 + dlat * np.arange(nlat)                 # -88.125 ... 88.125
    data = np.fromfile(file, dtype="<f4").reshape(ntime, nvar, nlat, nlon)
    value = data[:, [i - 1 for i in ivar]]
    return {"time": tstamps, "variable": varname, "lon": lon, "lat": lat, "value": value}


def console_line(year: float, co2: float, gmean: float, point: float) -> str:
    """The yearly line of f:954 (list-directed `print *` of four reals)."""
    return f"   {year:12.6f}   {co2:12.6f}   {gmean:12.8f}   {point:12.8f}"


# ------------------------------------------------------------------------------------------------
# batched driver: N namelists -> one ensemble on the GPU
# ------------------------------------------------------------------------------------------------
def run_namelists(namelists: Sequence[str], input_dir: str = "input", workdir: str = ".", device: int = 0,
                  write_files: bool = True, verbose: bool = False):
    """`./greb nml_1`, `./greb nml_2`, ... as ONE batch.  All members must share `time_flux` and
    `time_scnr` (one launch advances every member by the same number of steps).  Returns a list of
    dicts(config, monthly [years][12][5][48][96], gmean [years], point [years], lines)."""
    cfgs = []
    for n in namelists:
        with open(n) as fh:
            cfgs.append(config_from_namelist(fh.read()))
    if not cfgs:
        return []
    tf, ts = cfgs[0].time_flux, cfgs[0].time_scnr
    if any(c.time_flux != tf or c.time_scnr != ts for c in cfgs):
        raise ValueError("run_namelists: all members of a batch must share time_flux and time_scnr")
    forcing = read_inputs(input_dir)
    ens = _lib.Ensemble(len(cfgs), device=device)
    try:
        ens.set_forcing(forcing)
        for m, c in enumerate(cfgs):
            ens.set_member(m, c.physics, c.co2_ppm if ts > 0 else [680.0], year0=c.year0)
        ens.init()
        ens.spinup(tf)                                              # f:221
        ens.reset_scenario()                                        # f:226-227
        results = [dict(config=c, monthly=np.zeros((ts, 12, 5, YD, XD), np.float32), gmean=np.zeros(ts, np.float32),
                        point=np.zeros(ts, np.float32), lines=[]) for c in cfgs]
        for y in range(ts):                                         # one simulated year per launch
            out, gm, _ = ens.run(1)
            for m, (c, r) in enumerate(zip(cfgs, results)):
                r["monthly"][y] = out[m, 0]
                r["gmean"][y] = gm[m, 0]
                # f:954 prints tsmn(ipx,ipy)-273.15, the annual mean of Tsurf at the diagnostic point; the
                # ABI returns monthly means, so the point value is their day-weighted mean (not bit-exact)
                w = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], dtype=np.float64)
                ts_pt = out[m, 0, :, 0, c.ipy - 1, c.ipx - 1].astype(np.float64)
                r["point"][y] = np.float32((ts_pt * w).sum() / w.sum() - 273.15)
                r["lines"].append(console_line(c.year0 + y, float(c.co2_ppm[y]), float(gm[m, 0]), float(r["point"][y])))
                if verbose:
                    print(r["lines"][-1])
        if write_files:
            for r in results:
                write_output(os.path.join(workdir, r["config"].output_file_full), r["monthly"])
        return results
    finally:
        ens.close()


# ------------------------------------------------------------------------------------------------
# greb.original.model.f90: the `log_exp` sensitivity experiments (namelist_original)
# ------------------------------------------------------------------------------------------------
# log_exp -> process switches of include/greb_b200.h.  The original model accumulates its switches
# with `log_exp <= k` tests, so experiment L switches off everything of the experiments above it.
# log_exp <= 4, 7 and 16: `circulation` returns before assigning its intent(out) result (orig:553-555), so
# time_loop adds whatever its local arrays hold.  The C ABI DEFINES the unassigned result as zero
# (GREB_SW_NO_HEAT_CIRCULATION / GREB_SW_NO_VAPOR_CIRCULATION) — the evidently intended "no circulation", and what
# the reference computes with zero-initialised static locals.
_SW = _lib
_NOCRCL = _SW.SW_NO_HEAT_CIRCULATION | _SW.SW_NO_VAPOR_CIRCULATION
_EBM = _SW.SW_NO_ICE_ALBEDO | _SW.SW_NO_HYDRO | _SW.SW_NO_DEEP_OCEAN | _NOCRCL   # orig:394, 453, 492, 514, 553
ORIGINAL_EXPERIMENTS = {
    1: _EBM,                                                                 # + constant topography, clouds, vapour (orig:162-164)
    2: _EBM,                                                                 # + constant clouds, vapour
    3: _EBM,                                                                 # + constant vapour
    4: _EBM,
    7: _SW.SW_NO_VAPOR_CIRCULATION | _SW.SW_NO_DEEP_OCEAN,                   # orig:554, 514
    16: _SW.SW_SST_PLUS_1K | _SW.SW_NO_DEEP_OCEAN | _SW.SW_NO_VAPOR_CIRCULATION,   # orig:225-226, 515, 555
    5: _SW.SW_NO_ICE_ALBEDO | _SW.SW_NO_HYDRO | _SW.SW_NO_DEEP_OCEAN,       # orig:394, 453, 492, 514 (+ mld = d_ocean :165)
    6: _SW.SW_NO_HYDRO | _SW.SW_NO_DEEP_OCEAN,                               # orig:453, 514
    8: _SW.SW_VAPOR_DIFFUSION_ONLY | _SW.SW_NO_DEEP_OCEAN,                   # orig:560-564, 514
    9: _SW.SW_NO_DEEP_OCEAN,                                                 # orig:165, 514
    10: 0,                                                                   # the full model
    11: _SW.SW_LINEAR_VAPOR_EMISSIVITY | _SW.SW_NO_DEEP_OCEAN,               # orig:166, 423, 430, 514
    12: 0,                                                                   # A1B CO2 pathway, orig:179, 946-951
    13: _SW.SW_NO_HYDRO,                                                     # A1B without hydro, orig:453
    14: _SW.SW_SST_PLUS_1K | _SW.SW_NO_DEEP_OCEAN,                           # orig:225-226, 515
    15: _SW.SW_SST_PLUS_1K | _SW.SW_NO_DEEP_OCEAN | _SW.SW_NO_HYDRO,         # + orig:453
}


def a1b_co2(year: float) -> np.float32:
    """co2_level of greb.original.model.f90:945-951 for log_exp 12/13 (fp32, `year` is a real)."""
    y = np.float32(year)
    co2 = np.float32(680.0)
    if y <= np.float32(2000.0):
        co2 = np.float32(310.0) + np.float32(np.float32(60.0) / np.float32(50.0)) * (y - np.float32(1950.0))
    if np.float32(2000.0) < y <= np.float32(2050.0):
        co2 = np.float32(370.0) + np.float32(np.float32(150.0) / np.float32(50.0)) * (y - np.float32(2000.0))
    if np.float32(2050.0) < y <= np.float32(2100.0):
        co2 = np.float32(520.0) + np.float32(np.float32(180.0) / np.float32(50.0)) * (y - np.float32(2050.0))
    return np.float32(co2)


def original_experiment(log_exp: int, forcing: "synth.Forcing", time_scnr: int, d_ocean: float = 50.0) -> dict:
    """What `log_exp` of greb.original.model.f90 means for the C ABI: the switch mask, the modified
    inputs (orig:162-166), CO2_ctrl (orig:178-179) and the scenario CO2 path (orig:225, 939-951;
    the scenario starts in 1940, orig:219)."""
    if log_exp not in ORIGINAL_EXPERIMENTS:
        raise ValueError(f"log_exp = {log_exp}: not an experiment of greb.original.model.f90 "
                         f"({sorted(ORIGINAL_EXPERIMENTS)})")
    f = forcing
    if log_exp == 1:                                                 # orig:162 constant topography
        f = dataclasses.replace(f, z_topo=np.where(f.z_topo > 1.0, np.float32(1.0), f.z_topo).astype(np.float32))
    if log_exp <= 2:                                                 # orig:163 constant cloud cover
        f = dataclasses.replace(f, cldclim=np.full_like(f.cldclim, np.float32(0.7)))
    if log_exp <= 3:                                                 # orig:164 constant water vapour
        f = dataclasses.replace(f, qclim=np.full_like(f.qclim, np.float32(0.0052)))
    if log_exp <= 9 or log_exp == 11:                                # orig:165-166 "no deep ocean"
        f = dataclasses.replace(f, mldclim=np.full_like(f.mldclim, np.float32(d_ocean)))
    co2_ctrl = 298.0 if log_exp in (12, 13) else 340.0
    if log_exp in (12, 13):
        path = np.array([a1b_co2(1940 + y) for y in range(time_scnr)], dtype=np.float32)
    elif 14 <= log_exp <= 16:
        path = np.full(time_scnr, co2_ctrl, dtype=np.float32)        # orig:225
    else:
        path = np.full(time_scnr, 680.0, dtype=np.float32)           # orig:943
    return {"switches": ORIGINAL_EXPERIMENTS[log_exp], "forcing": f, "co2_ctrl": co2_ctrl, "co2_scenario": path}


def run_original(log_exp: int, forcing: "synth.Forcing", time_flux: int = 3, time_ctrl: int = 3, time_scnr: int = 50,
                 device: int = 0, arith: str = "exact") -> dict:
    """greb.original.model.f90:138-233 through the C ABI: flux-correction spin-up and control run
    at CO2_ctrl, then the scenario from the same initial fields (cap_surf carries over, orig:219).
    Returns the record streams of units 21 and 22 as [years][12][5][48][96] plus the annual means."""
    ex = original_experiment(log_exp, forcing, time_scnr)
    ens = _lib.Ensemble(1, device=device)
    try:
        ens.set_arithmetic(arith)
        ens.set_forcing(ex["forcing"])
        p = _lib.original_physics()
        p.co2_flux = ex["co2_ctrl"]
        co2 = np.concatenate([np.full(time_ctrl, ex["co2_ctrl"], dtype=np.float32), ex["co2_scenario"]])
        ens.set_member(0, p, co2 if len(co2) else [680.0], year0=1970)
        ens.set_switches(0, ex["switches"] & ~_lib.SW_SST_PLUS_1K)   # orig:226 applies to the scenario only
        ens.init()
        ens.spinup(time_flux)                                        # orig:201
        tf = ens.get_fluxcorr(0, 0)                                  # orig:204-206
        ini = {n: ens.get_state(0, n) for n in ("Ts", "Ta", "To", "q")}
        ens.reset_scenario()
        ctrl, gmc, _ = ens.run(time_ctrl)                            # orig:209-215
        for n, a in ini.items():                                     # orig:219
            ens.set_state(0, n, a)
        ens.set_switches(0, ex["switches"])
        scen, gm, gmw = ens.run(time_scnr)                           # orig:220-231
        return {"control": ctrl[0], "scenario": scen[0], "gmean_control": gmc[0], "gmean": gm[0],
                "gmean_coslat": gmw[0], "tf_correct": tf, "flags": ens.flags(), "experiment": ex}
    finally:
        ens.close()


def write_control(path: str, tf_correct: np.ndarray, control_monthly: np.ndarray) -> None:
    """`output/control` of greb.original.model.f90: 730 records of TF_correct (orig:204-206), then the
    control run's monthly means written over them from record 1 (orig:209-215, unit 21)."""
    recs = np.ascontiguousarray(tf_correct, dtype="<f4").reshape(-1, YD, XD)
    mon = np.ascontiguousarray(control_monthly, dtype="<f4").reshape(-1, YD, XD)
    n = max(len(recs), len(mon))
    out = np.zeros((n, YD, XD), dtype="<f4")
    out[:len(recs)] = recs
    out[:len(mon)] = mon
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    out.tofile(path)


def run_original_namelist(namelist: str = "namelist_original", input_dir: str = "input", workdir: str = ".",
                          device: int = 0, arith: str = "exact", write_files: bool = True) -> dict:
    """`./greb_original` work-alike (greb.original.shell.web-public.f90:16-60): reads the NUMERICS and
    PHYSICS groups of `namelist_original` and the ten input files, runs spin-up, control and
    scenario on the GPU and writes `output/control` and `output/scenario`."""
    with open(namelist) as fh:
        groups = parse_namelist(fh.read())
    num = {k.lower(): v for k, v in groups.get("numerics", {}).items()}
    phy = {k.lower(): v for k, v in groups.get("physics", {}).items()}
    tf = int(num.get("time_flux", [0])[0])
    tc = int(num.get("time_ctrl", [0])[0])
    ts = int(num.get("time_scnr", [0])[0])
    log_exp = int(phy.get("log_exp", [0])[0])                       # orig:60 default 0
    r = run_original(log_exp, read_inputs(input_dir), tf, tc, ts, device=device, arith=arith)
    if write_files:
        write_control(os.path.join(workdir, "output", "control"), r["tf_correct"], r["control"])
        write_output(os.path.join(workdir, "output", "scenario"), r["scenario"])
    return r


# ------------------------------------------------------------------------------------------------
# spin-up cache on disk: qflux_correction (f:311-364) once per (inputs, physics, years)
# ------------------------------------------------------------------------------------------------
def spinup_key(forcing: "synth.Forcing", physics: "_lib.Physics", time_flux: int, switches: int = 0,
               arith: str = "exact") -> str:
    """Everything the flux-correction spin-up depends on: the ten input fields, physics_par incl.
    co2_flux (the spin-up CO2, f:221), the spin-up-relevant process switches, its length, and the
    arithmetic mode (exact and fast corrections differ in the last bits)."""
    import ctypes
    import hashlib
    h = hashlib.sha256()
    h.update(forcing.digest().encode())
    h.update(ctypes.string_at(ctypes.byref(physics), ctypes.sizeof(physics)))
    h.update(f"|{int(time_flux)}|{int(switches) & ~_lib.SW_SST_PLUS_1K}|{arith}".encode())
    return h.hexdigest()[:32]


def save_spinup(path: str, ens: "_lib.Ensemble", member: int, key: str) -> None:
    """Write member `member`'s post-spin-up state (Ts, Ta, To, q, cap_surf) and the three flux
    correction fields of its physics group (TF, qF, ToF; 40.4 MB) to `path` (.npz)."""
    d = {"key": np.array(key), "state": np.stack([ens.get_state(member, n) for n in _lib.STATE])}
    for w, name in enumerate(("tf_correct", "qf_correct", "tof_correct")):
        d[name] = ens.get_fluxcorr(member, w)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **d)
    os.replace(tmp, path)


def load_spinup(path: str, ens: "_lib.Ensemble", members: Sequence[int], key: str) -> bool:
    """Restore a saved spin-up into `members` (all of one physics group: the corrections are set
    through the first of them) of an initialised handle instead of calling `spinup`.  Returns False
    — and touches nothing — if the file is missing or was written for another key."""
    if not os.path.exists(path):
        return False
    with np.load(path) as z:
        if str(z["key"]) != key:
            return False
        state = z["state"]
        corr = [z[n] for n in ("tf_correct", "qf_correct", "tof_correct")]
    for w, a in enumerate(corr):
        ens.set_fluxcorr(members[0], w, a)
    for m in members:
        for i, n in enumerate(_lib.STATE):
            ens.set_state(m, n, state[i])
    return True


# ------------------------------------------------------------------------------------------------
# checkpoint / resume of a scenario (loop state of f:226-234)
# ------------------------------------------------------------------------------------------------
def save_checkpoint(path: str, ens: "_lib.Ensemble", with_fluxcorr: bool = True) -> None:
    """Everything `greb_model`'s scenario loop carries from one step to the next, for every member of
    the handle: Ts1,Ta1,To1,q1,cap_surf, the step counter `it` of the next step (mon, year, irec, ityr,
    jday are functions of it, f:241-252, 975-985), the monthly accumulators and tsmn (f:145-149) — plus,
    unless `with_fluxcorr` is False, the three flux-correction fields of every physics group (40.4 MB
    each; a member's corrections are read through the member itself).  Written atomically as .npz."""
    d = {"version": np.array(1), "n_members": np.array(ens.n), "it_next": np.array(ens.get_calendar()),
         "state": ens.get_states(), "acc": ens.get_accumulators()}
    if with_fluxcorr:
        for w, name in enumerate(("tf_correct", "qf_correct", "tof_correct")):
            d[name] = np.stack([ens.get_fluxcorr(m, w) for m in range(ens.n)])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **d)
    os.replace(tmp, path)


def load_checkpoint(path: str, ens: "_lib.Ensemble") -> int:
    """Restore a checkpoint into an INITIALISED handle with the same members (same set_member calls:
    physics and CO2 paths are configuration, not state).  Returns the step counter `it` of the next step;
    continue with `ens.run(years)` (at a year boundary) or `ens.time_steps(it, n)` to the next one."""
    with np.load(path) as z:
        if int(z["version"]) != 1 or int(z["n_members"]) != ens.n:
            raise ValueError(f"{path}: checkpoint of {int(z['n_members'])} members, handle has {ens.n}")
        ens.set_states(z["state"])
        ens.set_accumulators(z["acc"])
        if "tf_correct" in z.files:
            for w, name in enumerate(("tf_correct", "qf_correct", "tof_correct")):
                a = z[name]
                for m in range(ens.n):
                    ens.set_fluxcorr(m, w, a[m])
        it = int(z["it_next"])
    ens.set_calendar(it)
    return it
