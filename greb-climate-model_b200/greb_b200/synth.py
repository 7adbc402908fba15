"""Synthetic GREB forcing set "S0" in the reference's binary input format.

The reference reads ten raw little-endian fp32 direct-access files from ``input/``
(reference ``src/greb.f90:1018-1027, 1073-1085``).  Seven of them are missing from the
reference mount and none of them exists on the GPU box, so every test and the bench run
on this generator's output.  The generator is *bit-reproducible*: it uses numpy's PCG64
stream plus float64 ``+ - * /`` and ``sqrt`` only (own polynomial sin/cos/exp below), so
the same seed gives byte-identical files on any host; ``forcing_digest`` lets tests pin that.

Layout (reference ``src/greb.f90:108-111``, ``R/functions.R:45-51``): a field is
``real(xdim=96, ydim=48)`` Fortran order == C ``[48][96]`` with longitude fastest;
row 0 is 88.125S, column 0 is 1.875E (cell centres).  Climatologies are ``[730][48][96]``;
``sw_solar`` is ``[730][48]``.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass

import numpy as np

XDIM, YDIM, NSTEP_YR = 96, 48, 730
FIELD_BYTES = 4 * XDIM * YDIM

# file name -> attribute (reference src/greb.f90:1018-1027)
INPUT_FILES = {
    "tsurf": "tclim",
    "vapor": "qclim",
    "topography": "z_topo",
    "soil.moisture": "swetclim",
    "solar.radiation": "sw_solar",
    "zonal.wind": "uclim",
    "meridional.wind": "vclim",
    "ocean.mld": "mldclim",
    "cloud.cover": "cldclim",
    "glacier.masks": "glacier",
}

_PI = 3.141592653589793
_LN2 = 0.6931471805599453


# ---- deterministic elementary functions (float64, basic IEEE ops only) -------------------

def _sin_cos(x):
    """sin and cos by reduction to [-pi/4, pi/4] octants and fixed-order Taylor sums."""
    x = np.asarray(x, dtype=np.float64)
    k = np.floor(x / (_PI / 2) + 0.5)
    r = x - k * (_PI / 2)          # |r| <= pi/4 (+ rounding); exact same ops everywhere
    r2 = r * r
    s = np.zeros_like(r)
    c = np.zeros_like(r)
    # Horner, degree 17/16 — far below fp64 rounding for |r| <= 0.8
    for n in (17, 15, 13, 11, 9, 7, 5, 3):
        s = (s + 1.0) * (-r2 / (n * (n - 1)))
    s = (s + 1.0) * r
    for n in (16, 14, 12, 10, 8, 6, 4, 2):
        c = (c + 1.0) * (-r2 / (n * (n - 1)))
    c = c + 1.0
    q = np.mod(k, 4.0)
    sin = np.where(q == 0, s, np.where(q == 1, c, np.where(q == 2, -s, -c)))
    cos = np.where(q == 0, c, np.where(q == 1, -s, np.where(q == 2, -c, s)))
    return sin, cos


def _sin(x):
    return _sin_cos(x)[0]


def _cos(x):
    return _sin_cos(x)[1]


def _exp(x):
    x = np.asarray(x, dtype=np.float64)
    k = np.floor(x / _LN2 + 0.5)
    r = x - k * _LN2
    p = np.zeros_like(r)
    for n in range(14, 0, -1):
        p = (p + 1.0) * (r / n)
    p = p + 1.0
    return p * np.exp2(k)  # exp2 of an integer-valued float is exact


# ---- container ----------------------------------------------------------------------------

@dataclass
class Forcing:
    """The ten input fields of the reference, all fp32, C order ``[time][lat][lon]``."""

    z_topo: np.ndarray      # [48][96]   m, ocean < 0
    glacier: np.ndarray     # [48][96]   {0,1}
    sw_solar: np.ndarray    # [730][48]  W/m^2
    tclim: np.ndarray       # [730][48][96] K
    qclim: np.ndarray       # kg/kg
    swetclim: np.ndarray    # 0..1
    uclim: np.ndarray       # m/s
    vclim: np.ndarray       # m/s
    mldclim: np.ndarray     # m
    cldclim: np.ndarray     # 0..1

    def fields(self):
        return {name: getattr(self, attr) for name, attr in INPUT_FILES.items()}

    def write(self, directory: str) -> None:
        """Write the ten files exactly as the reference's direct-access reads expect."""
        os.makedirs(directory, exist_ok=True)
        for name, arr in self.fields().items():
            np.ascontiguousarray(arr, dtype="<f4").tofile(os.path.join(directory, name))

    @staticmethod
    def read(directory: str) -> "Forcing":
        kw = {}
        for name, attr in INPUT_FILES.items():
            a = np.fromfile(os.path.join(directory, name), dtype="<f4")
            if attr in ("z_topo", "glacier"):
                a = a.reshape(YDIM, XDIM)
            elif attr == "sw_solar":
                a = a.reshape(NSTEP_YR, YDIM)
            else:
                a = a.reshape(NSTEP_YR, YDIM, XDIM)
            kw[attr] = a
        return Forcing(**kw)

    def digest(self) -> str:
        h = hashlib.sha256()
        for name in sorted(INPUT_FILES):
            h.update(np.ascontiguousarray(getattr(self, INPUT_FILES[name]), dtype="<f4").tobytes())
        return h.hexdigest()


# ---- generator ----------------------------------------------------------------------------

def _grid():
    lat = (np.arange(YDIM, dtype=np.float64) + 0.5) * (180.0 / YDIM) - 90.0   # -88.125 .. 88.125
    lon = (np.arange(XDIM, dtype=np.float64) + 0.5) * (360.0 / XDIM)          # 1.875 .. 358.125
    return lat, lon


def _smooth_noise(rng, ntime_modes=2, nmodes=6):
    """Smooth, zonally periodic random field generator: returns f(day)[48][96] callable.

    Sum of a few low-wavenumber lon/lat harmonics with random amplitudes and phases and an
    annual / semi-annual modulation, normalised to roughly unit variance."""
    lat, lon = _grid()
    phi = lat * (_PI / 180.0)
    lam = lon * (_PI / 180.0)
    comps = []
    for _ in range(nmodes):
        kx = int(rng.integers(1, 6))
        ky = int(rng.integers(1, 5))
        amp = float(rng.random()) + 0.5
        p1, p2, p3 = (float(rng.random()) * 2 * _PI for _ in range(3))
        kt = int(rng.integers(0, ntime_modes + 1))
        comps.append((kx, ky, amp, p1, p2, p3, kt))
    norm = np.sqrt(sum(c[2] ** 2 for c in comps) / 4.0)

    def f(day):
        out = np.zeros((YDIM, XDIM))
        for kx, ky, amp, p1, p2, p3, kt in comps:
            a = _sin(kx * lam + p1)[None, :] * _sin(ky * phi + p2)[:, None]
            out = out + amp * a * float(_cos(np.float64(2 * _PI * kt * day / 365.0 + p3)))
        return out / norm

    return f


def make_topography(rng):
    """Synthetic continents: ocean is exactly -0.1 m (as in the real file), land 1..~4500 m.
    No cell is exactly 0 (the reference's land/ocean predicates disagree at 0, greb.f90:190-191,384,453)."""
    lat, lon = _grid()
    phi = lat * (_PI / 180.0)
    lam = lon * (_PI / 180.0)
    f = np.zeros((YDIM, XDIM))
    for _ in range(10):
        kx = int(rng.integers(1, 5))
        ky = int(rng.integers(1, 5))
        amp = float(rng.random()) + 0.3
        p1 = float(rng.random()) * 2 * _PI
        p2 = float(rng.random()) * 2 * _PI
        f = f + amp * _sin(kx * lam + p1)[None, :] * _cos(ky * phi + p2)[:, None]
    f = f / 2.0
    sinphi = _sin(phi)[:, None]
    # Antarctic cap + more land in the north
    f = f + 2.5 * (np.abs(lat)[:, None] > 72.0) * (lat[:, None] < 0) + 0.35 * sinphi - 0.25
    land = f > 0.35
    height = 1.0 + 1800.0 * (f - 0.35) ** 2 + 900.0 * (f - 0.35)
    height = np.minimum(height, 4500.0)
    z = np.where(land, height, -0.1)
    return z.astype(np.float32)


def make_glacier(z_topo):
    lat, _ = _grid()
    g = (z_topo > 0) & ((lat[:, None] < -66.0) | ((lat[:, None] > 62.0) & (z_topo > 900.0)))
    return g.astype(np.float32)


def make_solar():
    """Daily-mean insolation [730][48], identical for the two half-day steps of a day."""
    lat, _ = _grid()
    phi = lat * (_PI / 180.0)
    out = np.zeros((NSTEP_YR, YDIM))
    S0 = 1365.0
    for d in range(365):
        g = 2 * _PI * (d + 0.5) / 365.0
        sg, cg = (float(v) for v in _sin_cos(np.float64(g)))
        s2g, c2g = (float(v) for v in _sin_cos(np.float64(2 * g)))
        decl = 0.006918 - 0.399912 * cg + 0.070257 * sg - 0.006758 * c2g + 0.000907 * s2g
        dist = 1.00011 + 0.034221 * cg + 0.00128 * sg + 0.000719 * c2g + 0.000077 * s2g
        sd, cd = (float(v) for v in _sin_cos(np.float64(decl)))
        sp, cp = _sin_cos(phi)
        x = -(sp * sd) / (cp * cd)          # cos(h0)
        x = np.clip(x, -1.0, 1.0)
        # h0 = acos(x) via atan2-free series is overkill: use sqrt identity + Newton on cos
        h0 = _acos(x)
        q = S0 / _PI * dist * (h0 * sp * sd + cp * cd * _sin(h0))
        q = np.maximum(q, 0.0)
        out[2 * d] = q
        out[2 * d + 1] = q
    return out.astype(np.float32)


def _acos(x):
    """acos on [-1,1] by 6 Newton steps on cos(h)=x from a sqrt-based start (deterministic)."""
    x = np.asarray(x, dtype=np.float64)
    h = np.sqrt(np.maximum(0.0, 2.0 * (1.0 - x)))     # good near x=1
    h = np.where(x < 0, _PI - np.sqrt(np.maximum(0.0, 2.0 * (1.0 + x))), h)
    for _ in range(8):
        s, c = _sin_cos(h)
        s = np.where(np.abs(s) < 1e-12, 1e-12, s)
        h = np.clip(h + (c - x) / s, 0.0, _PI)
    h = np.where(x >= 1.0, 0.0, np.where(x <= -1.0, _PI, h))
    return h


def make_forcing(seed: int = 20110101, topo: str = "synthetic", reference_input: str | None = None,
                 reference_fields=None) -> Forcing:
    """Generate the S0 set.

    topo="synthetic": all ten fields synthetic (what the GPU box and the bench use).
    topo="aquaplanet": z_topo == -0.1 everywhere, no glacier.
    topo="reference": take topography / glacier.masks / solar.radiation from the directory
    ``reference_input`` (the reference mount's input/) or from ``reference_fields`` =
    (z_topo [48][96], glacier [48][96], sw_solar [730][48]) — the arrays a test fixture carries; the other
    seven fields are generated around that orography."""
    rng = np.random.default_rng(seed)
    lat, lon = _grid()
    phi = lat * (_PI / 180.0)
    sinphi, cosphi = _sin_cos(phi)

    if topo == "reference":
        if reference_fields is not None:
            z_topo, glacier, sw_solar = (np.asarray(a, dtype=np.float32) for a in reference_fields)
        else:
            z_topo = np.fromfile(os.path.join(reference_input, "topography"), dtype="<f4").reshape(YDIM, XDIM)
            glacier = np.fromfile(os.path.join(reference_input, "glacier.masks"), dtype="<f4").reshape(YDIM, XDIM)
            sw_solar = np.fromfile(os.path.join(reference_input, "solar.radiation"), dtype="<f4").reshape(NSTEP_YR, YDIM)
        _ = make_topography(rng)  # keep the random stream aligned with the synthetic variant
    elif topo == "aquaplanet":
        _ = make_topography(rng)
        z_topo = np.full((YDIM, XDIM), -0.1, dtype=np.float32)
        glacier = np.zeros((YDIM, XDIM), dtype=np.float32)
        sw_solar = make_solar()
    else:
        z_topo = make_topography(rng)
        glacier = make_glacier(z_topo)
        sw_solar = make_solar()

    z = z_topo.astype(np.float64)
    land = z > 0
    zpos = np.maximum(z, 0.0)

    nT = _smooth_noise(rng)
    nU = _smooth_noise(rng)
    nV = _smooth_noise(rng)
    nC = _smooth_noise(rng)

    tclim = np.empty((NSTEP_YR, YDIM, XDIM), dtype=np.float32)
    qclim = np.empty_like(tclim)
    swet = np.empty_like(tclim)
    ucl = np.empty_like(tclim)
    vcl = np.empty_like(tclim)
    mld = np.empty_like(tclim)
    cld = np.empty_like(tclim)

    s2 = (sinphi * sinphi)[:, None]
    abss = np.abs(sinphi)[:, None]
    sgn = np.where(lat >= 0, 1.0, -1.0)[:, None]
    amp = np.where(land, 18.0, 5.0) * abss
    cos3, sin6 = _cos(3 * phi)[:, None], _sin(6 * phi)[:, None]
    taper = np.where(np.abs(lat)[:, None] > 75.0, 1.0 / 3.0, 1.0)
    vtaper = np.ones((YDIM, 1))
    vtaper[0, 0] = 0.0
    vtaper[-1, 0] = 0.0
    vtaper[1, 0] = 0.5
    vtaper[-2, 0] = 0.5
    sin2 = np.abs(_sin(2 * phi))[:, None]

    for d in range(365):
        season = float(_cos(np.float64(2 * _PI * (d - 15) / 365.0)))   # +1 mid-January
        t = 300.0 - 45.0 * s2 - 6.5e-3 * zpos - amp * season * sgn + 0.5 * nT(d)
        t = np.where(land, t, np.maximum(t, 271.35))
        # saturation humidity with the model's own formula (greb.f90:457-458)
        qs = 3.75e-3 * _exp(17.08085 * (t - 273.15) / (t - 273.15 + 234.175)) * _exp(-z / 8400.0)
        q = 0.75 * qs
        sw = np.where(land, 0.15 + 0.35 * (cosphi * cosphi)[:, None], 1.0) * np.ones((YDIM, XDIM))
        u = (-6.0 * cos3 * (1.0 + 0.2 * season * sgn) + 1.0 * nU(d)) * taper
        u = np.clip(u, -12.0, 12.0)
        v = (2.0 * sin6 * cosphi[:, None] * (1.0 + 0.2 * season) + 1.0 * nV(d)) * vtaper
        v = np.clip(v, -3.0, 3.0)
        winter = 0.5 * (1.0 + season * sgn)                              # 1 in local winter
        m = 40.0 + 60.0 * winter * abss + 100.0 * winter * (np.abs(lat)[:, None] > 45.0)
        m = np.where(land, 50.0, np.clip(m, 20.0, 250.0)) * np.ones((YDIM, XDIM))
        c = np.clip(0.45 + 0.25 * sin2 + 0.08 * nC(d), 0.1, 0.9)
        for arr, val in ((tclim, t), (qclim, q), (swet, sw), (ucl, u), (vcl, v), (mld, m), (cld, c)):
            arr[2 * d] = val.astype(np.float32)
            arr[2 * d + 1] = arr[2 * d]

    return Forcing(z_topo=z_topo.astype(np.float32), glacier=glacier.astype(np.float32),
                   sw_solar=sw_solar.astype(np.float32), tclim=tclim, qclim=qclim, swetclim=swet,
                   uclim=ucl, vclim=vcl, mldclim=mld, cldclim=cld)


_CACHE: dict = {}


def cached_forcing(seed: int = 20110101, topo: str = "synthetic", cache_dir: str | None = None) -> Forcing:
    """make_forcing with an in-process and optional on-disk cache (generation takes a few seconds)."""
    key = (seed, topo)
    if key in _CACHE:
        return _CACHE[key]
    if cache_dir is not None:
        d = os.path.join(cache_dir, f"S0_{topo}_{seed}")
        if os.path.exists(os.path.join(d, "cloud.cover")):
            f = Forcing.read(d)
        else:
            f = make_forcing(seed, topo)
            f.write(d)
    else:
        f = make_forcing(seed, topo)
    _CACHE[key] = f
    return f
