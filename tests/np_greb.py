"""Independent NumPy-float32 transcription of the reference kernels (array semantics).

TEST INFRASTRUCTURE.  Written directly from /root/reference/src/greb.f90 (line numbers
"f:NNN"), *not* from oracle/greb_oracle.c, using whole-array expressions (np.roll for the
periodic longitude wrap, np.where for Fortran `where`) so that a transcription slip in either
restatement shows up as a mismatch between the two.

All arrays are float32 [48][96]; python float literals are NEP-50 "weak" scalars, so every
operation is a correctly rounded IEEE fp32 operation in the order written.
"""
import math

import numpy as np

f32 = np.float32
XD, YD, NT = 96, 48, 730
DT, DT_CRCL = f32(43200), f32(1800)
DLON = f32(360.0) / f32(XD)
DLAT = f32(180.0) / f32(YD)


def _cosf(x):
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float
    libm.cosf.argtypes = [ctypes.c_float]
    return f32(libm.cosf(ctypes.c_float(float(x))))


def nint(x):
    return int(math.floor(float(x) + 0.5)) if x >= 0 else -int(math.floor(-float(x) + 0.5))


class Geo:
    def __init__(self, pi=3.1416, kappa=8e5):
        pi, kappa = f32(pi), f32(kappa)
        self.deg = f32(2.0) * pi * f32(6.371e6) / f32(360.0)          # f:578
        self.dyy = DLAT * self.deg                                   # f:579
        k = np.arange(1, YD + 1).astype(np.float32)
        self.lat = DLAT * k - DLAT / f32(2.0) - f32(90.0)             # f:580
        ang = f32(2.0) * pi / f32(360.0) * self.lat
        self.dxlat = (DLON * self.deg) * np.array([_cosf(a) for a in ang], dtype=np.float32)
        self.ccy_d = kappa * DT_CRCL / (self.dyy * self.dyy)                 # f:581
        self.ccx_d = kappa * DT_CRCL / (self.dxlat * self.dxlat)               # f:582
        self.ccy_a = DT_CRCL / self.dyy / f32(2.0)                   # f:752
        self.ccx_a = DT_CRCL / self.dxlat / f32(2.0)                 # f:753
        self.polar = ~(self.dxlat > f32(2.5e5))                      # f:592
        self.t2_d, self.ccx2_d, self.t2_a, self.ccx2_a = [], [], [], []
        for d in self.dxlat:
            dd = f32(max(1, nint(DT_CRCL / (f32(1.0) * (d * d) / kappa))))   # f:652
            dtdff2 = int(DT_CRCL / dd)
            self.t2_d.append(max(1, nint(f32(1800) / f32(dtdff2))))         # f:653
            self.ccx2_d.append(kappa * f32(dtdff2) / (d * d))                # f:654
            dd = f32(max(1, nint(DT_CRCL / (d / f32(10.0) / f32(1.0)))))    # f:838
            dtdff2 = int(DT_CRCL / dd)
            self.t2_a.append(max(1, nint(f32(1800) / f32(dtdff2))))         # f:839
            self.ccx2_a.append(f32(dtdff2) / d / f32(2))                    # f:840


def _sh(a, n):
    """a(j+n) with periodic wrap along the last axis."""
    return np.roll(a, -n, axis=-1)


def _diff_x(T, w, cc):
    """f:620-625 for whole rows; T, w are [..., 96]."""
    return cc * (10 * (_sh(w, -1) * (_sh(T, -1) - T) + _sh(w, 1) * (_sh(T, 1) - T))
                 + 4 * (_sh(w, -2) * (_sh(T, -2) - _sh(T, -1)) + _sh(w, -1) * (T - _sh(T, -1)))
                 + 4 * (_sh(w, 1) * (T - _sh(T, 1)) + _sh(w, 2) * (_sh(T, 2) - _sh(T, 1)))
                 + 1 * (_sh(w, -3) * (_sh(T, -3) - _sh(T, -2)) + _sh(w, -2) * (_sh(T, -1) - _sh(T, -2)))
                 + 1 * (_sh(w, 2) * (_sh(T, 1) - _sh(T, 2)) + _sh(w, 3) * (_sh(T, 3) - _sh(T, 2)))) / f32(20.0)


def diffusion(T1, wz, geo):
    T1 = T1.astype(np.float32)
    dTy = np.empty_like(T1)
    dTx = np.empty_like(T1)
    ccy = geo.ccy_d
    dTy[1:-1] = ccy * (wz[:-2] * (T1[:-2] - T1[1:-1]) + wz[2:] * (T1[2:] - T1[1:-1]))    # f:587
    dTy[0] = ccy * wz[1] * (-T1[0] + T1[1])                                              # f:589
    dTy[-1] = ccy * wz[-2] * (T1[-2] - T1[-1])                                           # f:590
    for k in range(YD):
        if not geo.polar[k]:
            dTx[k] = _diff_x(T1[k], wz[k], geo.ccx_d[k])
        else:
            T1h = T1[k].copy()
            for _ in range(geo.t2_d[k]):
                d = _diff_x(T1h, wz[k], geo.ccx2_d[k])
                d = np.where(d <= -T1h, f32(-0.9) * T1h, d)                               # f:715
                T1h = T1h + d                                                            # f:716
            dTx[k] = T1h - T1[k]                                                         # f:718
    return wz * (dTx + dTy)                                                              # f:721


def advection(T1, wz, um, up, vm, vp, geo):
    T1 = T1.astype(np.float32)
    dTy = np.empty_like(T1)
    dTx = np.empty_like(T1)
    ccy = geo.ccy_a
    T = T1
    # f:756-795
    dTy[0] = ccy * (vp[0] * (wz[1] * (T[0] - T[1]) + wz[2] * (T[0] - T[2]))) / f32(3.0)
    dTy[1] = ccy * (-vm[1] * (wz[0] * (T[1] - T[0]))
                    + vp[1] * (wz[2] * (T[1] - T[2]) + wz[3] * (T[1] - T[3])) / f32(3.0))
    c = slice(2, YD - 2)
    dTy[c] = ccy * (-vm[c] * (wz[1:YD - 3] * (T[c] - T[1:YD - 3]) + wz[0:YD - 4] * (T[c] - T[0:YD - 4]))
                    + vp[c] * (wz[3:YD - 1] * (T[c] - T[3:YD - 1]) + wz[4:YD] * (T[c] - T[4:YD]))) / f32(3.0)
    k = YD - 2
    dTy[k] = ccy * (-vm[k] * (wz[k - 1] * (T[k] - T[k - 1]) + wz[k - 2] * (T[k] - T[k - 2])) / f32(3.0)
                    + vp[k] * (wz[k + 1] * (T[k] - T[k + 1])))
    k = YD - 1
    dTy[k] = ccy * (-vm[k] * (wz[k - 1] * (T[k] - T[k - 1]) + wz[k - 2] * (T[k] - T[k - 2]))) / f32(3.0)
    for k in range(YD):
        w = wz[k]
        if not geo.polar[k]:                                                             # f:816-820
            t = T[k]
            dTx[k] = geo.ccx_a[k] * (-um[k] * (_sh(w, -1) * (t - _sh(t, -1)) + _sh(w, -2) * (t - _sh(t, -2)))
                                     + up[k] * (_sh(w, 1) * (t - _sh(t, 1)) + _sh(w, 2) * (t - _sh(t, 2)))) / f32(3.0)
        else:
            h = T[k].copy()
            for _ in range(geo.t2_a[k]):
                d = geo.ccx2_a[k] * (-um[k] * (10 * _sh(w, -1) * (h - _sh(h, -1))
                                               + 4 * _sh(w, -2) * (_sh(h, -1) - _sh(h, -2))
                                               + 1 * _sh(w, -3) * (_sh(h, -2) - _sh(h, -3)))
                                     + up[k] * (10 * _sh(w, 1) * (h - _sh(h, 1))
                                                + 4 * _sh(w, 2) * (_sh(h, 1) - _sh(h, 2))
                                                + 1 * _sh(w, 3) * (_sh(h, 2) - _sh(h, 3)))) / f32(20.0)
                # f:880-888, 1-based j=xdim-2=94: jp1=95, jp2=95 (sic), jp3=1  -> 0-based 93; 94, 94, 0
                j = XD - 3
                d[j] = geo.ccx2_a[k] * (-um[k, j] * (10 * w[j - 1] * (h[j] - h[j - 1])
                                                     + 4 * w[j - 2] * (h[j - 1] - h[j - 2])
                                                     + 1 * w[j - 3] * (h[j - 2] - h[j - 3]))
                                        + up[k, j] * (10 * w[XD - 2] * (h[j] - h[XD - 2])
                                                      + 4 * w[XD - 2] * (h[XD - 2] - h[XD - 2])
                                                      + 1 * w[0] * (h[XD - 2] - h[0]))) / f32(20.0)
                d = np.where(d <= -h, f32(-0.9) * h, d)                                   # f:907
                h = h + d
            dTx[k] = h - T[k]                                                            # f:910
    return dTx + dTy                                                                     # f:913


def circulation(X_in, wz, u, v, geo):
    um = np.where(u >= 0, u, f32(0)).astype(np.float32)   # f:203-216
    up = np.where(u >= 0, f32(0), u).astype(np.float32)
    vm = np.where(v >= 0, v, f32(0)).astype(np.float32)
    vp = np.where(v >= 0, f32(0), v).astype(np.float32)
    X = X_in.astype(np.float32).copy()
    for _ in range(max(1, nint(DT / DT_CRCL))):           # f:543-550
        X = X + diffusion(X, wz, geo) + advection(X, wz, um, up, vm, vp, geo)
    return X - X_in                                       # f:551


# ---- column physics (transcendentals via numpy: compare with a few-ulp tolerance) -----------

def SWradiation(Ts, cld, sw_solar_row, z_topo, glacier, p):
    a_atmos = cld * f32(p.a_cloud)
    a_ice = f32(p.a_no_ice) + f32(p.da_ice)
    a = np.zeros_like(Ts)
    land = z_topo >= 0
    oce = z_topo < 0
    a = np.where(land & (Ts <= f32(p.Tl_ice1)), a_ice, a)
    a = np.where(land & (Ts >= f32(p.Tl_ice2)), f32(p.a_no_ice), a)
    a = np.where(land & (Ts > f32(p.Tl_ice1)) & (Ts < f32(p.Tl_ice2)),
                 f32(p.a_no_ice) + f32(p.da_ice) * (1 - (Ts - f32(p.Tl_ice1)) / (f32(p.Tl_ice2) - f32(p.Tl_ice1))), a)
    a = np.where(oce & (Ts <= f32(p.To_ice1)), a_ice, a)
    a = np.where(oce & (Ts >= f32(p.To_ice2)), f32(p.a_no_ice), a)
    a = np.where(oce & (Ts > f32(p.To_ice1)) & (Ts < f32(p.To_ice2)),
                 f32(p.a_no_ice) + f32(p.da_ice) * (1 - (Ts - f32(p.To_ice1)) / (f32(p.To_ice2) - f32(p.To_ice1))), a)
    a = np.where(glacier > 0.5, a_ice, a)
    albedo = a + a_atmos - a * a_atmos
    sw = sw_solar_row[:, None] * (1 - albedo)
    return sw.astype(np.float32), albedo.astype(np.float32)


def LWradiation(Ts, Ta, q, co2, cld, dTrad, z_topo, p):
    pe = [f32(x) for x in p.p_emi]
    ez = np.exp(-z_topo / f32(p.z_air))
    e_co2 = ez * f32(co2)
    e_vapor = ez * f32(p.r_qviwv) * q
    em = (pe[3] * np.log(pe[0] * e_co2 + pe[1] * e_vapor + pe[2]) + pe[6]
          + pe[4] * np.log(pe[0] * e_co2 + pe[2]) + pe[5] * np.log(pe[1] * e_vapor + pe[2]))
    em = (pe[7] - cld) / pe[8] * (em - pe[9]) + pe[9]
    LWsurf = -f32(p.sig) * ((Ts * Ts) * (Ts * Ts))
    Tr = Ta + dTrad
    LWdown = -em * f32(p.sig) * ((Tr * Tr) * (Tr * Tr))
    return LWsurf, LWdown.copy(), LWdown, em


def hydro(Ts, q, u, v, swet, z_topo, p):
    absw = np.sqrt(u * u + v * v)
    absw = np.where(z_topo > 0, np.sqrt(absw * absw + f32(4.0)), absw)
    absw = np.where(z_topo < 0, np.sqrt(absw * absw + f32(9.0)), absw)
    qs = f32(3.75e-3) * np.exp(f32(17.08085) * (Ts - f32(273.15)) / (Ts - f32(273.15) + f32(234.175)))
    qs = qs * np.exp(-z_topo / f32(p.z_air))
    Qlat = (q - qs) * absw * f32(p.cq_latent) * f32(p.rho_air) * f32(p.ce) * swet
    dq_eva = -Qlat / f32(p.cq_latent) / f32(p.r_qviwv)
    dq_rain = f32(p.cq_rain) * q
    Qlat_air = -dq_rain * f32(p.cq_latent) * f32(p.r_qviwv)
    return Qlat, Qlat_air, dq_eva, dq_rain


def deep_ocean(Ts, To, mld, mld_prev, z_ocean, z_topo, p, cap_ocean):
    dmld = mld - mld_prev
    oce = (z_topo < 0) & (Ts >= f32(p.To_ice2))
    dTo = np.where(oce & (dmld < 0), -dmld / (z_ocean - mld) * (Ts - To), f32(0))
    dT_ocean = np.where(oce & (dmld > 0), dmld / mld * (To - Ts), f32(0))
    dTo = f32(0.5) * dTo
    dT_ocean = f32(0.5) * dT_ocean
    Tx = np.maximum(f32(p.To_ice2), Ts)
    dTo = dTo + DT * f32(p.co_turb) * (Tx - To) / (cap_ocean * (z_ocean - mld))
    dT_ocean = dT_ocean + DT * f32(p.co_turb) * (To - Tx) / (cap_ocean * mld)
    return dT_ocean.astype(np.float32), dTo.astype(np.float32)


def seaice(cap_surf, Ts, mld, z_topo, glacier, p, cap_land, cap_ocean):
    oce = z_topo < 0
    c = cap_surf.copy()
    c = np.where(oce & (Ts <= f32(p.To_ice1)), cap_land, c)
    c = np.where(oce & (Ts >= f32(p.To_ice2)), cap_ocean * mld, c)
    c = np.where(oce & (Ts > f32(p.To_ice1)) & (Ts < f32(p.To_ice2)),
                 cap_land + (cap_ocean * mld - cap_land) / (f32(p.To_ice2) - f32(p.To_ice1)) * (Ts - f32(p.To_ice1)), c)
    c = np.where(glacier > 0.5, cap_land, c)
    return c.astype(np.float32)
