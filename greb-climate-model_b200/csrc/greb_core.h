// greb_core.h — warp-level implementation of the GREB 12-hourly step for one ensemble member.
//
// One CTA (12 warps) integrates one member.  Warp w owns latitude rows [row0, row0+nrow) of the
// 96x48 grid; a lane owns 3 consecutive longitudes of each of those rows, so a row is exactly one
// warp wide and the periodic longitude wrap is a lane rotation (SHFL), while the poles are just
// the first / last rows with their one-sided formulas.  During the 24 circulation sub-steps the
// advected field, its weights wz and the winds stay in registers; only the two rows next to a
// warp's band are exchanged through a double-buffered shared-memory copy of the field, with one
// CTA barrier per sub-step.  Everything else of the step is column-local.
//
// Arithmetic contract ("exact mode"): the reference is gfortran -O3 without -ffast-math, i.e.
// IEEE fp32, no FMA contraction, expression order as written.  This file is compiled with
// -fmad=false; v_fma is used only where it is provably identical to the written form
// (multiplication by 4 is exact), and the divisions by the literals 3. and 20. use a
// correctly-rounded 3-instruction sequence (div_c).  Identities used to share work between
// cells — x-(y) == x+(-y), (-a)*b == -(a*b), RN(-x) == -RN(x), a+b == b+a — are exact in IEEE
// arithmetic, so results are bit-identical to the as-written evaluation (signs of zeros aside).
//
// Reference: /root/reference/src/greb.f90 ("f:NNN" below).
#pragma once

#include "greb_simt.h"
#include "greb_types.h"

#define GREB_DT 43200.0f  // f:38  (integer dt in real expressions)

// ---- correctly rounded x/3 and x/20 ---------------------------------------------------------
// q0 = x*RN(1/d); r = fma(-d,q0,x); q = fma(r,RN(1/d),q0) equals RN(x/d) for every float x whose
// quotient is a normal number (exhaustively verified for d = 3 and d = 20 in tests/test_divc.py).
#if GREB_DEVICE
GDEV vf div3(vf x) {
  const float r = 0.3333333432674407958984375f;
  float q = __fmul_rn(x, r);
  float e = __fmaf_rn(-3.0f, q, x);
  return __fmaf_rn(e, r, q);
}
GDEV vf div20(vf x) {
  const float r = 0.0500000007450580596923828125f;
  float q = __fmul_rn(x, r);
  float e = __fmaf_rn(-20.0f, q, x);
  return __fmaf_rn(e, r, q);
}
#else
GDEV vf div3(vf x) { return x / 3.0f; }
GDEV vf div20(vf x) { return x / 20.0f; }
#endif

// clamp of the polar sub-sub-steps: where(d <= -T) d = -0.9*T   (f:715, f:907)
GDEV vf polar_clamp(vf d, vf T) { return v_sel(d <= -T, -0.9f * T, d); }

// ---------------------------------------------------------------------------------------------
// x-direction products of one row segment (3 own cells + neighbours).
//   d(m) = T(m+1)-T(m),  P(m) = wz(m)*d(m),  Q(m) = wz(m+1)*d(m)
// With c0 the lane's first cell: Pm3..P1 = P(c0-3..c0+1), Q0..Qp2 = Q(c0..c0+4).
// ---------------------------------------------------------------------------------------------
struct XRow {
  vf xm2, xm1, xp1, xp2;            // T(c0-2), T(c0-1), T(c0+3), T(c0+4)
  vf dm1, d0, d1, d2;               // d(c0-1..c0+2)
  vf Pm3, Pm2, Pm1, P0, P1;
  vf Q0, Q1, Q2, Qp1, Qp2;
};

GDEV void xrow_products(XRow& x, const vf (&T)[3], const vf (&W)[3], const vf (&WX)[4], vi lane_l, vi lane_r) {
  x.xm2 = v_shfl(T[1], lane_l);
  x.xm1 = v_shfl(T[2], lane_l);
  x.xp1 = v_shfl(T[0], lane_r);
  x.xp2 = v_shfl(T[1], lane_r);
  const vf dm2 = x.xm1 - x.xm2;
  x.dm1 = T[0] - x.xm1;
  x.d0 = T[1] - T[0];
  x.d1 = T[2] - T[1];
  x.d2 = x.xp1 - T[2];
  const vf dp1 = x.xp2 - x.xp1;
  x.Pm2 = WX[0] * dm2;
  x.Pm1 = WX[1] * x.dm1;
  x.P0 = W[0] * x.d0;
  x.P1 = W[1] * x.d1;
  x.Q0 = W[1] * x.d0;
  x.Q1 = W[2] * x.d1;
  x.Q2 = WX[2] * x.d2;
  x.Qp1 = WX[3] * dp1;
  x.Pm3 = v_shfl(x.P0, lane_l);   // left lane's P(c0') = P(c0-3)
  x.Qp2 = v_shfl(x.Q1, lane_r);   // right lane's Q(c0'+1) = Q(c0+4)
}

// the bracket of f:620-625 for the 3 own cells:
//   10*(Q(j)-P(j-1)) + 4*(P(j-1)-P(j-2)) + 4*(Q(j+1)-Q(j)) + (P(j-2)-P(j-3)) + (Q(j+2)-Q(j+1))
GDEV void xdiff_bracket(vf (&S)[3], const XRow& x) {
  const vf Am2 = x.Pm2 - x.Pm3, Am1 = x.Pm1 - x.Pm2, A0 = x.P0 - x.Pm1, A1 = x.P1 - x.P0;  // A(m)=P(m)-P(m-1)
  const vf B0 = x.Q1 - x.Q0, B1 = x.Q2 - x.Q1, B2 = x.Qp1 - x.Q2, B3 = x.Qp2 - x.Qp1;     // B(m)=Q(m+1)-Q(m)
  const vf G0 = x.Q0 - x.Pm1, G1 = x.Q1 - x.P0, G2 = x.Q2 - x.P1;
  S[0] = v_fma(4.0f, B0, v_fma(4.0f, Am1, 10.0f * G0)) + Am2 + B1;
  S[1] = v_fma(4.0f, B1, v_fma(4.0f, A0, 10.0f * G1)) + Am1 + B2;
  S[2] = v_fma(4.0f, B2, v_fma(4.0f, A1, 10.0f * G2)) + A0 + B3;
}

// one extra polar diffusion sub-sub-step on the row copy h (f:656-717), generic slow path
GDEV void xdiff_polar_iter(vf (&h)[3], const vf (&W)[3], const vf (&WX)[4], vi lane_l, vi lane_r, float ccx2) {
  XRow x;
  xrow_products(x, h, W, WX, lane_l, lane_r);
  vf S[3];
  xdiff_bracket(S, x);
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    vf d = div20(ccx2 * S[c]);
    d = polar_clamp(d, h[c]);
    h[c] = h[c] + d;
  }
}

// polar x-advection bracket (f:872-878 and the wrap cases incl. the index bug of f:881) for the
// 3 own cells of a row copy h; returns X = -um*(...) + up*(...) per cell.
GDEV void xadv_polar_X(vf (&X)[3], const XRow& x, const vf (&T)[3], const vf (&W)[3], const vf (&WX)[4],
                       const vf (&U)[3], vb is_bug_lane) {
  // per own cell: P(j-1), P(j-2), P(j-3) / Q(j), Q(j+1), Q(j+2), d(j-1) / d(j), wz(j-1) / wz(j+1)
  const vf dlo[3] = {x.dm1, x.d0, x.d1}, dhi[3] = {x.d0, x.d1, x.d2};
  const vf wlo[3] = {WX[1], W[0], W[1]}, whi[3] = {W[1], W[2], WX[2]};
  const vf P2[3] = {x.Pm2, x.Pm1, x.P0}, P3[3] = {x.Pm3, x.Pm2, x.Pm1};
  const vf Qn1[3] = {x.Q1, x.Q2, x.Qp1}, Qn2[3] = {x.Q2, x.Qp1, x.Qp2};
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vb pu = U[c] >= 0.0f;
    const vf near10 = (10.0f * v_sel(pu, wlo[c], whi[c])) * v_sel(pu, dlo[c], dhi[c]);
    vf mid = v_sel(pu, P2[c], Qn1[c]);
    vf far = v_sel(pu, P3[c], Qn2[c]);
    if (c == 0) {
      // f:881: Fortran j = xdim-2 (0-based 93 = lane 31, cell 0) uses jp1 = jp2 = xdim-1, jp3 = 1:
      // the 4* term vanishes and the 1* term is wz(1)*(T(xdim-1)-T(1)).
      const vb bug = is_bug_lane && !pu;
      mid = v_sel(bug, v_bcast(0.0f), mid);
      far = v_sel(bug, WX[2] * (x.xp1 - T[1]), far);
    }
    const vf S = v_fma(4.0f, mid, near10) + far;
    X[c] = (-U[c]) * S;
  }
}

GDEV void xadv_polar_iter(vf (&h)[3], const vf (&W)[3], const vf (&WX)[4], const vf (&U)[3], vi lane_l,
                          vi lane_r, vb is_bug_lane, float ccx2) {
  XRow x;
  xrow_products(x, h, W, WX, lane_l, lane_r);
  vf X[3];
  xadv_polar_X(X, x, h, W, WX, U, is_bug_lane);
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    vf d = div20(ccx2 * X[c]);
    d = polar_clamp(d, h[c]);
    h[c] = h[c] + d;
  }
}

// ---------------------------------------------------------------------------------------------
// One circulation sub-step for the 3 cells a lane owns in latitude row k:
//   out = (T + dX_diffuse) + dX_advec                                    (f:547-549)
// Tm2..Tp2 / Wm2..Wp2 are the rows k-2..k+2 of the field and of wz (values of rows outside
// 0..47 are never used: the boundary rows have their own formulas).
// ---------------------------------------------------------------------------------------------
GDEV void row_update(vf (&out)[3], const GrebMemberConst& mc, int k, const vf (&Tm2)[3], const vf (&Tm1)[3],
                     const vf (&T)[3], const vf (&Tp1)[3], const vf (&Tp2)[3], const vf (&Wm2)[3],
                     const vf (&Wm1)[3], const vf (&W)[3], const vf (&Wp1)[3], const vf (&Wp2)[3],
                     const vf (&WX)[4], const vf (&U)[3], const vf (&V)[3], vi lane_l, vi lane_r,
                     vb is_bug_lane) {
  const bool polar = mc.polar[k] != 0;
  XRow x;
  xrow_products(x, T, W, WX, lane_l, lane_r);

  // ---------------- diffusion, longitudinal (f:592-719) ----------------
  vf dTx[3];
  {
    vf S[3];
    xdiff_bracket(S, x);
    if (!polar) {
      const float cc = mc.ccx_diff[k];
      GUNROLL
      for (int c = 0; c < 3; ++c) dTx[c] = div20(cc * S[c]);
    } else {
      const float cc2 = mc.ccx2_diff[k];
      vf h[3];
      GUNROLL
      for (int c = 0; c < 3; ++c) {
        vf d = div20(cc2 * S[c]);
        d = polar_clamp(d, T[c]);   // f:715
        h[c] = T[c] + d;            // f:716
      }
      const int time2 = mc.time2_diff[k];
      GNOUNROLL
      for (int tt2 = 1; tt2 < time2; ++tt2) xdiff_polar_iter(h, W, WX, lane_l, lane_r, cc2);
      GUNROLL
      for (int c = 0; c < 3; ++c) dTx[c] = h[c] - T[c];  // f:718
    }
  }

  // ---------------- y-direction edge products ----------------
  vf Pym1[3], Qy0[3];  // wz(k-1)*(T(k)-T(k-1)),  wz(k+1)*(T(k+1)-T(k))
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    Pym1[c] = Wm1[c] * (T[c] - Tm1[c]);
    Qy0[c] = Wp1[c] * (Tp1[c] - T[c]);
  }

  // ---------------- diffusion, latitudinal (f:587-590) ----------------
  vf dTy[3];
  {
    const float ccy = mc.ccy_diff;
    if (k >= 1 && k <= GY - 2) {
      GUNROLL
      for (int c = 0; c < 3; ++c) dTy[c] = ccy * (Qy0[c] - Pym1[c]);
    } else if (k == 0) {
      GUNROLL
      for (int c = 0; c < 3; ++c) dTy[c] = (ccy * Wp1[c]) * (Tp1[c] - T[c]);
    } else {
      GUNROLL
      for (int c = 0; c < 3; ++c) dTy[c] = (ccy * Wm1[c]) * (Tm1[c] - T[c]);
    }
  }

  // ---------------- advection, latitudinal (f:756-795) ----------------
  vf aTy[3];
  {
    const float ccy = mc.ccy_adv;
    if (k >= 2 && k <= GY - 3) {  // f:774-778
      GUNROLL
      for (int c = 0; c < 3; ++c) {
        const vb pv = V[c] >= 0.0f;
        const vf near = v_sel(pv, Pym1[c], -Qy0[c]);
        const vf far = v_sel(pv, Wm2[c], Wp2[c]) * (T[c] - v_sel(pv, Tm2[c], Tp2[c]));
        const vf Xv = (-v_abs(V[c])) * (near + far);
        aTy[c] = div3(ccy * Xv);
      }
    } else {
      GUNROLL
      for (int c = 0; c < 3; ++c) {
        const vb pv = V[c] >= 0.0f;
        const vf vm = v_sel(pv, V[c], v_bcast(0.0f));   // f:210-216
        const vf vp = v_sel(pv, v_bcast(0.0f), V[c]);
        if (k == 0) {  // f:759-761
          const vf s2 = (-Qy0[c]) + Wp2[c] * (T[c] - Tp2[c]);
          aTy[c] = div3(ccy * (vp * s2));
        } else if (k == 1) {  // f:766-769
          const vf s2 = (-Qy0[c]) + Wp2[c] * (T[c] - Tp2[c]);
          aTy[c] = ccy * (-(vm * Pym1[c]) + div3(vp * s2));
        } else if (k == GY - 2) {  // f:784-787
          const vf s1 = Pym1[c] + Wm2[c] * (T[c] - Tm2[c]);
          aTy[c] = ccy * (-div3(vm * s1) + vp * (-Qy0[c]));
        } else {  // k == GY-1, f:792-794
          const vf s1 = Pym1[c] + Wm2[c] * (T[c] - Tm2[c]);
          aTy[c] = div3(ccy * (-(vm * s1)));
        }
      }
    }
  }

  // ---------------- advection, longitudinal (f:798-911) ----------------
  vf aTx[3];
  if (!polar) {  // f:816-820 and wrap cases
    const float cc = mc.ccx_adv[k];
    const vf nearP[3] = {x.Pm1, x.P0, x.P1}, nearQ[3] = {x.Q0, x.Q1, x.Q2};
    const vf tfm[3] = {x.xm2, x.xm1, T[0]}, tfp[3] = {T[2], x.xp1, x.xp2};
    const vf wfm[3] = {WX[0], WX[1], W[0]}, wfp[3] = {W[2], WX[2], WX[3]};
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      const vb pu = U[c] >= 0.0f;
      const vf near = v_sel(pu, nearP[c], -nearQ[c]);
      const vf far = v_sel(pu, wfm[c], wfp[c]) * (T[c] - v_sel(pu, tfm[c], tfp[c]));
      const vf Xu = (-v_abs(U[c])) * (near + far);
      aTx[c] = div3(cc * Xu);
    }
  } else {  // f:838-910
    const float cc2 = mc.ccx2_adv[k];
    vf X[3], h[3];
    xadv_polar_X(X, x, T, W, WX, U, is_bug_lane);
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      vf d = div20(cc2 * X[c]);
      d = polar_clamp(d, T[c]);  // f:907
      h[c] = T[c] + d;           // f:908
    }
    const int time2 = mc.time2_adv[k];
    GNOUNROLL
    for (int tt2 = 1; tt2 < time2; ++tt2) xadv_polar_iter(h, W, WX, U, lane_l, lane_r, is_bug_lane, cc2);
    GUNROLL
    for (int c = 0; c < 3; ++c) aTx[c] = h[c] - T[c];  // f:910
  }

  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vf dXd = W[c] * (dTx[c] + dTy[c]);  // f:721
    const vf dXa = aTx[c] + aTy[c];           // f:913
    out[c] = (T[c] + dXd) + dXa;              // f:549
  }
}

// ---------------------------------------------------------------------------------------------
// Per-warp register tile for the circulation.
// Row slots: Y[0..1] = rows k0-2, k0-1 (lower halo); Y[2..2+nr) = own rows; then 2 upper halo rows.
// ---------------------------------------------------------------------------------------------
struct CircTile {
  vf Y[GREB_MAXR + 4][3];
  vf WY[GREB_MAXR + 4][3];
  vf WX[GREB_MAXR][4];
  vf U[GREB_MAXR][3], V[GREB_MAXR][3];
};

struct WarpGeom {
  int k0, nr;          // owned rows [k0, k0+nr)
  vi col;              // 3*lane: first owned longitude
  vi lane_l, lane_r;   // rotating neighbours
  vb is_bug_lane;      // lane 31
};

GDEV WarpGeom warp_geom(const SimtCtx& ctx, const GrebMemberConst& mc) {
  WarpGeom g;
  g.k0 = warp_uniform(mc.row0[ctx.warp]);
  g.nr = warp_uniform(mc.nrow[ctx.warp]);
  const vi lane = ctx_lane(ctx);
  g.col = lane * 3;
  g.lane_l = (lane + 31) & 31;
  g.lane_r = (lane + 1) & 31;
  g.is_bug_lane = (lane == 31);
  return g;
}

// wz rows (own + halo) and wz x-halos from a [GNC] field in global memory
GDEV void circ_load_wz(CircTile& t, const WarpGeom& g, const float* wz) {
  GUNROLL
  for (int s = 0; s < GREB_MAXR + 4; ++s) {
    const int k = g.k0 - 2 + s;
    const bool ok = (k >= 0) && (k < GY) && (s < g.nr + 4);
    GUNROLL
    for (int c = 0; c < 3; ++c) t.WY[s][c] = ok ? v_ldg(wz, k * GX + g.col + c) : v_bcast(0.0f);
  }
  // x-halo columns c0-2, c0-1, c0+3, c0+4 (periodic)
  const vi cm2 = v_seli(g.col == 0, vi(GX - 2), g.col - 2);
  const vi cm1 = v_seli(g.col == 0, vi(GX - 1), g.col - 1);
  const vi cp1 = v_seli(g.col == GX - 3, vi(0), g.col + 3);
  const vi cp2 = v_seli(g.col == GX - 3, vi(1), g.col + 4);
  GUNROLL
  for (int r = 0; r < GREB_MAXR; ++r) {
    const bool ok = r < g.nr;
    const int k = ok ? g.k0 + r : g.k0;
    t.WX[r][0] = v_ldg(wz, k * GX + cm2);
    t.WX[r][1] = v_ldg(wz, k * GX + cm1);
    t.WX[r][2] = v_ldg(wz, k * GX + cp1);
    t.WX[r][3] = v_ldg(wz, k * GX + cp2);
  }
}

GDEV void circ_load_uv(CircTile& t, const WarpGeom& g, const float* u, const float* v) {
  GUNROLL
  for (int r = 0; r < GREB_MAXR; ++r) {
    const int k = (r < g.nr) ? g.k0 + r : g.k0;
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      t.U[r][c] = v_ldg(u, k * GX + g.col + c);
      t.V[r][c] = v_ldg(v, k * GX + g.col + c);
    }
  }
}

// own rows + both halos of the field from any [GNC] buffer (global at the start of a circulation)
GDEV void circ_load_field(CircTile& t, const WarpGeom& g, const float* X) {
  GUNROLL
  for (int s = 0; s < GREB_MAXR + 4; ++s) {
    const int k = g.k0 - 2 + s;
    const bool ok = (k >= 0) && (k < GY) && (s < g.nr + 4);
    GUNROLL
    for (int c = 0; c < 3; ++c) t.Y[s][c] = ok ? v_ld(X, k * GX + g.col + c) : v_bcast(0.0f);
  }
}

// halo rows only, from the shared-memory copy written by the neighbouring warps
GDEV void circ_load_halo(CircTile& t, const WarpGeom& g, const float* X) {
  GUNROLL
  for (int s = 0; s < GREB_MAXR + 4; ++s) {
    const int k = g.k0 - 2 + s;
    const bool halo = (s < 2) || (s >= g.nr + 2 && s < g.nr + 4);
    if (halo && k >= 0 && k < GY) {
      GUNROLL
      for (int c = 0; c < 3; ++c) t.Y[s][c] = v_ld(X, k * GX + g.col + c);
    }
  }
}

GDEV void circ_store_own(const CircTile& t, const WarpGeom& g, float* X) {
  GUNROLL
  for (int r = 0; r < GREB_MAXR; ++r) {
    if (r < g.nr) {
      GUNROLL
      for (int c = 0; c < 3; ++c) v_st(X, (g.k0 + r) * GX + g.col + c, t.Y[r + 2][c]);
    }
  }
}

GDEV void circ_substep(CircTile& t, const WarpGeom& g, const GrebMemberConst& mc) {
  vf Tn[GREB_MAXR][3];
  GUNROLL
  for (int r = 0; r < GREB_MAXR; ++r) {
    if (r < g.nr) {
      row_update(Tn[r], mc, g.k0 + r, t.Y[r], t.Y[r + 1], t.Y[r + 2], t.Y[r + 3], t.Y[r + 4], t.WY[r], t.WY[r + 1],
                 t.WY[r + 2], t.WY[r + 3], t.WY[r + 4], t.WX[r], t.U[r], t.V[r], g.lane_l, g.lane_r, g.is_bug_lane);
    }
  }
  GUNROLL
  for (int r = 0; r < GREB_MAXR; ++r) {
    if (r < g.nr) {
      GUNROLL
      for (int c = 0; c < 3; ++c) t.Y[r + 2][c] = Tn[r][c];
    }
  }
}

// circulation (f:528-553): 24 sub-steps on the tile; `hb` = two [GNC] shared-memory buffers.
// On entry the tile holds the field (own rows + halos); on exit the own rows hold X after 24
// sub-steps.  All warps of the CTA must call it together (CTA barriers inside).
GDEV void circulation_run(const SimtCtx& ctx, CircTile& t, const WarpGeom& g, const GrebMemberConst& mc, float* hb) {
  GNOUNROLL
  for (int tt = 0; tt < GSUB; ++tt) {
    circ_substep(t, g, mc);
    if (tt + 1 < GSUB) {
      float* buf = hb + (tt & 1) * GNC;
      circ_store_own(t, g, buf);
      cta_sync(ctx);
      circ_load_halo(t, g, buf);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Column physics + state update of one cell column (everything of time_loop / qflux_correction
// that is not the circulation).  Phase A runs before the circulations, B after circulation(Ta),
// C after circulation(q).
// ---------------------------------------------------------------------------------------------
GDEV vf pow4(vf x) {
  const vf x2 = x * x;
  return x2 * x2;  // gfortran expands x**4 as (x*x)*(x*x)
}

struct StepInfo {
  int ityr;     // 0-based step of year
  int month_end;  // 1 if a month ends at this step (f:975-976)
  float ndm;    // days of that month * 2
  int out_rec;  // month slot of this launch to write
  float co2;
  int spinup;
};

GDEV void column_phase_a(const SimtCtx& ctx, const GrebKernelArgs& a, const GrebMemberConst& mc, int member,
                         const StepInfo& si, int k, vi col, float* stash) {
  const float* forc = a.forc + (size_t)si.ityr * GF_COUNT * GNC;
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  const float* wz_air = a.wz + (size_t)mc.group * 2 * GNC;
  float* corr = a.corr + ((size_t)mc.group * GNT + si.ityr) * GC_COUNT * GNC;
  const float solar = a.sw_solar[si.ityr * GY + k];
  const float* pe = mc.p_emi;
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vi idx = k * GX + col + c;
    const vf Ts = v_ld(st + GS_TS * GNC, idx), Ta = v_ld(st + GS_TA * GNC, idx);
    const vf To = v_ld(st + GS_TO * GNC, idx), q = v_ld(st + GS_Q * GNC, idx);
    const vf cap = v_ld(st + GS_CAP * GNC, idx);
    const vf cld = v_ldg(forc + GF_CLD * GNC, idx), dTrad = v_ldg(forc + GF_DTRAD * GNC, idx);
    const vf swet = v_ldg(forc + GF_SWET * GNC, idx), absw = v_ldg(forc + GF_ABSWIND * GNC, idx);
    const vf mld = v_ldg(forc + GF_MLD * GNC, idx), dmld = v_ldg(forc + GF_DMLD * GNC, idx);
    const vi mask = v_ldgi(a.mask, idx);
    const vf zoc = v_ldg(a.z_ocean, idx), ez = v_ldg(wz_air, idx);
    const vb land_ge0 = v_bit(mask, 0), ocean = v_bit(mask, 1), glac = v_bit(mask, 2);

    // ---- SWradiation, f:380-401
    const vf a_atmos = cld * mc.a_cloud;
    const float a_ice = mc.a_no_ice + mc.da_ice;
    const vf T1 = v_sel(land_ge0, v_bcast(mc.Tl_ice1), v_bcast(mc.To_ice1));
    const vf T2 = v_sel(land_ge0, v_bcast(mc.Tl_ice2), v_bcast(mc.To_ice2));
    vf a_surf = mc.a_no_ice + mc.da_ice * (1.0f - (Ts - T1) / (T2 - T1));
    a_surf = v_sel(Ts <= T1, v_bcast(a_ice), a_surf);
    a_surf = v_sel(Ts >= T2, v_bcast(mc.a_no_ice), a_surf);
    a_surf = v_sel(glac, v_bcast(a_ice), a_surf);
    const vf albedo = a_surf + a_atmos - a_surf * a_atmos;
    const vf sw = solar * (1.0f - albedo);

    // ---- LWradiation, f:420-432
    const vf e_co2 = ez * si.co2;
    const vf e_vapor = ez * mc.r_qviwv * q;
    vf em = pe[3] * v_log(pe[0] * e_co2 + pe[1] * e_vapor + pe[2]) + pe[6] + pe[4] * v_log(pe[0] * e_co2 + pe[2]) +
            pe[5] * v_log(pe[1] * e_vapor + pe[2]);
    em = (pe[7] - cld) / pe[8] * (em - pe[9]) + pe[9];
    const vf LWsurf = -(mc.sig * pow4(Ts));
    const vf LWdown = -(em * mc.sig * pow4(Ta + dTrad));
    const vf LWup = LWdown;

    // ---- sensible heat, f:295
    const vf Qsens = mc.ct_sens * (Ta - Ts);

    // ---- hydro, f:457-467 (abswind incl. gustiness is precomputed on the host, f:452-454)
    vf qs = 3.75e-3f * v_exp(17.08085f * (Ts - 273.15f) / (Ts - 273.15f + 234.175f));
    qs = qs * ez;
    const vf Qlat = (q - qs) * absw * mc.cq_latent * mc.rho_air * mc.ce * swet;
    const vf dq_eva = -(Qlat / mc.cq_latent / mc.r_qviwv);
    const vf dq_rain = mc.cq_rain * q;
    const vf Qlat_air = -(dq_rain * mc.cq_latent * mc.r_qviwv);

    // ---- deep_ocean, f:505-523
    const vb warm = ocean && (Ts >= mc.To_ice2);
    vf dTo = v_sel(warm && (dmld < 0.0f), -(dmld / (zoc - mld) * (Ts - To)), v_bcast(0.0f));
    vf dToc = v_sel(warm && (dmld > 0.0f), dmld / mld * (To - Ts), v_bcast(0.0f));
    dTo = 0.5f * dTo;
    dToc = 0.5f * dToc;
    const vf Tx = v_max(v_bcast(mc.To_ice2), Ts);
    dTo = dTo + GREB_DT * mc.co_turb * (Tx - To) / (mc.cap_ocean * (zoc - mld));
    dToc = dToc + GREB_DT * mc.co_turb * (To - Tx) / (mc.cap_ocean * mld);

    vf Ts0, To0, tendA, tq;
    if (!si.spinup) {  // time_loop, f:258-264
      const vf TF = v_ld(corr + GC_TF * GNC, idx), ToF = v_ld(corr + GC_TOF * GNC, idx);
      Ts0 = Ts + dToc + GREB_DT * (sw + LWsurf - LWdown + Qlat + Qsens + TF) / cap;
      tendA = GREB_DT * (LWup + LWdown - em * LWsurf + Qlat_air - Qsens) / mc.cap_air;
      To0 = To + dTo + ToF;
      tq = GREB_DT * (dq_eva + dq_rain);
    } else {  // qflux_correction, f:333-351
      const vf Tclim = v_ldg(a.tclim + (size_t)si.ityr * GNC, idx), Toclim = v_ldg(a.toclim, idx);
      const vf dTs = GREB_DT * (sw + LWsurf - LWdown + Qlat + Qsens) / cap;
      const vf ts0 = Ts + dTs + dToc;
      tendA = GREB_DT * (LWup + LWdown - em * LWsurf + Qlat_air - Qsens) / mc.cap_air;
      const vf to0 = To + dTo;
      tq = GREB_DT * (dq_eva + dq_rain);
      const vf T_error = Tclim - ts0;
      const vf tf = T_error * cap / GREB_DT;
      v_st(corr + GC_TF * GNC, idx, tf);
      Ts0 = Ts + dTs + dToc + tf * GREB_DT / cap;
      const vf tof = Toclim - to0;
      v_st(corr + GC_TOF * GNC, idx, tof);
      To0 = To + dTo + tof;
    }

    // ---- seaice(Ts0), f:483-490
    vf capn = cap;
    {
      const vf capo = mc.cap_ocean * mld;
      vf ramp = mc.cap_land + (capo - mc.cap_land) / (mc.To_ice2 - mc.To_ice1) * (Ts0 - mc.To_ice1);
      ramp = v_sel(Ts0 <= mc.To_ice1, v_bcast(mc.cap_land), ramp);
      ramp = v_sel(Ts0 >= mc.To_ice2, capo, ramp);
      capn = v_sel(ocean, ramp, capn);
      capn = v_sel(glac, v_bcast(mc.cap_land), capn);
    }

    v_st(st + GS_TS * GNC, idx, Ts0);
    v_st(st + GS_TO * GNC, idx, To0);
    v_st(st + GS_CAP * GNC, idx, capn);
    v_st(stash, idx, tendA);
    v_st(stash + GNC, idx, tq);
    // diagnostics (f:945) runs in both loops; output (f:974) only in time_loop
    v_st(acc + GA_TSMN * GNC, idx, v_ld(acc + GA_TSMN * GNC, idx) + Ts0);
    if (!si.spinup) {
      v_st(acc + GA_TMM * GNC, idx, v_ld(acc + GA_TMM * GNC, idx) + Ts0);
      v_st(acc + GA_TOMM * GNC, idx, v_ld(acc + GA_TOMM * GNC, idx) + To0);
      v_st(acc + GA_APMM * GNC, idx, v_ld(acc + GA_APMM * GNC, idx) + albedo);
    }
  }
}

// after circulation(Ta): X holds the circulated air temperature of the 3 cells of row k
GDEV void column_phase_b(const GrebKernelArgs& a, int member, const StepInfo& si, int k, vi col, const vf (&X)[3],
                         const float* stash) {
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vi idx = k * GX + col + c;
    const vf Ta1 = v_ld(st + GS_TA * GNC, idx);
    const vf dTa_crcl = X[c] - Ta1;  // f:551
    const vf tendA = v_ld(stash, idx);
    const vf Ta0 = si.spinup ? (Ta1 + tendA + dTa_crcl)    // f:337
                             : (Ta1 + dTa_crcl + tendA);   // f:260
    v_st(st + GS_TA * GNC, idx, Ta0);
    if (!si.spinup) v_st(acc + GA_TAMM * GNC, idx, v_ld(acc + GA_TAMM * GNC, idx) + Ta0);
  }
}

// after circulation(q)
GDEV void column_phase_c(const GrebKernelArgs& a, const GrebMemberConst& mc, int member, const StepInfo& si, int k,
                         vi col, const vf (&X)[3], const float* stash) {
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  float* corr = a.corr + ((size_t)mc.group * GNT + si.ityr) * GC_COUNT * GNC;
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vi idx = k * GX + col + c;
    const vf q1 = v_ld(st + GS_Q * GNC, idx);
    const vf dq_crcl = X[c] - q1;  // f:551
    const vf tq = v_ld(stash + GNC, idx);
    vf q0;
    if (!si.spinup) {  // f:264-266
      vf dq = tq + dq_crcl + v_ld(corr + GC_QF * GNC, idx);
      dq = v_sel(dq <= -q1, -0.9f * q1, dq);
      q0 = q1 + dq;
      v_st(acc + GA_QMM * GNC, idx, v_ld(acc + GA_QMM * GNC, idx) + q0);
    } else {  // f:342, 353-355
      const vf qq0 = q1 + tq + dq_crcl;
      const vf qf = v_ldg(a.qclim + (size_t)si.ityr * GNC, idx) - qq0;
      v_st(corr + GC_QF * GNC, idx, qf);
      q0 = q1 + tq + dq_crcl + qf;
    }
    v_st(st + GS_Q * GNC, idx, q0);
  }
}

// month end (f:977-983): write the five means, zero the accumulators
GDEV void column_month_end(const GrebKernelArgs& a, int member, const StepInfo& si, int k, vi col) {
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  float* out = a.out ? a.out + ((size_t)member * a.out_months + si.out_rec) * 5 * GNC : nullptr;
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vi idx = k * GX + col + c;
    GUNROLL
    for (int f = 0; f < 5; ++f) {
      if (out) v_st(out + f * GNC, idx, v_ld(acc + f * GNC, idx) / si.ndm);
      v_st(acc + f * GNC, idx, v_bcast(0.0f));
    }
  }
}

GDEV StepInfo step_info(const GrebKernelArgs& a, const GrebMemberConst& mc, int member, int it) {
  // calendar of f:251-252 and the month-end test of f:975-976
  const int cum[12] = {31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 365};
  const int dim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  StepInfo si;
  si.ityr = (it - 1) % GNT;
  const int jday = ((it - 1) / 2) % 365 + 1;
  si.month_end = 0;
  si.ndm = 1.0f;
  si.spinup = a.spinup;
  if ((it & 1) == 0) {
    for (int m = 0; m < 12; ++m)
      if (jday == cum[m]) {
        si.month_end = 1;
        si.ndm = (float)(dim[m] * 2);
      }
  }
  // months completed before this step since the launch began (launches start on a year boundary
  // or anywhere: count month ends in (it0-1, it-1])
  int rec = 0;
  {
    const int first = a.it0;
    // month ends occur at it = 730*y + 2*cum[m]
    const int y0 = (first - 1) / GNT, y1 = (it - 1) / GNT;
    for (int y = y0; y <= y1; ++y)
      for (int m = 0; m < 12; ++m) {
        const int e = y * GNT + 2 * cum[m];
        if (e >= first && e < it) ++rec;
      }
  }
  si.out_rec = rec;
  si.co2 = a.spinup ? mc.co2_flux : a.co2[(size_t)member * a.co2_stride + (it - 1) / GNT];  // f:924
  return si;
}

// ---------------------------------------------------------------------------------------------
// The whole member integration: `nsteps` steps starting at step counter it0.
// smem layout: [0, 2*GNC) halo double buffer, [2*GNC, 4*GNC) stash (tendA, tq).
// ---------------------------------------------------------------------------------------------
GDEV void member_run(const SimtCtx& ctx, const GrebKernelArgs& a, const GrebMemberConst& mc, int member) {
  float* hb = ctx.smem;
  float* stash = ctx.smem + 2 * GNC;
  const WarpGeom g = warp_geom(ctx, mc);
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  const float* wzg = a.wz + (size_t)mc.group * 2 * GNC;
  CircTile t;

  GNOUNROLL
  for (int it = a.it0; it < a.it0 + a.nsteps; ++it) {
    const StepInfo si = step_info(a, mc, member, it);
    const float* forc = a.forc + (size_t)si.ityr * GF_COUNT * GNC;

    // ---- phase A: column physics, Ts/To/cap update
    GNOUNROLL
    for (int r = 0; r < g.nr; ++r) column_phase_a(ctx, a, mc, member, si, g.k0 + r, g.col, stash);

    // ---- circulation of air temperature (f:301), then of humidity (f:303): one code instance
    circ_load_uv(t, g, forc + GF_U * GNC, forc + GF_V * GNC);
    GNOUNROLL
    for (int fld = 0; fld < 2; ++fld) {
      circ_load_wz(t, g, wzg + fld * GNC);
      circ_load_field(t, g, st + (fld == 0 ? GS_TA : GS_Q) * GNC);
      circulation_run(ctx, t, g, mc, hb);
      if (fld == 0) {
        GUNROLL
        for (int r = 0; r < GREB_MAXR; ++r)
          if (r < g.nr) column_phase_b(a, member, si, g.k0 + r, g.col, t.Y[r + 2], stash);
        cta_sync(ctx);  // halo buffers are reused by the next circulation
      } else {
        GUNROLL
        for (int r = 0; r < GREB_MAXR; ++r)
          if (r < g.nr) column_phase_c(a, mc, member, si, g.k0 + r, g.col, t.Y[r + 2], stash);
      }
    }

    // ---- output (f:975-985)
    if (!si.spinup && si.month_end) {
      GNOUNROLL
      for (int r = 0; r < g.nr; ++r) column_month_end(a, member, si, g.k0 + r, g.col);
    }

    // ---- annual mean diagnostics (f:948-956)
    if (si.ityr == GNT - 1) {
      float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
      cta_sync(ctx);
      GNOUNROLL
      for (int r = 0; r < g.nr; ++r) {
        GUNROLL
        for (int c = 0; c < 3; ++c) {
          const vi idx = (g.k0 + r) * GX + g.col + c;
          v_st(hb, idx, v_ld(acc + GA_TSMN * GNC, idx) / (float)GNT);  // f:949
          v_st(acc + GA_TSMN * GNC, idx, v_bcast(0.0f));               // f:955
        }
      }
      cta_sync(ctx);
      if (ctx.warp == 0 && lane0(ctx)) {
        float s = 0.0f, sw = 0.0f;
        for (int k = 0; k < GY; ++k) {
          float rs = 0.0f;
          for (int i = 0; i < GX; ++i) {
            s = s + hb[k * GX + i];  // f:954 sum() in array element order
            rs = rs + hb[k * GX + i];
          }
          sw = sw + a.coslat_w[k] * (rs / (float)GX);
        }
        const float g0 = s / (float)(GX * GY) - 273.15f;
        a.diag[member * 2 + 0] = g0;
        a.diag[member * 2 + 1] = sw - 273.15f;
        if (!(g0 == g0) || g0 > 1e4f || g0 < -1e4f) a.flags[member] = 1;
      }
    }
    cta_sync(ctx);  // next step's phase A may overwrite the stash / reuse hb
  }
}
