// greb_core.h — warp-level implementation of the GREB 12-hourly step for one ensemble member.
//
// One CTA integrates one member.  Thread mapping ("v2"):
//   * 12 main warps.  A lane group of 8 lanes owns one latitude row; each lane owns 12 consecutive
//     longitudes of it, so every thread runs the SAME straight-line code on a 12-cell register tile
//     (the hot loop is ~10 KB of SASS instead of an unrolled per-row body per warp).  The periodic
//     longitude wrap is a rotation inside the 8-lane group (SHFL); the rows above and below come from
//     a double-buffered shared-memory copy of the field (LDS.128), published once per sub-step.
//   * 2 helper warps own the circulation of the two pole rows and of any other row whose polar
//     x-diffusion needs several sub-sub-steps (time2_diff > 1: 8 dependent iterations on the pole
//     rows at the default kappa): a whole row per warp at 3 cells per lane, so that serial chain
//     runs inside one warp, in parallel with the other 46 rows, instead of stalling a main warp.
//   * ONE CTA barrier per sub-step: a thread publishes its row, does the whole x-direction part of
//     the next sub-step (which needs only its own row) and only then synchronises (BAR.SYNC; a
//     polled split-phase mbarrier variant is kept behind GREB_BAR_MBARRIER, greb_simt.h).
//
// Arithmetic contract ("exact mode"): the reference is gfortran -O3 without -ffast-math, i.e.
// IEEE fp32, no FMA contraction, expression order as written.  This file is compiled with
// -fmad=false; v_fma is used only where it is provably identical to the written form
// (multiplication by 4 is exact), and the divisions by the literals 3. and 20. use a
// correctly-rounded 3-instruction sequence.  Identities used to share work between cells —
// x-(y) == x+(-y), (-a)*b == -(a*b), RN(-x) == -RN(x), a+b == b+a, x+0 == x — are exact in IEEE
// arithmetic, so results are bit-identical to the as-written evaluation (signs of zeros aside).
//
// Reference: /root/reference/src/greb.f90 ("f:NNN" below).
#pragma once

#include "greb_simt.h"
#include "greb_types.h"
#include "../../include/greb_b200.h"  // GREB_SW_*

#define GREB_DT 43200.0f  // f:38  (integer dt in real expressions)

// ---- correctly rounded x/3 and x/20 ---------------------------------------------------------
// q0 = x*RN(1/d); r = fma(-d,q0,x); q = fma(r,RN(1/d),q0) equals RN(x/d) for every float x whose
// quotient is a normal number (verified exhaustively over all 2^32 inputs, tests/test_divc.py and
// tools/divc_exhaustive.c).
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

// =============================================================================================
//                    main warps: one 12-cell tile of one latitude row per thread
// =============================================================================================

struct RowGeom {
  int k;           // latitude row (uniform within the 8-lane group)
  int polar;       // f:592 / f:799 branch
  int ykind;       // 0 interior, 1: k==0, 2: k==1, 3: k==GY-2, 4: k==GY-1   (f:756-795, f:587-590)
  int owned;       // 0 if the circulation of this row runs on a helper warp (the group then only does the column phases)
  int km2, km1, kp1, kp2;  // neighbour rows clamped into the grid (absent rows get zero weights)
  vi col;          // first owned longitude: 12 * (lane & 7)
  vi lane_l, lane_r;  // SHFL sources: the lanes owning the 12 cells to the west / east
  vb is_bug;       // the lane group member that owns longitude 93 (its cell 9), see f:881
  vi tid4;         // 4 * (index of this thread among the 384 main threads): private shared-memory slot
};

#if GREB_DEVICE
GDEV bool ctx_is_helper(const SimtCtx& c) { return c.warp >= GREB_NMAIN; }
GDEV int ctx_helper_index(const SimtCtx& c) { return c.warp - GREB_NMAIN; }
GDEV int ctx_group(const SimtCtx& c) { return c.warp * 4 + (c.lane_u >> 3); }
GDEV RowGeom row_geom(const SimtCtx& c, const GrebMemberConst& mc) {
  RowGeom g;
  const int seg = c.lane_u & 7, base = c.lane_u & 24;
  g.k = mc.row_of_group[ctx_group(c)];
  g.col = 12 * seg;
  g.lane_l = base | ((seg + 7) & 7);
  g.lane_r = base | ((seg + 1) & 7);
  g.is_bug = (seg == 7);
  g.tid4 = 4 * (c.warp * 32 + c.lane_u);
#else
GDEV bool ctx_is_helper(const SimtCtx& c) { return c.warp >= GY; }       // emu: unit index
GDEV int ctx_helper_index(const SimtCtx& c) { return c.warp - GY; }
GDEV int ctx_group(const SimtCtx& c) { return c.warp; }
GDEV RowGeom row_geom(const SimtCtx& c, const GrebMemberConst& mc) {
  RowGeom g;
  const vi seg = ctx_lane(c) & 7;   // emu: lanes 0..7 of the vector are the group, the rest is unused
  g.k = mc.row_of_group[ctx_group(c)];
  g.col = seg * 12;
  g.lane_l = (seg + 7) & 7;
  g.lane_r = (seg + 1) & 7;
  g.is_bug = (seg == 7);
  g.tid4 = (seg + 8 * ctx_group(c)) * 4;
#endif
  g.polar = mc.polar[g.k];
  g.ykind = g.k == 0 ? 1 : g.k == 1 ? 2 : g.k == GY - 2 ? 3 : g.k == GY - 1 ? 4 : 0;
  g.owned = mc.hslot_of_row[g.k] < 0;
  g.km2 = g.k >= 2 ? g.k - 2 : 0;
  g.km1 = g.k >= 1 ? g.k - 1 : 0;
  g.kp1 = g.k <= GY - 2 ? g.k + 1 : GY - 1;
  g.kp2 = g.k <= GY - 3 ? g.k + 2 : GY - 1;
  return g;
}

struct Tile {
  vf T[GREB_CPT];        // the circulating field, own cells
  vf W[GREB_CPT];        // wz, own cells
  vf U[GREB_CPT];        // zonal wind of this step
  vf wxl[3], wxr[3];     // wz of the 3 cells west / east of the tile
  vi pvmask;             // GREB_YCOEF (fast mode): bit j set <=> v >= 0 at own cell j
  // The y-direction constants live in a private shared-memory slot of the thread (GSM_PRIV), not in
  // registers: V, wz(k-1), wz(k+1) (0 where the row does not exist) and the upstream far-row weight
  // WFY = v>=0 ? wz(k-2) : wz(k+2) (0 if absent).  They are read back with LDS.128 in substep_y.
};

GDEV float* priv_ptr(float* smem, int which, int q) { return smem + GSM_PRIV + (which * 3 + q) * (GREB_NMAIN * 32 * 4); }

// `uscale` = 1 in the exact mode; -cadv in the fast mode (Tile::U then holds CU, see substep_x_fast).
// `vA`, `vB` = 1 in the exact mode; in the fast mode the latitudinal advection coefficient for v >= 0 /
// v < 0, so that the private V slot holds v*cy: its sign selects the wind branch, -|v*cy| is the factor.
GDEV void tile_load_uv(Tile& t, const RowGeom& g, const float* u, const float* v, float* smem, float uscale = 1.0f,
                       float vA = 1.0f, float vB = 1.0f) {
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf a[4], b[4];
    v_ldg4(a, u, g.k * GX + g.col + 4 * q);
    v_ldg4(b, v, g.k * GX + g.col + 4 * q);
    GUNROLL
    for (int i = 0; i < 4; ++i) {
      t.U[4 * q + i] = (uscale == 1.0f) ? a[i] : a[i] * uscale;
      if (vA != 1.0f || vB != 1.0f) b[i] = b[i] * v_sel(b[i] >= 0.0f, v_bcast(vA), v_bcast(vB));
    }
    v_st4(priv_ptr(smem, PRIV_V, q), g.tid4, b[0], b[1], b[2], b[3]);
  }
}

// wz-dependent constants of one circulation (needs V of tile_load_uv)
GDEV void tile_load_wz(Tile& t, const RowGeom& g, const float* wz, float* smem) {
  const bool has_m1 = g.k >= 1, has_m2 = g.k >= 2, has_p1 = g.k <= GY - 2, has_p2 = g.k <= GY - 3;
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf w0[4], wm1[4], wp1[4], wm2[4], wp2[4], v[4], wfy[4];
    v_ldg4(w0, wz, g.k * GX + g.col + 4 * q);
    v_ldg4(wm1, wz, g.km1 * GX + g.col + 4 * q);
    v_ldg4(wp1, wz, g.kp1 * GX + g.col + 4 * q);
    v_ldg4(wm2, wz, g.km2 * GX + g.col + 4 * q);
    v_ldg4(wp2, wz, g.kp2 * GX + g.col + 4 * q);
    v_ld4(v, priv_ptr(smem, PRIV_V, q), g.tid4);
    GUNROLL
    for (int i = 0; i < 4; ++i) {
      t.W[4 * q + i] = w0[i];
      if (!has_m1) wm1[i] = v_bcast(0.0f);
      if (!has_p1) wp1[i] = v_bcast(0.0f);
      const vf a = has_m2 ? wm2[i] : v_bcast(0.0f);
      const vf b = has_p2 ? wp2[i] : v_bcast(0.0f);
      wfy[i] = v_sel(v[i] >= 0.0f, a, b);
    }
    v_st4(priv_ptr(smem, PRIV_WM1, q), g.tid4, wm1[0], wm1[1], wm1[2], wm1[3]);
    v_st4(priv_ptr(smem, PRIV_WP1, q), g.tid4, wp1[0], wp1[1], wp1[2], wp1[3]);
    v_st4(priv_ptr(smem, PRIV_WFY, q), g.tid4, wfy[0], wfy[1], wfy[2], wfy[3]);
  }
  GUNROLL
  for (int i = 0; i < 3; ++i) {
    // periodic wrap of the longitude index
    const vi cl = v_seli(g.col == 0, vi(GX - 3 + i), g.col - 3 + i);
    const vi cr = v_seli(g.col == GX - 12, vi(i), g.col + 12 + i);
    t.wxl[i] = v_ldg(wz, g.k * GX + cl);
    t.wxr[i] = v_ldg(wz, g.k * GX + cr);
  }
}

#if GREB_YCOEF
// Fast mode, once per circulation: fold the latitudinal constants of the step into three per-cell factors
//   T' = T + X + CA*(T(k+1)-T) + CB*(T(k-1)-T) + CF*(Tfar-T),   X = wz*dTx + aTx from the x part,
//   CA = wz*ccy_diff*wz(k+1) + (v<0  ? |v*cy|*wz(k+1) : 0)      f:587-588 + the v<0 branch of f:771-780
//   CB = wz*ccy_diff*wz(k-1) + (v>=0 ? |v*cy|*wz(k-1) : 0)
//   CF = |v*cy|*WFY,  Tfar = v>=0 ? T(k-2) : T(k+2)
// (overwrites the PRIV_WP1 / PRIV_WM1 / PRIV_WFY slots; PRIV_V is no longer read in the sub-steps)
GDEV void tile_fold_ycoef(Tile& t, const RowGeom& g, float ccyd, float* smem) {
  t.pvmask = vi(0);
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf V[4], Wm1[4], Wp1[4], WFY[4], CA[4], CB[4], CF[4];
    v_ld4(V, priv_ptr(smem, PRIV_V, q), g.tid4);
    v_ld4(Wm1, priv_ptr(smem, PRIV_WM1, q), g.tid4);
    v_ld4(Wp1, priv_ptr(smem, PRIV_WP1, q), g.tid4);
    v_ld4(WFY, priv_ptr(smem, PRIV_WFY, q), g.tid4);
    GUNROLL
    for (int i = 0; i < 4; ++i) {
      const vb pv = V[i] >= 0.0f;
      const vf av = v_abs(V[i]);
      const vf wd = t.W[4 * q + i] * ccyd;
      CA[i] = v_fma(wd, Wp1[i], v_sel(pv, v_bcast(0.0f), av * Wp1[i]));
      CB[i] = v_fma(wd, Wm1[i], v_sel(pv, av * Wm1[i], v_bcast(0.0f)));
      CF[i] = av * WFY[i];
      t.pvmask = v_seli(pv, t.pvmask | (1 << (4 * q + i)), t.pvmask);
    }
    v_st4(priv_ptr(smem, PRIV_WP1, q), g.tid4, CA[0], CA[1], CA[2], CA[3]);
    v_st4(priv_ptr(smem, PRIV_WM1, q), g.tid4, CB[0], CB[1], CB[2], CB[3]);
    v_st4(priv_ptr(smem, PRIV_WFY, q), g.tid4, CF[0], CF[1], CF[2], CF[3]);
  }
}
#endif

GDEV void tile_load_field(Tile& t, const RowGeom& g, const float* X) {
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf a[4];
    v_ld4(a, X, g.k * GX + g.col + 4 * q);
    GUNROLL
    for (int i = 0; i < 4; ++i) t.T[4 * q + i] = a[i];
  }
}

GDEV void tile_publish(const Tile& t, const RowGeom& g, float* buf) {
  GUNROLL
  for (int q = 0; q < 3; ++q)
    v_st4(buf, g.k * GX + g.col + 4 * q, t.T[4 * q], t.T[4 * q + 1], t.T[4 * q + 2], t.T[4 * q + 3]);
}

// ---------------------------------------------------------------------------------------------
// x-direction part of one sub-step (needs only the thread's own row):
//   dTx = longitudinal diffusion increment (f:592-719), aTx = longitudinal advection (f:798-911)
// Notation: d(m) = T(m+1)-T(m),  P(m) = wz(m)*d(m),  Q(m) = wz(m+1)*d(m); arrays are indexed
// e = m + 3 where m is the cell index relative to the tile (m = -3 .. 14).
// ---------------------------------------------------------------------------------------------
GDEV void substep_x(vf (&dTx)[GREB_CPT], vf (&aTx)[GREB_CPT], const Tile& t, const RowGeom& g,
                    const GrebMemberConst& mc) {
  vf TT[18], WW[18];
  GUNROLL
  for (int i = 0; i < 3; ++i) {
    TT[i] = v_shfl(t.T[9 + i], g.lane_l);
    TT[15 + i] = v_shfl(t.T[i], g.lane_r);
    WW[i] = t.wxl[i];
    WW[15 + i] = t.wxr[i];
  }
  GUNROLL
  for (int i = 0; i < GREB_CPT; ++i) {
    TT[3 + i] = t.T[i];
    WW[3 + i] = t.W[i];
  }
  vf d[17], P[17], Q[17], A[17], B[17], S[GREB_CPT];
  GUNROLL
  for (int e = 0; e < 17; ++e) d[e] = TT[e + 1] - TT[e];
  GUNROLL
  for (int e = 0; e <= 13; ++e) P[e] = WW[e] * d[e];
  GUNROLL
  for (int e = 3; e <= 16; ++e) Q[e] = WW[e + 1] * d[e];
  GUNROLL
  for (int e = 1; e <= 13; ++e) A[e] = P[e] - P[e - 1];  // A(m) = P(m)-P(m-1)
  GUNROLL
  for (int e = 3; e <= 15; ++e) B[e] = Q[e + 1] - Q[e];  // B(m) = Q(m+1)-Q(m)
  // bracket of f:620-625: 10*(Q(j)-P(j-1)) + 4*A(j-1) + 4*B(j) + A(j-2) + B(j+1)
  GUNROLL
  for (int j = 0; j < GREB_CPT; ++j) {
    const int e = j + 3;
    const vf G = Q[e] - P[e - 1];
    S[j] = v_fma(4.0f, B[e], v_fma(4.0f, A[e - 1], 10.0f * G)) + A[e - 2] + B[e + 1];
  }
  if (!g.polar) {
    const float cc = mc.ccx_diff[g.k], cca = mc.ccx_adv[g.k];
    GUNROLL
    for (int j = 0; j < GREB_CPT; ++j) {
      const int e = j + 3;
      dTx[j] = div20(cc * S[j]);
      // f:816-820: -um*(wz(j-1)*(T-T(j-1)) + wz(j-2)*(T-T(j-2))) + up*(wz(j+1)*(T-T(j+1)) + wz(j+2)*(T-T(j+2)))
      // exactly one of um, up is non-zero: select the upstream side, multiply by -|u|
      const vb pu = t.U[j] >= 0.0f;
      const vf near = v_sel(pu, P[e - 1], -Q[e]);
      const vf far = v_sel(pu, WW[e - 2], WW[e + 2]) * (TT[e] - v_sel(pu, TT[e - 2], TT[e + 2]));
      const vf Xu = (-v_abs(t.U[j])) * (near + far);
      aTx[j] = div3(cca * Xu);
    }
  } else {
    const float cc2 = mc.ccx2_diff[g.k], cca2 = mc.ccx2_adv[g.k];
    GUNROLL
    for (int j = 0; j < GREB_CPT; ++j) {
      const int e = j + 3;
      {  // f:659-718 with time2 == 1 (rows with more sub-sub-steps are overwritten from the helper)
        vf dd = div20(cc2 * S[j]);
        dd = polar_clamp(dd, t.T[j]);
        const vf h = t.T[j] + dd;
        dTx[j] = h - t.T[j];
      }
      // f:872-878: -um*(10*wz(j-1)*(T(j)-T(j-1)) + 4*wz(j-2)*(T(j-1)-T(j-2)) + wz(j-3)*(T(j-2)-T(j-3)))
      //            + up*(10*wz(j+1)*(T(j)-T(j+1)) + 4*wz(j+2)*(T(j+1)-T(j+2)) + wz(j+3)*(T(j+2)-T(j+3)))
      const vb pu = t.U[j] >= 0.0f;
      const vf near10 = (10.0f * v_sel(pu, WW[e - 1], WW[e + 1])) * v_sel(pu, d[e - 1], d[e]);
      vf mid = v_sel(pu, P[e - 2], Q[e + 1]);
      vf far = v_sel(pu, P[e - 3], Q[e + 2]);
      if (j == 9) {
        // f:881: at Fortran j = xdim-2 (0-based longitude 93 = cell 9 of the last lane of the group)
        // the reference sets jp1 = jp2 = xdim-1, jp3 = 1: the 4* term vanishes and the 1* term is
        // wz(1)*(T(xdim-1)-T(1)).  Reproduced verbatim.
        const vb bug = g.is_bug && !pu;
        mid = v_sel(bug, v_bcast(0.0f), mid);
        far = v_sel(bug, WW[15] * (TT[15] - TT[13]), far);
      }
      const vf Sa = v_fma(4.0f, mid, near10) + far;
      const vf Xu = (-t.U[j]) * Sa;
      vf dd = div20(cca2 * Xu);
      dd = polar_clamp(dd, t.T[j]);  // f:907
      const vf h = t.T[j] + dd;      // f:908
      aTx[j] = h - t.T[j];           // f:910
    }
  }
}

// ---------------------------------------------------------------------------------------------
// y-direction part + update: needs rows k-2..k+2 of the published field `buf`.
//   out = (T + wz*(dTx+dTy)) + (aTx+aTy)                                  (f:721, f:913, f:549)
// SPECIAL = the thread's row is row 2 or row ydim-1 (1-based), whose advection divides only one of
// the two wind branches by 3 (f:766-769, f:784-787).  Rows 1 and ydim are always helper-owned.
// ---------------------------------------------------------------------------------------------
template <bool SPECIAL>
GDEV void substep_y(Tile& t, const vf (&dTx)[GREB_CPT], const vf (&aTx)[GREB_CPT], const RowGeom& g,
                    const GrebMemberConst& mc, const float* buf, float* smem) {
  const float ccyd = mc.ccy_diff, ccya = mc.ccy_adv;
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf tm2[4], tm1[4], tp1[4], tp2[4], V[4], Wm1[4], Wp1[4], WFY[4];
    v_ld4(V, priv_ptr(smem, PRIV_V, q), g.tid4);
    v_ld4(Wm1, priv_ptr(smem, PRIV_WM1, q), g.tid4);
    v_ld4(Wp1, priv_ptr(smem, PRIV_WP1, q), g.tid4);
    v_ld4(WFY, priv_ptr(smem, PRIV_WFY, q), g.tid4);
    v_ld4(tm1, buf, g.km1 * GX + g.col + 4 * q);
    v_ld4(tp1, buf, g.kp1 * GX + g.col + 4 * q);
    v_ld4(tm2, buf, g.km2 * GX + g.col + 4 * q);
    v_ld4(tp2, buf, g.kp2 * GX + g.col + 4 * q);
    GUNROLL
    for (int i = 0; i < 4; ++i) {
      const int j = 4 * q + i;
      const vf T = t.T[j];
      const vf Pym1 = Wm1[i] * (T - tm1[i]);    // wz(k-1)*(T(k)-T(k-1))
      const vf Qy0 = Wp1[i] * (tp1[i] - T);     // wz(k+1)*(T(k+1)-T(k))
      // diffusion, latitudinal (f:587-588)
      const vf dTy = ccyd * (Qy0 - Pym1);
      // advection, latitudinal (f:771-780): upstream side selected by the sign of v
      const vb pv = V[i] >= 0.0f;
      const vf near = v_sel(pv, Pym1, -Qy0);
      const vf far = WFY[i] * (T - v_sel(pv, tm2[i], tp2[i]));
      const vf Xv = (-v_abs(V[i])) * (near + far);
      vf aTy;
      if (!SPECIAL) {
        aTy = div3(ccya * Xv);
      } else {
        // row 2: v>=0 branch is not divided at all, v<0 branch is divided before the ccy multiply;
        // row ydim-1: the other way round
        const vb plain = (g.ykind == 2) ? pv : !pv;
        aTy = v_sel(plain, ccya * Xv, ccya * div3(Xv));
      }
      const vf dXd = t.W[j] * (dTx[j] + dTy);   // f:721
      const vf dXa = aTx[j] + aTy;              // f:913
      t.T[j] = (T + dXd) + dXa;                 // f:549
    }
  }
}

// The same y part on PAIRS of adjacent cells (exact mode, rows without the special advection formula;
// build option GREB_PACKED_Y, greb_types.h): every add, subtract, multiply and fused multiply-add is one
// packed instruction for two cells (FADD2 / FMUL2 / FFMA2, greb_simt.h) — the IEEE operation on each half, so
// the results are the scalar code's bit for bit.  The y part has no coupling along the row, and LDS.128
// delivers two ready-made pairs, so no data is permuted; only the wind-branch selections stay scalar.
GDEV vf2 p_div3(vf2 x) {
#if GREB_DEVICE
  const vf2 r = p_bcast(0.3333333432674407958984375f);
  const vf2 q = p_mul(x, r);
  const vf2 e = p_fma(p_bcast(-3.0f), q, x);
  return p_fma(e, r, q);
#else
  return p_pack(div3(p_lo(x)), div3(p_hi(x)));
#endif
}
GDEV void substep_y_packed(Tile& t, const vf (&dTx)[GREB_CPT], const vf (&aTx)[GREB_CPT], const RowGeom& g,
                           const GrebMemberConst& mc, const float* buf, float* smem) {
  const vf2 ccyd = p_bcast(mc.ccy_diff), ccya = p_bcast(mc.ccy_adv);
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf2 tm2[2], tm1[2], tp1[2], tp2[2], V[2], Wm1[2], Wp1[2], WFY[2];
    p_ld2(V, priv_ptr(smem, PRIV_V, q), g.tid4);
    p_ld2(Wm1, priv_ptr(smem, PRIV_WM1, q), g.tid4);
    p_ld2(Wp1, priv_ptr(smem, PRIV_WP1, q), g.tid4);
    p_ld2(WFY, priv_ptr(smem, PRIV_WFY, q), g.tid4);
    p_ld2(tm1, buf, g.km1 * GX + g.col + 4 * q);
    p_ld2(tp1, buf, g.kp1 * GX + g.col + 4 * q);
    p_ld2(tm2, buf, g.km2 * GX + g.col + 4 * q);
    p_ld2(tp2, buf, g.kp2 * GX + g.col + 4 * q);
    GUNROLL
    for (int h = 0; h < 2; ++h) {
      const int j = 4 * q + 2 * h;
      const vf2 T = p_pack(t.T[j], t.T[j + 1]);
      const vf2 Pym1 = p_mul(Wm1[h], p_sub(T, tm1[h]));        // wz(k-1)*(T(k)-T(k-1))
      const vf2 Qy0 = p_mul(Wp1[h], p_sub(tp1[h], T));         // wz(k+1)*(T(k+1)-T(k))
      const vf2 dTy = p_mul(ccyd, p_sub(Qy0, Pym1));           // f:587-588
      const vb pv0 = p_lo(V[h]) >= 0.0f, pv1 = p_hi(V[h]) >= 0.0f;
      const vf2 NV = p_pack(-v_abs(p_lo(V[h])), -v_abs(p_hi(V[h])));
      const vf2 near = p_pack(v_sel(pv0, p_lo(Pym1), -p_lo(Qy0)), v_sel(pv1, p_hi(Pym1), -p_hi(Qy0)));
      const vf2 tfar = p_pack(v_sel(pv0, p_lo(tm2[h]), p_lo(tp2[h])), v_sel(pv1, p_hi(tm2[h]), p_hi(tp2[h])));
      const vf2 far = p_mul(WFY[h], p_sub(T, tfar));
      const vf2 Xv = p_mul(NV, p_add(near, far));              // (-|v|)*(near+far), f:771-780
      const vf2 aTy = p_div3(p_mul(ccya, Xv));
      const vf2 dXd = p_mul(p_pack(t.W[j], t.W[j + 1]), p_add(p_pack(dTx[j], dTx[j + 1]), dTy));   // f:721
      const vf2 dXa = p_add(p_pack(aTx[j], aTx[j + 1]), aTy);                                       // f:913
      const vf2 Tn = p_add(p_add(T, dXd), dXa);                                                     // f:549
      t.T[j] = p_lo(Tn);
      t.T[j + 1] = p_hi(Tn);
    }
  }
}

// =============================================================================================
//   FAST arithmetic mode (GREB_ARITH_FAST): the same stencils, algebraically factored and with FMA
//   contraction — NOT bit-identical to the reference, held to the north_star tolerances instead
//   (tests/test_gpu_fast_mode.py).  SURVEY.md A.3: the x-diffusion bracket of f:620-625 equals
//   6(Q(j)-P(j-1)) + 3(Q(j+1)-P(j-2)) + (Q(j+2)-P(j-3)); the divisions by 20 and 3 and the factors
//   ccx, ccy fold into one coefficient per row; (T+d)-T of the polar branches becomes d.  About 38
//   instead of 91 instructions per cell and sub-step.  The clamps of f:715/f:907 are kept.
// =============================================================================================
struct FastRow {
  float cdiff;     // ccx_diff/20 (main rows) or ccx2_diff/20 (polar rows)
  float cadv;      // ccx_adv/3 (main rows) or ccx2_adv/20 (polar rows)
  float cyA, cyB;  // latitudinal advection coefficient for v >= 0 / v < 0: ccy_adv/3 except on rows 2 and
                   // ydim-1, where one branch is not divided by 3 (f:766-769, f:784-787)
  float ccyd;
};
GDEV FastRow fast_row(int k, const GrebMemberConst& mc) {
  FastRow f;
  const int polar = mc.polar[k];
  f.cdiff = (polar ? mc.ccx2_diff[k] : mc.ccx_diff[k]) * 0.05f;
  f.cadv = polar ? mc.ccx2_adv[k] * 0.05f : mc.ccx_adv[k] * (1.0f / 3.0f);
  const float c3 = mc.ccy_adv * (1.0f / 3.0f);
  f.cyA = (k == 1) ? mc.ccy_adv : c3;
  f.cyB = (k == GY - 2) ? mc.ccy_adv : c3;
  f.ccyd = mc.ccy_diff;
  return f;
}

// In the fast mode Tile::U holds CU = -u*cadv: the coefficient of the active wind branch of f:816-820 /
// f:872-878 (u >= 0 <=> CU <= 0 selects the western branch).
GDEV void substep_x_fast(vf (&dTx)[GREB_CPT], vf (&aTx)[GREB_CPT], const Tile& t, const RowGeom& g, const FastRow& fr) {
  vf TT[18], WW[18];
  GUNROLL
  for (int i = 0; i < 3; ++i) {
    TT[i] = v_shfl(t.T[9 + i], g.lane_l);
    TT[15 + i] = v_shfl(t.T[i], g.lane_r);
    WW[i] = t.wxl[i];
    WW[15 + i] = t.wxr[i];
  }
  GUNROLL
  for (int i = 0; i < GREB_CPT; ++i) {
    TT[3 + i] = t.T[i];
    WW[3 + i] = t.W[i];
  }
  vf d[17], P[17], Q[17];
  GUNROLL
  for (int e = 0; e < 17; ++e) d[e] = TT[e + 1] - TT[e];
  GUNROLL
  for (int e = 0; e <= 13; ++e) P[e] = WW[e] * d[e];
  GUNROLL
  for (int e = 3; e <= 16; ++e) Q[e] = WW[e + 1] * d[e];
  // bracket = 6(Q(j)-P(j-1)) + 3(Q(j+1)-P(j-2)) + (Q(j+2)-P(j-3))
#define GREB_FAST_S(e) \
  v_fma(6.0f, Q[e], v_fma(-6.0f, P[(e) - 1], v_fma(3.0f, Q[(e) + 1], v_fma(-3.0f, P[(e) - 2], Q[(e) + 2] - P[(e) - 3]))))
  if (!g.polar) {
    GUNROLL
    for (int j = 0; j < GREB_CPT; ++j) {
      const int e = j + 3;
      dTx[j] = fr.cdiff * GREB_FAST_S(e);
      const vf SL = v_fma(WW[e - 2], TT[e] - TT[e - 2], P[e - 1]);
      const vf SR = v_fma(WW[e + 2], TT[e + 2] - TT[e], Q[e]);
      aTx[j] = t.U[j] * v_sel(t.U[j] <= 0.0f, SL, SR);
#if GREB_YCOEF
      dTx[j] = v_fma(t.W[j], dTx[j], aTx[j]);   // X = wz*dTx + aTx: one array crosses the barrier
#endif
    }
  } else {
    GUNROLL
    for (int j = 0; j < GREB_CPT; ++j) {
      const int e = j + 3;
      dTx[j] = polar_clamp(fr.cdiff * GREB_FAST_S(e), t.T[j]);                         // f:715
      const vf LL = v_fma(10.0f, P[e - 1], v_fma(4.0f, P[e - 2], P[e - 3]));
      vf RR = v_fma(10.0f, Q[e], v_fma(4.0f, Q[e + 1], Q[e + 2]));
      if (j == 9) RR = v_sel(g.is_bug, v_fma(10.0f, Q[e], WW[15] * (TT[15] - TT[13])), RR);   // f:881
      aTx[j] = polar_clamp(t.U[j] * v_sel(t.U[j] <= 0.0f, LL, RR), t.T[j]);          // f:907
#if GREB_YCOEF
      dTx[j] = v_fma(t.W[j], dTx[j], aTx[j]);
#endif
    }
  }
#undef GREB_FAST_S
}

GDEV void substep_y_fast(Tile& t, const vf (&dTx)[GREB_CPT], const vf (&aTx)[GREB_CPT], const RowGeom& g,
                         const FastRow& fr, const float* buf, float* smem) {
#if GREB_YCOEF
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf tm2[4], tm1[4], tp1[4], tp2[4], CA[4], CB[4], CF[4];
    v_ld4(CA, priv_ptr(smem, PRIV_WP1, q), g.tid4);
    v_ld4(CB, priv_ptr(smem, PRIV_WM1, q), g.tid4);
    v_ld4(CF, priv_ptr(smem, PRIV_WFY, q), g.tid4);
    v_ld4(tm1, buf, g.km1 * GX + g.col + 4 * q);
    v_ld4(tp1, buf, g.kp1 * GX + g.col + 4 * q);
    v_ld4(tm2, buf, g.km2 * GX + g.col + 4 * q);
    v_ld4(tp2, buf, g.kp2 * GX + g.col + 4 * q);
    GUNROLL
    for (int i = 0; i < 4; ++i) {
      const int j = 4 * q + i;
      const vf T = t.T[j];
      const vf far = v_sel(v_bit(t.pvmask, j), tm2[i], tp2[i]);
      const vf acc = v_fma(CA[i], tp1[i] - T, v_fma(CB[i], tm1[i] - T, T + dTx[j]));   // dTx holds X here
      t.T[j] = v_fma(CF[i], far - T, acc);
    }
  }
  (void)aTx; (void)fr;
  return;
#endif
  GUNROLL
  for (int q = 0; q < 3; ++q) {
    vf tm2[4], tm1[4], tp1[4], tp2[4], V[4], Wm1[4], Wp1[4], WFY[4];
    v_ld4(V, priv_ptr(smem, PRIV_V, q), g.tid4);
    v_ld4(Wm1, priv_ptr(smem, PRIV_WM1, q), g.tid4);
    v_ld4(Wp1, priv_ptr(smem, PRIV_WP1, q), g.tid4);
    v_ld4(WFY, priv_ptr(smem, PRIV_WFY, q), g.tid4);
    v_ld4(tm1, buf, g.km1 * GX + g.col + 4 * q);
    v_ld4(tp1, buf, g.kp1 * GX + g.col + 4 * q);
    v_ld4(tm2, buf, g.km2 * GX + g.col + 4 * q);
    v_ld4(tp2, buf, g.kp2 * GX + g.col + 4 * q);
    GUNROLL
    for (int i = 0; i < 4; ++i) {
      const int j = 4 * q + i;
      const vf T = t.T[j];
      const vf PyS = Wm1[i] * (T - tm1[i]);
      const vf QyN = Wp1[i] * (tp1[i] - T);
      const vb pv = V[i] >= 0.0f;
      const vf Sv = v_fma(WFY[i], T - v_sel(pv, tm2[i], tp2[i]), v_sel(pv, PyS, -QyN));
      const vf t1 = v_fma(t.W[j], v_fma(fr.ccyd, QyN - PyS, dTx[j]), T);   // T + wz*(dTx+dTy)
      t.T[j] = t1 + v_fma(-v_abs(V[i]), Sv, aTx[j]);                        // + (aTx+aTy); V holds v*cy
    }
  }
}

// =============================================================================================
//   helper warps: whole latitude rows at 3 cells per lane (the pole rows and every row whose polar
//   x-diffusion needs several sub-sub-steps).  Only polar-branch rows are ever helper-owned.
// =============================================================================================
struct XRow {
  vf xm1, xp1;                 // T(c0-1), T(c0+3)
  vf dm1, d0, d1, d2;          // d(c0-1..c0+2)
  vf Pm3, Pm2, Pm1, P0, P1;
  vf Q0, Q1, Q2, Qp1, Qp2;
};

// With c0 the lane's first cell: Pm3..P1 = P(c0-3..c0+1), Q0..Qp2 = Q(c0..c0+4).
GDEV void xrow_products(XRow& x, const vf (&T)[3], const vf (&W)[3], const vf (&WX)[4], vi lane_l, vi lane_r) {
  const vf xm2 = v_shfl(T[1], lane_l);
  x.xm1 = v_shfl(T[2], lane_l);
  x.xp1 = v_shfl(T[0], lane_r);
  const vf xp2 = v_shfl(T[1], lane_r);
  const vf dm2 = x.xm1 - xm2, dp1 = xp2 - x.xp1;
  x.dm1 = T[0] - x.xm1;
  x.d0 = T[1] - T[0];
  x.d1 = T[2] - T[1];
  x.d2 = x.xp1 - T[2];
  x.Pm2 = WX[0] * dm2;
  x.Pm1 = WX[1] * x.dm1;
  x.P0 = W[0] * x.d0;
  x.P1 = W[1] * x.d1;
  x.Q0 = W[1] * x.d0;
  x.Q1 = W[2] * x.d1;
  x.Q2 = WX[2] * x.d2;
  x.Qp1 = WX[3] * dp1;
  x.Pm3 = v_shfl(x.P0, lane_l);  // left lane's P(c0') = P(c0-3)
  x.Qp2 = v_shfl(x.Q1, lane_r);  // right lane's Q(c0'+1) = Q(c0+4)
}

GDEV void xdiff_bracket3(vf (&S)[3], const XRow& x) {
  const vf Am2 = x.Pm2 - x.Pm3, Am1 = x.Pm1 - x.Pm2, A0 = x.P0 - x.Pm1, A1 = x.P1 - x.P0;
  const vf B0 = x.Q1 - x.Q0, B1 = x.Q2 - x.Q1, B2 = x.Qp1 - x.Q2, B3 = x.Qp2 - x.Qp1;
  const vf G0 = x.Q0 - x.Pm1, G1 = x.Q1 - x.P0, G2 = x.Q2 - x.P1;
  S[0] = v_fma(4.0f, B0, v_fma(4.0f, Am1, 10.0f * G0)) + Am2 + B1;
  S[1] = v_fma(4.0f, B1, v_fma(4.0f, A0, 10.0f * G1)) + Am1 + B2;
  S[2] = v_fma(4.0f, B2, v_fma(4.0f, A1, 10.0f * G2)) + A0 + B3;
}

#define GREB_HROWS (GREB_MAXH / GREB_NHELP)  // rows per helper warp

struct HelperRow {
  vf T[3];                                   // the circulating field
  vf W[3], WX[4];                            // wz of the own cells and of c0-2, c0-1, c0+3, c0+4
  vf U[3], V[3];
  vf Wm1[3], Wp1[3], WFY[3];                 // as in the main tile (0 where the row does not exist)
  vf dTx[3], aTx[3];                         // x-direction results of the current sub-step
};

struct HelperGeom {
  int n;                       // rows served by this helper warp
  int k[GREB_HROWS];
  vi col, lane_l, lane_r;
  vb is_bug;                   // lane 31 owns longitude 93 as its first cell (f:881)
};

GDEV HelperGeom helper_geom(const SimtCtx& ctx, const GrebMemberConst& mc) {
  HelperGeom g;
  const vi lane = ctx_lane(ctx);
  g.col = lane * 3;
  g.lane_l = (lane + 31) & 31;
  g.lane_r = (lane + 1) & 31;
  g.is_bug = (lane == 31);
  g.n = 0;
  GUNROLL
  for (int i = 0; i < GREB_HROWS; ++i) {
    const int s = ctx_helper_index(ctx) + i * GREB_NHELP;
    g.k[i] = (s < mc.n_hslots) ? mc.helper_row[s] : 0;
    if (s < mc.n_hslots) g.n = i + 1;
  }
  return g;
}

GDEV void helper_load_uv(HelperRow (&hr)[GREB_HROWS], const HelperGeom& g, const float* u, const float* v) {
  GUNROLL
  for (int i = 0; i < GREB_HROWS; ++i)
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      hr[i].U[c] = v_ldg(u, g.k[i] * GX + g.col + c);
      hr[i].V[c] = v_ldg(v, g.k[i] * GX + g.col + c);
    }
}

GDEV void helper_load_wz(HelperRow (&hr)[GREB_HROWS], const HelperGeom& g, const float* wz) {
  const vi cm2 = v_seli(g.col == 0, vi(GX - 2), g.col - 2);
  const vi cm1 = v_seli(g.col == 0, vi(GX - 1), g.col - 1);
  const vi cp1 = v_seli(g.col == GX - 3, vi(0), g.col + 3);
  const vi cp2 = v_seli(g.col == GX - 3, vi(1), g.col + 4);
  GUNROLL
  for (int i = 0; i < GREB_HROWS; ++i) {
    const int k = g.k[i];
    const int km1 = k >= 1 ? k - 1 : 0, km2 = k >= 2 ? k - 2 : 0;
    const int kp1 = k <= GY - 2 ? k + 1 : GY - 1, kp2 = k <= GY - 3 ? k + 2 : GY - 1;
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      hr[i].W[c] = v_ldg(wz, k * GX + g.col + c);
      hr[i].Wm1[c] = (k >= 1) ? v_ldg(wz, km1 * GX + g.col + c) : v_bcast(0.0f);
      hr[i].Wp1[c] = (k <= GY - 2) ? v_ldg(wz, kp1 * GX + g.col + c) : v_bcast(0.0f);
      const vf a = (k >= 2) ? v_ldg(wz, km2 * GX + g.col + c) : v_bcast(0.0f);
      const vf b = (k <= GY - 3) ? v_ldg(wz, kp2 * GX + g.col + c) : v_bcast(0.0f);
      hr[i].WFY[c] = v_sel(hr[i].V[c] >= 0.0f, a, b);
    }
    hr[i].WX[0] = v_ldg(wz, k * GX + cm2);
    hr[i].WX[1] = v_ldg(wz, k * GX + cm1);
    hr[i].WX[2] = v_ldg(wz, k * GX + cp1);
    hr[i].WX[3] = v_ldg(wz, k * GX + cp2);
  }
}

// x-direction part for one helper-owned (polar) row: all time2 diffusion sub-sub-steps (f:655-718)
// and the polar advection (f:838-910)
GDEV void helper_x(HelperRow& r, const HelperGeom& g, const GrebMemberConst& mc, int k) {
  const float cc2 = mc.ccx2_diff[k], cca2 = mc.ccx2_adv[k];
#ifdef GREB_DBG_TIME2_1
  const int time2 = 1;  // timing experiment only (wrong results)
#else
  const int time2 = mc.time2_diff[k];
#endif
  XRow x;
  xrow_products(x, r.T, r.W, r.WX, g.lane_l, g.lane_r);
  vf S[3], h[3];
  xdiff_bracket3(S, x);
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    vf dd = div20(cc2 * S[c]);
    dd = polar_clamp(dd, r.T[c]);  // f:715
    h[c] = r.T[c] + dd;            // f:716
  }
  // polar advection bracket from the same products (f:872-878 + the f:881 index bug)
  {
    const vf dlo[3] = {x.dm1, x.d0, x.d1}, dhi[3] = {x.d0, x.d1, x.d2};
    const vf wlo[3] = {r.WX[1], r.W[0], r.W[1]}, whi[3] = {r.W[1], r.W[2], r.WX[2]};
    const vf P2[3] = {x.Pm2, x.Pm1, x.P0}, P3[3] = {x.Pm3, x.Pm2, x.Pm1};
    const vf Qn1[3] = {x.Q1, x.Q2, x.Qp1}, Qn2[3] = {x.Q2, x.Qp1, x.Qp2};
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      const vb pu = r.U[c] >= 0.0f;
      const vf near10 = (10.0f * v_sel(pu, wlo[c], whi[c])) * v_sel(pu, dlo[c], dhi[c]);
      vf mid = v_sel(pu, P2[c], Qn1[c]);
      vf far = v_sel(pu, P3[c], Qn2[c]);
      if (c == 0) {  // longitude 93 is cell 0 of lane 31: jp1 = jp2 = 94, jp3 = 0
        const vb bug = g.is_bug && !pu;
        mid = v_sel(bug, v_bcast(0.0f), mid);
        far = v_sel(bug, r.WX[2] * (x.xp1 - r.T[1]), far);
      }
      const vf Sa = v_fma(4.0f, mid, near10) + far;
      const vf Xu = (-r.U[c]) * Sa;
      vf dd = div20(cca2 * Xu);
      dd = polar_clamp(dd, r.T[c]);     // f:907
      const vf ha = r.T[c] + dd;        // f:908
      r.aTx[c] = ha - r.T[c];           // f:910
    }
  }
  GNOUNROLL
  for (int tt2 = 1; tt2 < time2; ++tt2) {
    XRow y;
    xrow_products(y, h, r.W, r.WX, g.lane_l, g.lane_r);
    xdiff_bracket3(S, y);
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      vf dd = div20(cc2 * S[c]);
      dd = polar_clamp(dd, h[c]);
      h[c] = h[c] + dd;
    }
  }
  GUNROLL
  for (int c = 0; c < 3; ++c) r.dTx[c] = h[c] - r.T[c];  // f:718
}

// y-direction part + update of one helper-owned row (all five row cases of f:756-795, f:587-590)
GDEV void helper_y(HelperRow& r, const HelperGeom& g, const GrebMemberConst& mc, int k, const float* buf) {
  const float ccyd = mc.ccy_diff, ccya = mc.ccy_adv;
  const int km1 = k >= 1 ? k - 1 : 0, km2 = k >= 2 ? k - 2 : 0;
  const int kp1 = k <= GY - 2 ? k + 1 : GY - 1, kp2 = k <= GY - 3 ? k + 2 : GY - 1;
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vf T = r.T[c];
    const vf tm1 = v_ld(buf, km1 * GX + g.col + c), tp1 = v_ld(buf, kp1 * GX + g.col + c);
    const vf tm2 = v_ld(buf, km2 * GX + g.col + c), tp2 = v_ld(buf, kp2 * GX + g.col + c);
    const vf Pym1 = r.Wm1[c] * (T - tm1);
    const vf Qy0 = r.Wp1[c] * (tp1 - T);
    vf dTy = ccyd * (Qy0 - Pym1);
    if (k == 0) dTy = (ccyd * r.Wp1[c]) * (tp1 - T);       // f:589
    if (k == GY - 1) dTy = (ccyd * r.Wm1[c]) * (tm1 - T);  // f:590
    const vb pv = r.V[c] >= 0.0f;
    const vf near = v_sel(pv, Pym1, -Qy0);
    const vf far = r.WFY[c] * (T - v_sel(pv, tm2, tp2));
    const vf Xv = (-v_abs(r.V[c])) * (near + far);
    vf aTy = div3(ccya * Xv);                                              // rows 1, 3..ydim-2, ydim
    if (k == 1) aTy = v_sel(pv, ccya * Xv, ccya * div3(Xv));               // f:766-769
    if (k == GY - 2) aTy = v_sel(pv, ccya * div3(Xv), ccya * Xv);          // f:784-787
    const vf dXd = r.W[c] * (r.dTx[c] + dTy);
    const vf dXa = r.aTx[c] + aTy;
    r.T[c] = (T + dXd) + dXa;
  }
}

// ---- FAST arithmetic mode, helper rows ----------------------------------------------------------
GDEV void xdiff_bracket3_fast(vf (&S)[3], const XRow& x) {
  S[0] = v_fma(6.0f, x.Q0 - x.Pm1, v_fma(3.0f, x.Q1 - x.Pm2, x.Q2 - x.Pm3));
  S[1] = v_fma(6.0f, x.Q1 - x.P0, v_fma(3.0f, x.Q2 - x.Pm1, x.Qp1 - x.Pm2));
  S[2] = v_fma(6.0f, x.Q2 - x.P1, v_fma(3.0f, x.Qp1 - x.P0, x.Qp2 - x.Pm1));
}

GDEV void helper_x_fast(HelperRow& r, const HelperGeom& g, const GrebMemberConst& mc, int k) {
  const FastRow fr = fast_row(k, mc);
  const int time2 = mc.time2_diff[k];
  XRow x;
  xrow_products(x, r.T, r.W, r.WX, g.lane_l, g.lane_r);
  vf S[3], h[3];
  xdiff_bracket3_fast(S, x);
  GUNROLL
  for (int c = 0; c < 3; ++c) h[c] = r.T[c] + polar_clamp(fr.cdiff * S[c], r.T[c]);
  {
    const vf P1_[3] = {x.Pm1, x.P0, x.P1}, P2[3] = {x.Pm2, x.Pm1, x.P0}, P3[3] = {x.Pm3, x.Pm2, x.Pm1};
    const vf Q0_[3] = {x.Q0, x.Q1, x.Q2}, Qn1[3] = {x.Q1, x.Q2, x.Qp1}, Qn2[3] = {x.Q2, x.Qp1, x.Qp2};
    GUNROLL
    for (int c = 0; c < 3; ++c) {
      const vb pu = r.U[c] >= 0.0f;
      const vf cu = (-r.U[c]) * fr.cadv;
      const vf LL = v_fma(10.0f, P1_[c], v_fma(4.0f, P2[c], P3[c]));
      vf RR = v_fma(10.0f, Q0_[c], v_fma(4.0f, Qn1[c], Qn2[c]));
      if (c == 0) RR = v_sel(g.is_bug, v_fma(10.0f, Q0_[c], r.WX[2] * (x.xp1 - r.T[1])), RR);   // f:881
      r.aTx[c] = polar_clamp(cu * v_sel(pu, LL, RR), r.T[c]);
    }
  }
  GNOUNROLL
  for (int tt2 = 1; tt2 < time2; ++tt2) {
    XRow y;
    xrow_products(y, h, r.W, r.WX, g.lane_l, g.lane_r);
    xdiff_bracket3_fast(S, y);
    GUNROLL
    for (int c = 0; c < 3; ++c) h[c] = h[c] + polar_clamp(fr.cdiff * S[c], h[c]);
  }
  GUNROLL
  for (int c = 0; c < 3; ++c) r.dTx[c] = h[c] - r.T[c];
}

GDEV void helper_y_fast(HelperRow& r, const HelperGeom& g, const GrebMemberConst& mc, int k, const float* buf) {
  const FastRow fr = fast_row(k, mc);
  const int km1 = k >= 1 ? k - 1 : 0, km2 = k >= 2 ? k - 2 : 0;
  const int kp1 = k <= GY - 2 ? k + 1 : GY - 1, kp2 = k <= GY - 3 ? k + 2 : GY - 1;
  GUNROLL
  for (int c = 0; c < 3; ++c) {
    const vf T = r.T[c];
    const vf tm1 = v_ld(buf, km1 * GX + g.col + c), tp1 = v_ld(buf, kp1 * GX + g.col + c);
    const vf tm2 = v_ld(buf, km2 * GX + g.col + c), tp2 = v_ld(buf, kp2 * GX + g.col + c);
    const vf PyS = r.Wm1[c] * (T - tm1);      // 0 on row 1 (Wm1 = 0)
    const vf QyN = r.Wp1[c] * (tp1 - T);      // 0 on row ydim
    const vb pv = r.V[c] >= 0.0f;
    const vf Sv = v_fma(r.WFY[c], T - v_sel(pv, tm2, tp2), v_sel(pv, PyS, -QyN));
    const vf cV = (-v_abs(r.V[c])) * v_sel(pv, v_bcast(fr.cyA), v_bcast(fr.cyB));
    const vf t1 = v_fma(r.W[c], v_fma(fr.ccyd, QyN - PyS, r.dTx[c]), T);
    r.T[c] = t1 + v_fma(cV, Sv, r.aTx[c]);
  }
}

// =============================================================================================
//                                     circulation drivers
// =============================================================================================
struct SyncState {
  SplitBar* bar;
  float* hb;    // [2][GNC]
  float* smem;  // CTA shared memory base (private slots)
  int phase;    // number of completed waits (runs on across circulations and steps)
};

// circulation (f:528-553) for a main-warp thread: on entry t.T holds X_in of the own cells, on exit
// X after the 24 sub-steps (for rows whose circulation runs on a helper warp: read back from the
// published field).  Every warp of the CTA (helpers via circulation_helper) must take part.
template <int MODE = 0>
GDEV void circulation_main(const SimtCtx& ctx, Tile& t, const RowGeom& g, const GrebMemberConst& mc, SyncState& ss) {
  const FastRow fr = fast_row(g.k, mc);
  const bool late = warp_uniform(ctx.late) != 0;   // stagger, greb_types.h
  if (g.owned) tile_publish(t, g, ss.hb + (ss.phase & 1) * GNC);
  sb_arrive(ctx, ss.bar);
#if defined(GREB_DBG_CLOCKS) && GREB_DEVICE
  // timing experiment (tests/debug/debug_circ.py): cycles per sub-step in the x part, the barrier and the y part
  long long c_x = 0, c_w = 0, c_y = 0;
#define GCLK(acc, t0) { const long long t1_ = clock64(); acc += t1_ - t0; t0 = t1_; }
  long long tc = clock64();
#else
#define GCLK(acc, t0)
#endif
  GNOUNROLL
  for (int tt = 0; tt < GSUB; ++tt) {
    vf dTx[GREB_CPT], aTx[GREB_CPT];
#ifdef GREB_DBG_MAIN_IDLE   // timing experiment only (wrong results): main warps just synchronise
    sb_wait(ctx, ss.bar, ss.phase);
    const float* buf = ss.hb + (ss.phase & 1) * GNC;
    (void)buf; (void)dTx; (void)aTx;
#else
    if (late) sb_wait(ctx, ss.bar, ss.phase);         // stagger: this warp's x part runs beside the others' y part
    if (MODE == 1) substep_x_fast(dTx, aTx, t, g, fr);
    else substep_x(dTx, aTx, t, g, mc);               // own row only: overlaps the barrier latency
    GCLK(c_x, tc)
    if (!late) sb_wait(ctx, ss.bar, ss.phase);
    GCLK(c_w, tc)
    const float* buf = ss.hb + (ss.phase & 1) * GNC;
    if (MODE == 1) substep_y_fast(t, dTx, aTx, g, fr, buf, ss.smem);
#if GREB_PACKED_Y
    else if (g.ykind == 0) substep_y_packed(t, dTx, aTx, g, mc, buf, ss.smem);
#else
    else if (g.ykind == 0) substep_y<false>(t, dTx, aTx, g, mc, buf, ss.smem);
#endif
    else substep_y<true>(t, dTx, aTx, g, mc, buf, ss.smem);
#endif
    ss.phase++;
    if (g.owned) tile_publish(t, g, ss.hb + (ss.phase & 1) * GNC);   // also after the last sub-step
    sb_arrive(ctx, ss.bar);
    GCLK(c_y, tc)
  }
#if defined(GREB_DBG_CLOCKS) && GREB_DEVICE
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ss.phase < 60)
    printf("warp %2d row %2d polar %d: x %lld wait %lld y %lld cycles per sub-step\n", (int)(threadIdx.x >> 5), g.k,
           g.polar, c_x / GSUB, c_w / GSUB, c_y / GSUB);
#endif
  sb_wait(ctx, ss.bar, ss.phase);   // everybody's final rows are published
  if (!g.owned) tile_load_field(t, g, ss.hb + (ss.phase & 1) * GNC);
  ss.phase++;
}

template <int MODE = 0>
GDEV void circulation_helper(const SimtCtx& ctx, HelperRow (&hr)[GREB_HROWS], const HelperGeom& g,
                             const GrebMemberConst& mc, const float* X, SyncState& ss) {
  GUNROLL
  for (int i = 0; i < GREB_HROWS; ++i)
    if (i < g.n) {
      GUNROLL
      for (int c = 0; c < 3; ++c) {
        hr[i].T[c] = v_ld(X, g.k[i] * GX + g.col + c);
        v_st(ss.hb + (ss.phase & 1) * GNC, g.k[i] * GX + g.col + c, hr[i].T[c]);
      }
    }
  sb_arrive(ctx, ss.bar);
#if defined(GREB_DBG_CLOCKS) && GREB_DEVICE
  long long c_x = 0, c_w = 0, c_y = 0;
  long long tc = clock64();
#endif
  GNOUNROLL
  for (int tt = 0; tt < GSUB; ++tt) {
    GUNROLL
    for (int i = 0; i < GREB_HROWS; ++i)
      if (i < g.n) {
        if (MODE == 1) helper_x_fast(hr[i], g, mc, g.k[i]);
        else helper_x(hr[i], g, mc, g.k[i]);
      }
    GCLK(c_x, tc)
    sb_wait(ctx, ss.bar, ss.phase);
    GCLK(c_w, tc)
    const float* buf = ss.hb + (ss.phase & 1) * GNC;
    ss.phase++;
    float* nxt = ss.hb + (ss.phase & 1) * GNC;
    GUNROLL
    for (int i = 0; i < GREB_HROWS; ++i)
      if (i < g.n) {
        if (MODE == 1) helper_y_fast(hr[i], g, mc, g.k[i], buf);
        else helper_y(hr[i], g, mc, g.k[i], buf);
        GUNROLL
        for (int c = 0; c < 3; ++c) v_st(nxt, g.k[i] * GX + g.col + c, hr[i].T[c]);
      }
    sb_arrive(ctx, ss.bar);
    GCLK(c_y, tc)
  }
#if defined(GREB_DBG_CLOCKS) && GREB_DEVICE
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ss.phase < 60)
    printf("helper warp %2d row %2d: x %lld wait %lld y %lld cycles per sub-step\n", (int)(threadIdx.x >> 5), g.k[0],
           c_x / GSUB, c_w / GSUB, c_y / GSUB);
#endif
  sb_wait(ctx, ss.bar, ss.phase);
  ss.phase++;
}

// =============================================================================================
//        column physics + state update (everything of time_loop / qflux_correction that is not
//        the circulation), 4 consecutive cells at a time
// =============================================================================================
GDEV vf pow4(vf x) {
  const vf x2 = x * x;
  return x2 * x2;  // gfortran expands x**4 as (x*x)*(x*x)
}

struct StepInfo {
  int ityr;       // 0-based step of year
  int month_end;  // 1 if a month ends at this step (f:975-976)
  float ndm;      // days of that month * 2
  int out_rec;    // month slot of this launch to write
  float co2;
  int spinup;
};

// Phase A (before the circulations): SW, LW, sensible, hydro, deep ocean; Ts/To/cap_surf update,
// flux corrections in spin-up mode; stashes the air-temperature and humidity tendencies.
// MODE 1 (GREB_ARITH_FAST): approximate division and transcendentals (2 ulp); MODE 0: IEEE division and
// the CUDA libm, like the reference's true divisions
template <int MODE, int SW>
GDEV void column_phase_a(const GrebKernelArgs& a, const GrebMemberConst& mc, int member, const StepInfo& si, int k,
                         vi idx0, float* stash, const float* corr_s) {
#define DIVF(x, y) (MODE == 1 ? v_div_fast((x), (y)) : (x) / (y))
#define DIVC(x, c, rc) (MODE == 1 ? v_div_fast((x), (c)) : v_divc((x), (c), (rc)))   // division by a member constant
#define LOGF(x) (MODE == 1 ? v_log_fast(x) : v_log(x))
#define EXPF(x) (MODE == 1 ? v_exp_fast(x) : v_exp(x))
  const float* forc = a.forc + (size_t)si.ityr * GF_COUNT * GNC;
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  const float* wz_air = a.wz + (size_t)mc.group * 2 * GNC;
  float* corr = a.corr + ((size_t)mc.group * GNT + si.ityr) * GC_COUNT * GNC;
  const float solar = a.sw_solar[si.ityr * GY + k];
  const float* pe = mc.p_emi;

  vf Ts4[4], Ta4[4], To4[4], q4[4], cap4[4], cld4[4], dTrad4[4], swet4[4], absw4[4], mld4[4], dmld4[4], zoc4[4], ez4[4];
  vf rdeep4[4], rmix4[4];
  vf c1[4], c2[4], tsmn4[4], tmm4[4], tomm4[4], apmm4[4];
  vi mask4[4];
  v_ld4(Ts4, st + GS_TS * GNC, idx0);
  v_ld4(Ta4, st + GS_TA * GNC, idx0);
  v_ld4(To4, st + GS_TO * GNC, idx0);
  v_ld4(q4, st + GS_Q * GNC, idx0);
  v_ld4(cap4, st + GS_CAP * GNC, idx0);
  v_ldg4(cld4, forc + GF_CLD * GNC, idx0);
  v_ldg4(dTrad4, forc + GF_DTRAD * GNC, idx0);
  v_ldg4(swet4, forc + GF_SWET * GNC, idx0);
  v_ldg4(absw4, forc + GF_ABSWIND * GNC, idx0);
  v_ldg4(mld4, forc + GF_MLD * GNC, idx0);
  v_ldg4(dmld4, forc + GF_DMLD * GNC, idx0);
  v_ldg4(rdeep4, forc + GF_RDEEP * GNC, idx0);
  v_ldg4(rmix4, forc + GF_RMIX * GNC, idx0);
  v_ldgi4(mask4, a.mask, idx0);
  v_ldg4(zoc4, a.z_ocean, idx0);
  v_ldg4(ez4, wz_air, idx0);
  if (!si.spinup) {
    v_ld4(c1, corr_s + GC_TF * GNC, idx0);    // GREB_TMA_CORR: shared memory, else the group's global slice
    v_ld4(c2, corr_s + GC_TOF * GNC, idx0);
    v_ld4(tmm4, acc + GA_TMM * GNC, idx0);
    v_ld4(tomm4, acc + GA_TOMM * GNC, idx0);
    v_ld4(apmm4, acc + GA_APMM * GNC, idx0);
  } else {
    v_ldg4(c1, a.tclim + (size_t)si.ityr * GNC, idx0);
    v_ldg4(c2, a.toclim, idx0);
  }
  v_ld4(tsmn4, acc + GA_TSMN * GNC, idx0);
  // process switches (greb.original.model.f90 log_exp experiments); sw == 0 is the full model
  const int sw_ = SW ? mc.switches : 0;  // SW == 0: the switch code is compiled out
  vf qcl4[4];
  if (sw_ & GREB_SW_LINEAR_VAPOR_EMISSIVITY) v_ldg4(qcl4, a.qclim + (size_t)si.ityr * GNC, idx0);
  if ((sw_ & GREB_SW_SST_PLUS_1K) && !si.spinup) {
    // orig:226 runs BEFORE time_loop updates the module variable ityr (orig:248), so the SST comes
    // from the climatology of the PREVIOUS step (step 730 for the first step of a year)
    vf tcl[4];
    v_ldg4(tcl, a.tclim + (size_t)((si.ityr + GNT - 1) % GNT) * GNC, idx0);
    GUNROLL
    for (int i = 0; i < 4; ++i) Ts4[i] = v_sel(v_bit(mask4[i], 1), tcl[i] + 1.0f, Ts4[i]);
  }

  vf Ts0o[4], To0o[4], capo4[4], tendA4[4], tq4[4], tfo[4], tofo[4], albo[4];
  GUNROLL
  for (int i = 0; i < 4; ++i) {
    const vf Ts = Ts4[i], Ta = Ta4[i], To = To4[i], q = q4[i], cap = cap4[i];
    const vf cld = cld4[i], dTrad = dTrad4[i], swet = swet4[i], absw = absw4[i], mld = mld4[i], dmld = dmld4[i];
    const vf zoc = zoc4[i], ez = ez4[i];
    const vb land_ge0 = v_bit(mask4[i], 0), ocean = v_bit(mask4[i], 1), glac = v_bit(mask4[i], 2);

    // ---- SWradiation, f:380-401
    const vf a_atmos = cld * mc.a_cloud;
    const float a_ice = mc.a_no_ice + mc.da_ice;
    const vf T1 = v_sel(land_ge0, v_bcast(mc.Tl_ice1), v_bcast(mc.To_ice1));
    const vf T2 = v_sel(land_ge0, v_bcast(mc.Tl_ice2), v_bcast(mc.To_ice2));
    vf a_surf = mc.a_no_ice + mc.da_ice * (1.0f - DIVC(Ts - T1, T2 - T1, v_seld(land_ge0, vd(mc.rc_alb_land), vd(mc.rc_alb_ocean))));
    a_surf = v_sel(Ts <= T1, v_bcast(a_ice), a_surf);
    a_surf = v_sel(Ts >= T2, v_bcast(mc.a_no_ice), a_surf);
    a_surf = v_sel(glac, v_bcast(a_ice), a_surf);
    if (sw_ & GREB_SW_NO_ICE_ALBEDO) a_surf = v_bcast(mc.a_no_ice);  // orig:394
    const vf albedo = a_surf + a_atmos - a_surf * a_atmos;
    const vf sw = solar * (1.0f - albedo);

    // ---- LWradiation, f:420-432
    const vf e_co2 = ez * si.co2;
    vf e_vapor = ez * mc.r_qviwv * q;
    if (sw_ & GREB_SW_LINEAR_VAPOR_EMISSIVITY) e_vapor = ez * mc.r_qviwv * qcl4[i];  // orig:423
    vf em = pe[3] * LOGF(pe[0] * e_co2 + pe[1] * e_vapor + pe[2]) + pe[6] + pe[4] * LOGF(pe[0] * e_co2 + pe[2]) +
            pe[5] * LOGF(pe[1] * e_vapor + pe[2]);
    em = DIVC(pe[7] - cld, v_bcast(pe[8]), vd(mc.rc_pe8)) * (em - pe[9]) + pe[9];
    if (sw_ & GREB_SW_LINEAR_VAPOR_EMISSIVITY)
      em = em + ((0.022f / (0.15f * 24.f)) * mc.r_qviwv) * (q - qcl4[i]);  // orig:430
    const vf LWsurf = -(mc.sig * pow4(Ts));
    const vf LWdown = -(em * mc.sig * pow4(Ta + dTrad));
    const vf LWup = LWdown;

    // ---- sensible heat, f:295
    const vf Qsens = mc.ct_sens * (Ta - Ts);

    // ---- hydro, f:457-467 (abswind incl. gustiness is precomputed on the host, f:452-454)
    vf qs = 3.75e-3f * EXPF(DIVF(17.08085f * (Ts - 273.15f), Ts - 273.15f + 234.175f));
    qs = qs * ez;
    vf Qlat = (q - qs) * absw * mc.cq_latent * mc.rho_air * mc.ce * swet;
    vf dq_eva = -(DIVC(DIVC(Qlat, v_bcast(mc.cq_latent), vd(mc.rc_cq_latent)), v_bcast(mc.r_qviwv), vd(mc.rc_r_qviwv)));
    vf dq_rain = mc.cq_rain * q;
    vf Qlat_air = -(dq_rain * mc.cq_latent * mc.r_qviwv);
    if (sw_ & GREB_SW_NO_HYDRO) {  // orig:452-453
      Qlat = v_bcast(0.0f);
      dq_eva = v_bcast(0.0f);
      dq_rain = v_bcast(0.0f);
      Qlat_air = v_bcast(0.0f);
    }

    // ---- deep_ocean, f:505-523
    const vb warm = ocean && (Ts >= mc.To_ice2);
    // dmld / (z_ocean - mld) and dmld / mld come precomputed with the forcing (GF_RDEEP, GF_RMIX)
    vf dTo = v_sel(warm && (dmld < 0.0f), -(rdeep4[i] * (Ts - To)), v_bcast(0.0f));
    vf dToc = v_sel(warm && (dmld > 0.0f), rmix4[i] * (To - Ts), v_bcast(0.0f));
    dTo = 0.5f * dTo;
    dToc = 0.5f * dToc;
    const vf Tx = v_max(v_bcast(mc.To_ice2), Ts);
    dTo = dTo + DIVF(GREB_DT * mc.co_turb * (Tx - To), mc.cap_ocean * (zoc - mld));
    dToc = dToc + DIVF(GREB_DT * mc.co_turb * (To - Tx), mc.cap_ocean * mld);
    if (sw_ & GREB_SW_NO_DEEP_OCEAN) {  // orig:513-515
      dTo = v_bcast(0.0f);
      dToc = v_bcast(0.0f);
    }

    vf Ts0, To0;
    tendA4[i] = DIVC(GREB_DT * (LWup + LWdown - em * LWsurf + Qlat_air - Qsens), v_bcast(mc.cap_air), vd(mc.rc_cap_air));  // f:260 / f:336
    tq4[i] = GREB_DT * (dq_eva + dq_rain);                                                // f:264 / f:341
    if (!si.spinup) {  // time_loop, f:258-262
      Ts0 = Ts + dToc + DIVF(GREB_DT * (sw + LWsurf - LWdown + Qlat + Qsens + c1[i]), cap);
      To0 = To + dTo + c2[i];
      tfo[i] = c1[i];
      tofo[i] = c2[i];
    } else {  // qflux_correction, f:333-351
      const vf dTs = DIVF(GREB_DT * (sw + LWsurf - LWdown + Qlat + Qsens), cap);
      const vf ts0 = Ts + dTs + dToc;
      const vf to0 = To + dTo;
      const vf T_error = c1[i] - ts0;
      const vf tf = DIVF(T_error * cap, v_bcast(GREB_DT));
      tfo[i] = tf;
      Ts0 = Ts + dTs + dToc + DIVF(tf * GREB_DT, cap);
      const vf tof = c2[i] - to0;
      tofo[i] = tof;
      To0 = To + dTo + tof;
    }

    // ---- seaice(Ts0), f:483-490
    vf capn = cap;
    {
      const vf capo = mc.cap_ocean * mld;
      vf ramp = mc.cap_land + DIVC(capo - mc.cap_land, v_bcast(mc.To_ice2 - mc.To_ice1), vd(mc.rc_alb_ocean)) * (Ts0 - mc.To_ice1);
      ramp = v_sel(Ts0 <= mc.To_ice1, v_bcast(mc.cap_land), ramp);
      ramp = v_sel(Ts0 >= mc.To_ice2, capo, ramp);
      capn = v_sel(ocean, ramp, capn);
      if (sw_ & GREB_SW_NO_ICE_ALBEDO) {  // orig:492-495
        capn = v_sel(v_bit(mask4[i], 3), v_bcast(mc.cap_land), capn);
        capn = v_sel(ocean, capo, capn);
      }
      capn = v_sel(glac, v_bcast(mc.cap_land), capn);
    }
    Ts0o[i] = Ts0;
    To0o[i] = To0;
    capo4[i] = capn;
    albo[i] = albedo;
  }
  v_st4(st + GS_TS * GNC, idx0, Ts0o[0], Ts0o[1], Ts0o[2], Ts0o[3]);
  v_st4(st + GS_TO * GNC, idx0, To0o[0], To0o[1], To0o[2], To0o[3]);
  v_st4(st + GS_CAP * GNC, idx0, capo4[0], capo4[1], capo4[2], capo4[3]);
  v_st4(stash, idx0, tendA4[0], tendA4[1], tendA4[2], tendA4[3]);
  v_st4(stash + GNC, idx0, tq4[0], tq4[1], tq4[2], tq4[3]);
  // diagnostics (f:945) runs in both loops; output (f:974) only in time_loop
  v_st4(acc + GA_TSMN * GNC, idx0, tsmn4[0] + Ts0o[0], tsmn4[1] + Ts0o[1], tsmn4[2] + Ts0o[2], tsmn4[3] + Ts0o[3]);
  if (!si.spinup) {
    v_st4(acc + GA_TMM * GNC, idx0, tmm4[0] + Ts0o[0], tmm4[1] + Ts0o[1], tmm4[2] + Ts0o[2], tmm4[3] + Ts0o[3]);
    v_st4(acc + GA_TOMM * GNC, idx0, tomm4[0] + To0o[0], tomm4[1] + To0o[1], tomm4[2] + To0o[2], tomm4[3] + To0o[3]);
    v_st4(acc + GA_APMM * GNC, idx0, apmm4[0] + albo[0], apmm4[1] + albo[1], apmm4[2] + albo[2], apmm4[3] + albo[3]);
  } else {
    v_st4(corr + GC_TF * GNC, idx0, tfo[0], tfo[1], tfo[2], tfo[3]);
    v_st4(corr + GC_TOF * GNC, idx0, tofo[0], tofo[1], tofo[2], tofo[3]);
  }
#undef DIVF
#undef DIVC
#undef LOGF
#undef EXPF
}

// Phase B: after circulation(Ta).  X = circulated air temperature of 4 cells.
GDEV void column_phase_b(const GrebKernelArgs& a, int member, const StepInfo& si, vi idx0, const vf* X,
                         const float* stash) {
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  vf Ta1[4], tendA[4], tamm[4], o[4];
  v_ld4(Ta1, st + GS_TA * GNC, idx0);
  v_ld4(tendA, stash, idx0);
  if (!si.spinup) v_ld4(tamm, acc + GA_TAMM * GNC, idx0);
  GUNROLL
  for (int i = 0; i < 4; ++i) {
    const vf dTa_crcl = X[i] - Ta1[i];                       // f:551
    o[i] = si.spinup ? (Ta1[i] + tendA[i] + dTa_crcl)        // f:337
                     : (Ta1[i] + dTa_crcl + tendA[i]);       // f:260
  }
  v_st4(st + GS_TA * GNC, idx0, o[0], o[1], o[2], o[3]);
  if (!si.spinup) v_st4(acc + GA_TAMM * GNC, idx0, tamm[0] + o[0], tamm[1] + o[1], tamm[2] + o[2], tamm[3] + o[3]);
}

// Phase C: after circulation(q)
GDEV void column_phase_c(const GrebKernelArgs& a, const GrebMemberConst& mc, int member, const StepInfo& si, vi idx0,
                         const vf* X, const float* stash, const float* corr_s) {
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  float* corr = a.corr + ((size_t)mc.group * GNT + si.ityr) * GC_COUNT * GNC;
  vf q1[4], tq[4], cq[4], qmm[4], o[4], qfo[4];
  v_ld4(q1, st + GS_Q * GNC, idx0);
  v_ld4(tq, stash + GNC, idx0);
  if (!si.spinup) {
    v_ld4(cq, corr_s + GC_QF * GNC, idx0);
    v_ld4(qmm, acc + GA_QMM * GNC, idx0);
  } else {
    v_ldg4(cq, a.qclim + (size_t)si.ityr * GNC, idx0);
  }
  GUNROLL
  for (int i = 0; i < 4; ++i) {
    const vf dq_crcl = X[i] - q1[i];  // f:551
    if (!si.spinup) {                 // f:264-266
      vf dq = tq[i] + dq_crcl + cq[i];
      dq = v_sel(dq <= -q1[i], -0.9f * q1[i], dq);
      o[i] = q1[i] + dq;
      qfo[i] = cq[i];
    } else {  // f:342, 353-355
      const vf qq0 = q1[i] + tq[i] + dq_crcl;
      const vf qf = cq[i] - qq0;
      qfo[i] = qf;
      o[i] = q1[i] + tq[i] + dq_crcl + qf;
    }
  }
  v_st4(st + GS_Q * GNC, idx0, o[0], o[1], o[2], o[3]);
  if (!si.spinup) v_st4(acc + GA_QMM * GNC, idx0, qmm[0] + o[0], qmm[1] + o[1], qmm[2] + o[2], qmm[3] + o[3]);
  else v_st4(corr + GC_QF * GNC, idx0, qfo[0], qfo[1], qfo[2], qfo[3]);
}

// month end (f:977-983): write the five means (16-byte coalesced stores), zero the accumulators
GDEV void column_month_end(const GrebKernelArgs& a, int member, const StepInfo& si, vi idx0) {
  float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
  float* out = a.out ? a.out + ((size_t)member * a.out_months + si.out_rec) * 5 * GNC : nullptr;
  GUNROLL
  for (int f = 0; f < 5; ++f) {
    vf x[4];
    v_ld4(x, acc + f * GNC, idx0);
    if (out) v_st4(out + f * GNC, idx0, x[0] / si.ndm, x[1] / si.ndm, x[2] / si.ndm, x[3] / si.ndm);
    v_st4(acc + f * GNC, idx0, v_bcast(0.0f), v_bcast(0.0f), v_bcast(0.0f), v_bcast(0.0f));
  }
}

GDEV float step_co2(const GrebKernelArgs& a, const GrebMemberConst& mc, int member, int it) {
  return a.spinup ? mc.co2_flux : a.co2[(size_t)member * a.co2_stride + (it - 1) / GNT];  // f:924
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
  // month ends in [it0, it): the slot of this launch's output buffer to write
  int rec = 0;
  {
    const int first = a.it0;
    const int y0 = (first - 1) / GNT, y1 = (it - 1) / GNT;
    for (int y = y0; y <= y1; ++y)
      for (int m = 0; m < 12; ++m) {
        const int e = y * GNT + 2 * cum[m];
        if (e >= first && e < it) ++rec;
      }
  }
  si.out_rec = rec;
  si.co2 = 0.0f;  // filled by the caller: loaded once per simulated year (step_co2)
  return si;
}

// =============================================================================================
// The whole member integration: `nsteps` steps starting at step counter it0.
// `smem` layout: greb_types.h GSM_*.  The SplitBar and the flags must have been initialised
// (sb_init with the number of arriving units, flags = 0) before the first call.
// =============================================================================================
template <int MODE = 0, int SW = 0>
GDEV void member_run_main(const SimtCtx& ctx, const GrebKernelArgs& a, const GrebMemberConst& mc, int member,
                          SyncState& ss) {
  float* smem = ctx.smem;
  float* stash = smem + GSM_STASH;
  float* st = a.state + (size_t)member * GS_COUNT * GNC;
  const float* wzg = a.wz + (size_t)mc.group * 2 * GNC;
  const RowGeom g = row_geom(ctx, mc);
  Tile t;
  float co2 = 0.0f;
  // Flux corrections (time_loop only): 55 KB per step and physics group, read once — the one per-member
  // HBM stream of the scenario (greb_types.h GREB_TMA_CORR).
  const bool tma_issuer = ctx.warp == 0 && lane0(ctx);
  const float* corr_g = a.corr + (size_t)mc.group * GNT * GC_COUNT * GNC;
#if GREB_TMA_CORR
  // The TMA engine copies them into shared memory while the circulations run: after the barrier that
  // follows phase A of step `it`, one thread issues qF(it) (needed by phase C of this step, two circulations
  // later) and TF, ToF(it+1) (needed by phase A of the next step); the consumers wait on the transaction
  // barriers' phase parity.
  float* corr_s = smem + GSM_CORR;
  unsigned long long* bar_a = reinterpret_cast<unsigned long long*>(smem + GSM_TMA_BAR_A);
  unsigned long long* bar_q = reinterpret_cast<unsigned long long*>(smem + GSM_TMA_BAR_Q);
  int par_a = 0, par_q = 0;
  if (!a.spinup && tma_issuer)
    tma_load(corr_s, corr_g + (size_t)((a.it0 - 1) % GNT) * GC_COUNT * GNC, 2 * GNC * (unsigned)sizeof(float), bar_a);
#endif
  cta_sync(ctx);

  GNOUNROLL
  for (int it = a.it0; it < a.it0 + a.nsteps; ++it) {
    StepInfo si = step_info(a, mc, member, it);
    if (it == a.it0 || (it - 1) % GNT == 0) co2 = step_co2(a, mc, member, it);
    si.co2 = co2;
    const float* forc = a.forc + (size_t)si.ityr * GF_COUNT * GNC;

#if defined(GREB_DBG_CLOCKS) && GREB_DEVICE
    // timing experiment: cycles per 12-h step in the phases of the step (block 0, thread 0)
    long long sc0 = clock64();
    static __shared__ long long dbg_acc[8];
    if (it == a.it0 && threadIdx.x == 0) for (int i = 0; i < 8; ++i) dbg_acc[i] = 0;
#define SCLK(i) { const long long t1_ = clock64(); if (threadIdx.x == 0) dbg_acc[i] += t1_ - sc0; sc0 = t1_; }
#else
#define SCLK(i)
#endif
    // ---- phase A: column physics, Ts/To/cap update
#if GREB_TMA_CORR
    if (!a.spinup) {
      tma_wait(bar_a, par_a);
      par_a ^= 1;
    }
#else
    const float* corr_s = corr_g + (size_t)si.ityr * GC_COUNT * GNC;   // ordinary loads of the (L2-prefetched) slice
#endif
    GNOUNROLL
    for (int q = 0; q < 3; ++q) column_phase_a<MODE, SW>(a, mc, member, si, g.k, g.k * GX + g.col + 4 * q, stash, corr_s);
    SCLK(0)
    {
      const FastRow fr = fast_row(g.k, mc);
      tile_load_uv(t, g, forc + GF_U * GNC, forc + GF_V * GNC, smem, MODE == 1 ? -fr.cadv : 1.0f,
                   MODE == 1 ? fr.cyA : 1.0f, MODE == 1 ? fr.cyB : 1.0f);
    }
    // the helper warps read the rows they circulate from global state written by the main warps
    cta_sync(ctx);
#if GREB_TMA_CORR
    if (!a.spinup && tma_issuer) {   // every thread is past its reads of TF, ToF (this step) and qF (previous step)
      const float* cg = corr_g + (size_t)si.ityr * GC_COUNT * GNC;
      tma_load(corr_s + GC_QF * GNC, cg + GC_QF * GNC, GNC * (unsigned)sizeof(float), bar_q);
      if (it + 1 < a.it0 + a.nsteps)
        tma_load(corr_s, corr_g + (size_t)(it % GNT) * GC_COUNT * GNC, 2 * GNC * (unsigned)sizeof(float), bar_a);
    }
#else
    // the NEXT step's 55 KB slice: HBM -> L2 by the bulk-copy engine while the two circulations run
    if (!a.spinup && tma_issuer)
      tma_prefetch_l2(corr_g + (size_t)(it % GNT) * GC_COUNT * GNC, GC_COUNT * GNC * (unsigned)sizeof(float));
#endif
    SCLK(1)

    // ---- circulation of air temperature (f:301), then of humidity (f:303): one code instance
    GNOUNROLL
    for (int fld = 0; fld < 2; ++fld) {
      // orig:560-564: with zero winds every advection term is +-0, so X + dx_diffuse + dx_advec == X + dx_diffuse
      if (SW && fld == 1 && (mc.switches & GREB_SW_VAPOR_DIFFUSION_ONLY))
        tile_load_uv(t, g, forc + GF_U * GNC, forc + GF_V * GNC, smem, 0.0f, 0.0f, 0.0f);
      // orig:553-555: `circulation` returns before assigning dX_crcl; defined here as dX_crcl = 0
      // (include/greb_b200.h GREB_SW_NO_*_CIRCULATION): X = X_in, no sub-steps, no barriers (the helper
      // warps skip the same ones)
      const bool crcl_off = SW && (mc.switches & (fld == 0 ? GREB_SW_NO_HEAT_CIRCULATION : GREB_SW_NO_VAPOR_CIRCULATION));
      if (!crcl_off) {
        tile_load_wz(t, g, wzg + fld * GNC, smem);
#if GREB_YCOEF
        if (MODE == 1) tile_fold_ycoef(t, g, mc.ccy_diff, smem);
#endif
      }
      tile_load_field(t, g, st + (fld == 0 ? GS_TA : GS_Q) * GNC);
      SCLK(2)
      if (!crcl_off) circulation_main<MODE>(ctx, t, g, mc, ss);
      SCLK(3)
#if GREB_TMA_CORR
      if (fld == 1 && !a.spinup) {
        tma_wait(bar_q, par_q);
        par_q ^= 1;
      }
#endif
      GUNROLL
      for (int q = 0; q < 3; ++q) {
        if (fld == 0) column_phase_b(a, member, si, g.k * GX + g.col + 4 * q, &t.T[4 * q], stash);
        else column_phase_c(a, mc, member, si, g.k * GX + g.col + 4 * q, &t.T[4 * q], stash, corr_s);
      }
      SCLK(4)
    }
#if defined(GREB_DBG_CLOCKS) && GREB_DEVICE
    if (it == a.it0 + a.nsteps - 1 && threadIdx.x == 0 && blockIdx.x == 0 && a.nsteps > 1)
      printf("per step: phaseA %lld, load_uv+sync %lld, load_wz/field %lld, circulations %lld, phaseB/C %lld cycles\n",
             dbg_acc[0] / a.nsteps, dbg_acc[1] / a.nsteps, dbg_acc[2] / a.nsteps, dbg_acc[3] / a.nsteps,
             dbg_acc[4] / a.nsteps);
#endif

    // ---- output (f:975-985)
    if (!si.spinup && si.month_end) {
      GNOUNROLL
      for (int q = 0; q < 3; ++q) column_month_end(a, member, si, g.k * GX + g.col + 4 * q);
    }

    // ---- annual mean diagnostics (f:948-956)
    if (si.ityr == GNT - 1) {
      float* acc = a.acc + (size_t)member * GA_COUNT * GNC;
      float* scratch = ss.hb;
      cta_sync(ctx);  // every thread is past its last read of the field buffers
      GUNROLL
      for (int q = 0; q < 3; ++q) {
        vf x[4];
        const vi idx0 = g.k * GX + g.col + 4 * q;
        v_ld4(x, acc + GA_TSMN * GNC, idx0);
        v_st4(scratch, idx0, x[0] / (float)GNT, x[1] / (float)GNT, x[2] / (float)GNT, x[3] / (float)GNT);  // f:949
        v_st4(acc + GA_TSMN * GNC, idx0, v_bcast(0.0f), v_bcast(0.0f), v_bcast(0.0f), v_bcast(0.0f));     // f:955
      }
      cta_sync(ctx);
      if (ctx.warp == 0 && lane0(ctx)) {
        float s = 0.0f, sw = 0.0f;
        for (int k = 0; k < GY; ++k) {
          float rs = 0.0f;
          for (int i = 0; i < GX; ++i) {
            s = s + scratch[k * GX + i];  // f:954 sum() in array element order
            rs = rs + scratch[k * GX + i];
          }
          sw = sw + a.coslat_w[k] * (rs / (float)GX);
        }
        const float g0 = s / (float)(GX * GY) - 273.15f;
        a.diag[member * 2 + 0] = g0;
        a.diag[member * 2 + 1] = sw - 273.15f;
        if (!(g0 == g0) || g0 > 1e4f || g0 < -1e4f) a.flags[member] = 1;
      }
      cta_sync(ctx);  // scratch (= field buffer) is free again
    }
  }
}

// the helper warps' view of the same step sequence (identical barrier pattern)
template <int MODE = 0, int SW = 0>
GDEV void member_run_helper(const SimtCtx& ctx, const GrebKernelArgs& a, const GrebMemberConst& mc, int member,
                            SyncState& ss) {
  const float* st = a.state + (size_t)member * GS_COUNT * GNC;
  const float* wzg = a.wz + (size_t)mc.group * 2 * GNC;
  const HelperGeom hg = helper_geom(ctx, mc);
  HelperRow hr[GREB_HROWS];
  cta_sync(ctx);   // pairs with the barrier after the first TMA issue in member_run_main
  GNOUNROLL
  for (int it = a.it0; it < a.it0 + a.nsteps; ++it) {
    const int ityr = (it - 1) % GNT;
    const float* forc = a.forc + (size_t)ityr * GF_COUNT * GNC;
    helper_load_uv(hr, hg, forc + GF_U * GNC, forc + GF_V * GNC);
    cta_sync(ctx);
    GNOUNROLL
    for (int fld = 0; fld < 2; ++fld) {
      if (SW && fld == 1 && (mc.switches & GREB_SW_VAPOR_DIFFUSION_ONLY)) {
        GUNROLL
        for (int i = 0; i < GREB_HROWS; ++i)
          GUNROLL
          for (int c = 0; c < 3; ++c) {
            hr[i].U[c] = v_bcast(0.0f);
            hr[i].V[c] = v_bcast(0.0f);
          }
      }
      if (SW && (mc.switches & (fld == 0 ? GREB_SW_NO_HEAT_CIRCULATION : GREB_SW_NO_VAPOR_CIRCULATION))) continue;
      helper_load_wz(hr, hg, wzg + fld * GNC);
      circulation_helper<MODE>(ctx, hr, hg, mc, st + (fld == 0 ? GS_TA : GS_Q) * GNC, ss);
    }
    if (ityr == GNT - 1) {
      cta_sync(ctx);
      cta_sync(ctx);
      cta_sync(ctx);
    }
  }
}

template <int MODE = 0, int SW = 0>
GDEV void member_run(const SimtCtx& ctx, const GrebKernelArgs& a, const GrebMemberConst& mc, int member) {
  SyncState ss;
  ss.bar = reinterpret_cast<SplitBar*>(ctx.smem + GSM_SYNC);
  ss.hb = ctx.smem + GSM_HB;
  ss.smem = ctx.smem;
  ss.phase = 0;
  if (ctx_is_helper(ctx)) member_run_helper<MODE, SW>(ctx, a, mc, member, ss);
  else member_run_main<MODE, SW>(ctx, a, mc, member, ss);
}
