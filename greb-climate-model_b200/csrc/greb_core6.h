// greb_core6.h — the circulation on 6-cell tiles: 24 warps per member instead of 12 + 2.
//
// The round-2 measurements say the circulation is bound by the latency of each warp's dependent chain at
// 3.5 warps per scheduler, not by issue slots (packing 20 % of the instructions away changed nothing,
// DESIGN.md section 5).  This variant halves the tile — a 16-lane group owns one latitude row, each lane 6
// consecutive longitudes, a warp owns the two rows k and 47-k (same |latitude|: same branch of f:592, same
// number of polar sub-sub-steps, so a warp never diverges) — and so doubles the resident warps at <= 80
// registers per thread.  There are no helper warps: the pole rows' sub-sub-steps (f:655-718) run as a loop
// inside warp 0, halos by shuffle.  Arithmetic: the exact mode of greb_core.h, operation for operation.
//
// RESULT (B200, tools/circ_bench.py, 1,184 fields x 24 sub-steps, exact mode): bit-identical to the oracle and
// to the 12-cell kernel, 80 registers, no spills — and SLOWER: 0.566 ms against 0.402 ms, and still 0.427 ms
// with the pole rows' extra sub-sub-steps compiled out (-DT6_DBG_TIME2_1).  Twice the warps issue ~8 % more
// instructions (halo redundancy: 11 edge differences per 6 cells instead of 17 per 12) at the same ~61 % of
// the issue slots, so occupancy is not what holds the circulation back: a sub-step is one warp's x part
// (shuffle -> stencil chain) plus its y part (24 LDS.128 behind the barrier, shared-memory-bound when every
// warp does it at once) in sequence, and neither gets shorter with narrower tiles.  Kept as an experiment
// behind GREB_B200_TILE6 in greb_b200_circulation (kernel-level entry only); the product kernels are
// greb_core.h's.
#pragma once

#include "greb_core.h"

#define T6_CPT 6
#define T6_LPR 16
#define T6_NWARP 24
#define T6_NTHREADS (T6_NWARP * 32)
// shared memory, floats: [2][GNC] field copies | sync slot | private y constants [4][3 chunks][768 threads][2]
#define T6_HB 0
#define T6_SYNC (2 * GNC)
#define T6_PRIV (T6_SYNC + 32)
#define T6_FLOATS (T6_PRIV + PRIV_COUNT * 3 * T6_NTHREADS * 2)

#if GREB_DEVICE
struct Row6 {
  int k, polar, ykind, time2;
  int km2, km1, kp1, kp2;
  int col, lane_l, lane_r, tid2;
  bool is_bug;
};

GDEV Row6 row6(int warp, int lane, const GrebMemberConst& mc) {
  Row6 g;
  const int seg = lane & 15, base = lane & 16;
  g.k = (lane & 16) ? (GY - 1 - warp) : warp;       // rows k and 47-k share their geometry
  g.col = T6_CPT * seg;
  g.lane_l = base | ((seg + 15) & 15);
  g.lane_r = base | ((seg + 1) & 15);
  g.is_bug = seg == 15;                              // owns longitude 93 as its cell 3 (f:881)
  g.tid2 = 2 * (warp * 32 + lane);
  g.polar = mc.polar[g.k];
#ifdef T6_DBG_TIME2_1   // timing experiment only (wrong results): no extra polar sub-sub-steps
  g.time2 = 1;
#else
  g.time2 = mc.time2_diff[g.k];
#endif
  g.ykind = g.k == 0 ? 1 : g.k == 1 ? 2 : g.k == GY - 2 ? 3 : g.k == GY - 1 ? 4 : 0;
  g.km2 = g.k >= 2 ? g.k - 2 : 0;
  g.km1 = g.k >= 1 ? g.k - 1 : 0;
  g.kp1 = g.k <= GY - 2 ? g.k + 1 : GY - 1;
  g.kp2 = g.k <= GY - 3 ? g.k + 2 : GY - 1;
  return g;
}

struct Tile6 {
  float T[T6_CPT], W[T6_CPT], U[T6_CPT], wxl[3], wxr[3];
};

GDEV float* priv6(float* smem, int which, int q) { return smem + T6_PRIV + (which * 3 + q) * (T6_NTHREADS * 2); }
GDEV float2 ld2(const float* p, int idx) { return *reinterpret_cast<const float2*>(p + idx); }
GDEV float2 ldg2(const float* p, int idx) { return __ldg(reinterpret_cast<const float2*>(p + idx)); }
GDEV void st2(float* p, int idx, float a, float b) { *reinterpret_cast<float2*>(p + idx) = make_float2(a, b); }

GDEV void t6_load_uv(Tile6& t, const Row6& g, const float* u, const float* v, float* smem) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float2 a = ldg2(u, g.k * GX + g.col + 2 * q), b = ldg2(v, g.k * GX + g.col + 2 * q);
    t.U[2 * q] = a.x;
    t.U[2 * q + 1] = a.y;
    st2(priv6(smem, PRIV_V, q), g.tid2, b.x, b.y);
  }
}

GDEV void t6_load_wz(Tile6& t, const Row6& g, const float* wz, float* smem) {
  const bool has_m1 = g.k >= 1, has_m2 = g.k >= 2, has_p1 = g.k <= GY - 2, has_p2 = g.k <= GY - 3;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int c = g.col + 2 * q;
    const float2 w0 = ldg2(wz, g.k * GX + c);
    float2 wm1 = ldg2(wz, g.km1 * GX + c), wp1 = ldg2(wz, g.kp1 * GX + c);
    const float2 wm2 = ldg2(wz, g.km2 * GX + c), wp2 = ldg2(wz, g.kp2 * GX + c);
    const float2 v = ld2(priv6(smem, PRIV_V, q), g.tid2);
    t.W[2 * q] = w0.x;
    t.W[2 * q + 1] = w0.y;
    if (!has_m1) wm1 = make_float2(0.f, 0.f);
    if (!has_p1) wp1 = make_float2(0.f, 0.f);
    const float ax = has_m2 ? wm2.x : 0.f, ay = has_m2 ? wm2.y : 0.f;
    const float bx = has_p2 ? wp2.x : 0.f, by = has_p2 ? wp2.y : 0.f;
    st2(priv6(smem, PRIV_WM1, q), g.tid2, wm1.x, wm1.y);
    st2(priv6(smem, PRIV_WP1, q), g.tid2, wp1.x, wp1.y);
    st2(priv6(smem, PRIV_WFY, q), g.tid2, v.x >= 0.f ? ax : bx, v.y >= 0.f ? ay : by);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int cl = g.col == 0 ? GX - 3 + i : g.col - 3 + i;
    const int cr = g.col == GX - T6_CPT ? i : g.col + T6_CPT + i;
    t.wxl[i] = __ldg(wz + g.k * GX + cl);
    t.wxr[i] = __ldg(wz + g.k * GX + cr);
  }
}

GDEV void t6_load_field(Tile6& t, const Row6& g, const float* X) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float2 a = ld2(X, g.k * GX + g.col + 2 * q);
    t.T[2 * q] = a.x;
    t.T[2 * q + 1] = a.y;
  }
}

GDEV void t6_publish(const Tile6& t, const Row6& g, float* buf) {
#pragma unroll
  for (int q = 0; q < 3; ++q) st2(buf, g.k * GX + g.col + 2 * q, t.T[2 * q], t.T[2 * q + 1]);
}

// the x-diffusion bracket of f:620-625 for the 6 own cells of a row segment whose values (own + 3 halo cells
// on each side) are TT[0..11]; also returns the edge differences and weighted products the advection reuses
struct XProd6 {
  float d[11], P[8], Q[11];
};
GDEV void t6_bracket(float (&S)[T6_CPT], XProd6& x, const float (&TT)[12], const float (&WW)[12]) {
  float A[8], B[10];
#pragma unroll
  for (int e = 0; e < 11; ++e) x.d[e] = TT[e + 1] - TT[e];
#pragma unroll
  for (int e = 0; e <= 7; ++e) x.P[e] = WW[e] * x.d[e];
#pragma unroll
  for (int e = 3; e <= 10; ++e) x.Q[e] = WW[e + 1] * x.d[e];
#pragma unroll
  for (int e = 1; e <= 7; ++e) A[e] = x.P[e] - x.P[e - 1];
#pragma unroll
  for (int e = 3; e <= 9; ++e) B[e] = x.Q[e + 1] - x.Q[e];
#pragma unroll
  for (int j = 0; j < T6_CPT; ++j) {
    const int e = j + 3;
    const float G = x.Q[e] - x.P[e - 1];
    S[j] = __fmaf_rn(4.0f, B[e], __fmaf_rn(4.0f, A[e - 1], 10.0f * G)) + A[e - 2] + B[e + 1];
  }
}

GDEV void t6_substep_x(float (&dTx)[T6_CPT], float (&aTx)[T6_CPT], const Tile6& t, const Row6& g, const GrebMemberConst& mc) {
  float TT[12], WW[12], S[T6_CPT];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    TT[i] = __shfl_sync(0xffffffffu, t.T[3 + i], g.lane_l);
    TT[9 + i] = __shfl_sync(0xffffffffu, t.T[i], g.lane_r);
    WW[i] = t.wxl[i];
    WW[9 + i] = t.wxr[i];
  }
#pragma unroll
  for (int i = 0; i < T6_CPT; ++i) {
    TT[3 + i] = t.T[i];
    WW[3 + i] = t.W[i];
  }
  XProd6 x;
  t6_bracket(S, x, TT, WW);
  if (!g.polar) {
    const float cc = mc.ccx_diff[g.k], cca = mc.ccx_adv[g.k];
#pragma unroll
    for (int j = 0; j < T6_CPT; ++j) {
      const int e = j + 3;
      dTx[j] = div20(cc * S[j]);
      const bool pu = t.U[j] >= 0.0f;
      const float near = pu ? x.P[e - 1] : -x.Q[e];
      const float far = (pu ? WW[e - 2] : WW[e + 2]) * (TT[e] - (pu ? TT[e - 2] : TT[e + 2]));
      const float Xu = (-fabsf(t.U[j])) * (near + far);
      aTx[j] = div3(cca * Xu);
    }
  } else {
    const float cc2 = mc.ccx2_diff[g.k], cca2 = mc.ccx2_adv[g.k];
    float h[T6_CPT];
#pragma unroll
    for (int j = 0; j < T6_CPT; ++j) {
      const int e = j + 3;
      float dd = div20(cc2 * S[j]);
      dd = polar_clamp(dd, t.T[j]);      // f:715
      h[j] = t.T[j] + dd;                // f:716
      const bool pu = t.U[j] >= 0.0f;
      const float near10 = (10.0f * (pu ? WW[e - 1] : WW[e + 1])) * (pu ? x.d[e - 1] : x.d[e]);
      float mid = pu ? x.P[e - 2] : x.Q[e + 1];
      float far = pu ? x.P[e - 3] : x.Q[e + 2];
      if (j == 3) {                      // longitude 93: jp1 = jp2 = xdim-1, jp3 = 1 (f:881)
        const bool bug = g.is_bug && !pu;
        mid = bug ? 0.0f : mid;
        far = bug ? WW[e + 3] * (TT[e + 3] - TT[e + 1]) : far;
      }
      const float Sa = __fmaf_rn(4.0f, mid, near10) + far;
      const float Xu = (-t.U[j]) * Sa;
      float da = div20(cca2 * Xu);
      da = polar_clamp(da, t.T[j]);      // f:907
      const float ha = t.T[j] + da;      // f:908
      aTx[j] = ha - t.T[j];              // f:910
    }
    // the remaining polar sub-sub-steps (f:655-718): time2 is the same for every lane of the warp
#pragma unroll 1
    for (int tt2 = 1; tt2 < g.time2; ++tt2) {
      float HH[12], S2[T6_CPT];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        HH[i] = __shfl_sync(0xffffffffu, h[3 + i], g.lane_l);
        HH[9 + i] = __shfl_sync(0xffffffffu, h[i], g.lane_r);
      }
#pragma unroll
      for (int i = 0; i < T6_CPT; ++i) HH[3 + i] = h[i];
      XProd6 y;
      t6_bracket(S2, y, HH, WW);
#pragma unroll
      for (int j = 0; j < T6_CPT; ++j) {
        float dd = div20(cc2 * S2[j]);
        dd = polar_clamp(dd, h[j]);
        h[j] = h[j] + dd;
      }
    }
#pragma unroll
    for (int j = 0; j < T6_CPT; ++j) dTx[j] = h[j] - t.T[j];   // f:718
  }
}

// y part + update for every row kind (f:587-590, f:756-795); mirrors substep_y / helper_y of greb_core.h
GDEV void t6_substep_y(Tile6& t, const float (&dTx)[T6_CPT], const float (&aTx)[T6_CPT], const Row6& g,
                       const GrebMemberConst& mc, const float* buf, float* smem) {
  const float ccyd = mc.ccy_diff, ccya = mc.ccy_adv;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int c = g.col + 2 * q;
    const float2 V = ld2(priv6(smem, PRIV_V, q), g.tid2), Wm1 = ld2(priv6(smem, PRIV_WM1, q), g.tid2);
    const float2 Wp1 = ld2(priv6(smem, PRIV_WP1, q), g.tid2), WFY = ld2(priv6(smem, PRIV_WFY, q), g.tid2);
    const float2 tm1 = ld2(buf, g.km1 * GX + c), tp1 = ld2(buf, g.kp1 * GX + c);
    const float2 tm2 = ld2(buf, g.km2 * GX + c), tp2 = ld2(buf, g.kp2 * GX + c);
    const float Vv[2] = {V.x, V.y}, wm1[2] = {Wm1.x, Wm1.y}, wp1[2] = {Wp1.x, Wp1.y}, wfy[2] = {WFY.x, WFY.y};
    const float am1[2] = {tm1.x, tm1.y}, ap1[2] = {tp1.x, tp1.y}, am2[2] = {tm2.x, tm2.y}, ap2[2] = {tp2.x, tp2.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int j = 2 * q + i;
      const float T = t.T[j];
      const float Pym1 = wm1[i] * (T - am1[i]);
      const float Qy0 = wp1[i] * (ap1[i] - T);
      float dTy = ccyd * (Qy0 - Pym1);
      if (g.ykind == 1) dTy = (ccyd * wp1[i]) * (ap1[i] - T);   // f:589
      if (g.ykind == 4) dTy = (ccyd * wm1[i]) * (am1[i] - T);   // f:590
      const bool pv = Vv[i] >= 0.0f;
      const float near = pv ? Pym1 : -Qy0;
      const float far = wfy[i] * (T - (pv ? am2[i] : ap2[i]));
      const float Xv = (-fabsf(Vv[i])) * (near + far);
      float aTy = div3(ccya * Xv);
      if (g.ykind == 2) aTy = pv ? ccya * Xv : ccya * div3(Xv);  // f:766-769
      if (g.ykind == 3) aTy = pv ? ccya * div3(Xv) : ccya * Xv;  // f:784-787
      const float dXd = t.W[j] * (dTx[j] + dTy);   // f:721
      const float dXa = aTx[j] + aTy;              // f:913
      t.T[j] = (T + dXd) + dXa;                    // f:549
    }
  }
}

// circulation (f:528-553): 24 sub-steps, ONE CTA barrier each (publish -> x part of the next -> barrier -> y part)
GDEV void t6_circulation(Tile6& t, const Row6& g, const GrebMemberConst& mc, float* hb, float* smem, int& phase, bool late) {
  t6_publish(t, g, hb + (phase & 1) * GNC);
#pragma unroll 1
  for (int tt = 0; tt < GSUB; ++tt) {
    float dTx[T6_CPT], aTx[T6_CPT];
    if (late) __syncthreads();
    t6_substep_x(dTx, aTx, t, g, mc);
    if (!late) __syncthreads();
    const float* buf = hb + (phase & 1) * GNC;
    t6_substep_y(t, dTx, aTx, g, mc, buf, smem);
    phase++;
    t6_publish(t, g, hb + (phase & 1) * GNC);
  }
  __syncthreads();
  phase++;
}
#endif  // GREB_DEVICE
