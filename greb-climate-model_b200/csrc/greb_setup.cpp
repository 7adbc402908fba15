// greb_setup.cpp — see greb_setup.h.  Reference: /root/reference/src/greb.f90 ("f:NNN").
#include "greb_setup.h"

#include <math.h>
#include <string.h>

#include <algorithm>

void greb_b200_physics_defaults(greb_physics_par* p) {  // f:68-104
  static const float pe[10] = {9.0721f, 106.7252f, 61.5562f, 0.0179f, 0.0028f,
                               0.0570f, 0.3462f, 2.3406f, 0.7032f, 1.0662f};
  p->pi = 3.1416f;
  p->sig = 5.6704e-8f;
  p->rho_ocean = 999.1f;
  p->rho_land = 2600.f;
  p->rho_air = 1.2f;
  p->cp_ocean = 4186.f;
  p->cp_land = 926.222f;
  p->cp_air = 1005.f;
  p->eps = 1.f;
  p->d_ocean = 50.f;
  p->d_land = 2.f;
  p->d_air = 5000.f;
  p->ct_sens = 22.5f;
  p->da_ice = 0.25f;
  p->a_no_ice = 0.1f;
  p->a_cloud = 0.35f;
  p->Tl_ice1 = 273.15f - 10.f;
  p->Tl_ice2 = 273.15f;
  p->To_ice1 = 273.15f - 7.f;
  p->To_ice2 = 273.15f - 1.7f;
  p->co_turb = 5.0f;
  p->kappa = 8e5f;
  p->ce = 2e-3f;
  p->cq_latent = 2.257e6f;
  p->cq_rain = -0.1f / 24.f / 3600.f;
  p->z_air = 8400.f;
  p->z_vapor = 5000.f;
  p->r_qviwv = 2.6736e3f;
  memcpy(p->p_emi, pe, sizeof pe);
  p->co2_flux = 298.f;
}

void greb_b200_physics_original(greb_physics_par* p) {  // src/greb.original.model.f90:63-101, :178
  greb_b200_physics_defaults(p);
  p->cp_land = p->cp_ocean / 4.5f;
  p->co2_flux = 340.f;
}

void greb_b200_pad_co2(const float* given, int n_given, float* co2, int n_years) {  // f:1047-1061
  for (int i = 0; i < n_years; ++i) co2[i] = (i < n_given) ? given[i] : -1.f;
  if (n_years > 0 && co2[0] == -1.f) co2[0] = 680.f;
  for (int i = 1; i < n_years; ++i)
    if (co2[i] < 0.f) {
      for (int j = i; j < n_years; ++j) co2[j] = co2[i - 1];
      break;
    }
}

bool greb_physics_equal(const greb_physics_par& a, const greb_physics_par& b) {
  return memcmp(&a, &b, sizeof(greb_physics_par)) == 0;
}

void greb_build_forcing(GrebHostForcing& F, const float* z_topo, const float* glacier, const float* sw_solar,
                        const float* tclim, const float* qclim, const float* swetclim, const float* uclim,
                        const float* vclim, const float* mldclim, const float* cldclim) {
  F.forc.assign((size_t)GNT * GF_COUNT * GNC, 0.f);
  F.sw_solar.assign(sw_solar, sw_solar + (size_t)GNT * GY);
  F.z_topo.assign(z_topo, z_topo + GNC);
  F.tclim.assign(tclim, tclim + (size_t)GNT * GNC);
  F.qclim.assign(qclim, qclim + (size_t)GNT * GNC);
  F.mld0.assign(mldclim, mldclim + GNC);
  F.mask.assign(GNC, 0);
  F.z_ocean.assign(GNC, 0.f);
  F.toclim.assign(GNC, 0.f);
  for (int c = 0; c < GNC; ++c) {
    int m = 0;
    if (z_topo[c] >= 0.f) m |= GM_TOPO_GE0;  // f:384
    if (z_topo[c] < 0.f) m |= GM_TOPO_LT0;   // f:389, 483, 511
    if (glacier[c] > 0.5f) m |= GM_GLACIER;  // f:395, 490
    if (z_topo[c] > 0.f) m |= GM_TOPO_GT0;   // greb.original.model.f90:493 (log_exp <= 5)
    F.mask[c] = m;
    // f:1087-1094 Toclim
    float mn = tclim[c];
    for (int n = 1; n < GNT; ++n) mn = std::min(mn, tclim[(size_t)n * GNC + c]);
    if (mn - 273.15f < -1.7f) mn = -1.7f + 273.15f;
    F.toclim[c] = mn;
    // f:179-183 z_ocean
    float zo = 0.f;
    for (int n = 0; n < GNT; ++n)
      if (mldclim[(size_t)n * GNC + c] > zo) zo = mldclim[(size_t)n * GNC + c];
    F.z_ocean[c] = 3.0f * zo;
  }
  for (int n = 0; n < GNT; ++n) {
    float* rec = &F.forc[(size_t)n * GF_COUNT * GNC];
    const int np = (n > 0) ? n - 1 : GNT - 1;  // f:507-508
    for (int c = 0; c < GNC; ++c) {
      const size_t i = (size_t)n * GNC + c;
      const float u = uclim[i], v = vclim[i];
      rec[GF_U * GNC + c] = u;
      rec[GF_V * GNC + c] = v;
      rec[GF_CLD * GNC + c] = cldclim[i];
      rec[GF_DTRAD * GNC + c] = -0.16f * tclim[i] - 5.f;  // f:176
      rec[GF_SWET * GNC + c] = swetclim[i];
      float aw = sqrtf(u * u + v * v);                                // f:452
      if (z_topo[c] > 0.f) aw = sqrtf(aw * aw + 2.0f * 2.0f);          // f:453
      if (z_topo[c] < 0.f) aw = sqrtf(aw * aw + 3.0f * 3.0f);          // f:454
      rec[GF_ABSWIND * GNC + c] = aw;
      rec[GF_MLD * GNC + c] = mldclim[i];
      const float dmld = mldclim[i] - mldclim[(size_t)np * GNC + c];
      rec[GF_DMLD * GNC + c] = dmld;
      rec[GF_RDEEP * GNC + c] = dmld / (F.z_ocean[c] - mldclim[i]);   // f:512 (may be inf/nan where it is never selected)
      rec[GF_RMIX * GNC + c] = dmld / mldclim[i];                      // f:513
    }
  }
  // cos-lat weights of the README's global mean (README.md:36-37), normalised
  F.coslat_w.resize(GY);
  double s = 0;
  for (int k = 0; k < GY; ++k) {
    const double lat = (k + 0.5) * 3.75 - 90.0;
    F.coslat_w[k] = (float)cos(lat * 3.14159265358979323846 / 180.0);
    s += F.coslat_w[k];
  }
  for (int k = 0; k < GY; ++k) F.coslat_w[k] = (float)(F.coslat_w[k] / s);
}

static int f_nint(float x) { return (int)lroundf(x); }

int greb_assign_rows(const int* polar, const int* time2_diff, const int* time2_adv, int* row_of_group,
                     int* hslot_of_row, int* helper_row, int* n_hslots, const int* warp_order) {
  // Lane groups (4 per main warp) -> rows.  Warps are kept homogeneous (all polar-branch rows or all
  // main-branch rows) so that the f:592/f:799 branch does not diverge inside a warp; the cheaper
  // main-row warps go to the SM sub-partitions (warp % 4) that also host a helper warp (12, 13) in the
  // 14-warp layout.
  int pol[GY], mainr[GY], np = 0, nm = 0;
  for (int k = 0; k < GY; ++k) (polar[k] ? pol[np++] : mainr[nm++]) = k;
  int rows[GY], n = 0;  // rows in group order of a virtual warp list: polar warps first
  for (int i = 0; i < np; ++i) rows[n++] = pol[i];
  for (int i = 0; i < nm; ++i) rows[n++] = mainr[i];
  const int n_polar_warps = (np + 3) / 4;
  // warp_order: the logical main warps in the order in which they receive quads of rows, polar rows first
  // (a property of the warp placement, greb_b200.cu greb_layouts); default = the 14-warp placement, where
  // polar warps take ids with (w % 4) in {2, 3} first
  int order[GREB_NMAIN], no = 0;
  if (warp_order) {
    for (int w = 0; w < GREB_NMAIN; ++w) order[no++] = warp_order[w];
  } else {
    for (int w = 0; w < GREB_NMAIN; ++w)
      if ((w & 3) >= 2) order[no++] = w;
    for (int w = 0; w < GREB_NMAIN; ++w)
      if ((w & 3) < 2) order[no++] = w;
  }
  (void)n_polar_warps;
  for (int v = 0; v < GREB_NMAIN; ++v)
    for (int s = 0; s < 4; ++s) row_of_group[order[v] * 4 + s] = rows[v * 4 + s];
  // helper-owned rows: the two pole rows (always) and every row whose polar diffusion needs more
  // than one sub-sub-step; the helper code implements the polar branch only
  int nh = 0;
  for (int k = 0; k < GY; ++k) {
    hslot_of_row[k] = -1;
    if (time2_adv[k] > 1) return -2;  // cannot happen on the 96x48 grid (f:838: dd = 1 for every row)
    if (k == 0 || k == GY - 1 || time2_diff[k] > 1) {
      if (!polar[k]) return -3;
      if (nh >= GREB_MAXH) return -1;
      hslot_of_row[k] = nh;
      helper_row[nh++] = k;
    }
  }
  for (int s = nh; s < GREB_MAXH; ++s) helper_row[s] = 0;
  *n_hslots = nh;
  return 0;
}

int greb_build_member_const(GrebMemberConst& mc, const greb_physics_par& p, int group, const int* warp_order) {
  memset(&mc, 0, sizeof mc);
  mc.sig = p.sig; mc.ct_sens = p.ct_sens; mc.da_ice = p.da_ice; mc.a_no_ice = p.a_no_ice; mc.a_cloud = p.a_cloud;
  mc.Tl_ice1 = p.Tl_ice1; mc.Tl_ice2 = p.Tl_ice2; mc.To_ice1 = p.To_ice1; mc.To_ice2 = p.To_ice2;
  mc.co_turb = p.co_turb; mc.ce = p.ce; mc.cq_latent = p.cq_latent; mc.cq_rain = p.cq_rain;
  mc.rho_air = p.rho_air; mc.r_qviwv = p.r_qviwv;
  memcpy(mc.p_emi, p.p_emi, sizeof mc.p_emi);
  mc.cap_ocean = p.cp_ocean * p.rho_ocean;          // f:186
  mc.cap_land = p.cp_land * p.rho_land * p.d_land;  // f:187
  mc.cap_air = p.cp_air * p.rho_air * p.d_air;      // f:188
  mc.co2_flux = p.co2_flux;
  mc.group = group;
  {
    const float dl = p.Tl_ice2 - p.Tl_ice1, dof = p.To_ice2 - p.To_ice1;   // the fp32 differences the reference divides by
    mc.rc_alb_land = 1.0 / (double)dl;
    mc.rc_alb_ocean = 1.0 / (double)dof;
    mc.rc_pe8 = 1.0 / (double)p.p_emi[8];
    mc.rc_cq_latent = 1.0 / (double)p.cq_latent;
    mc.rc_r_qviwv = 1.0 / (double)p.r_qviwv;
    mc.rc_cap_air = 1.0 / (double)mc.cap_air;
  }
  // geometry, f:578-582 / f:749-753 (the reference recomputes it in every call)
  const float DT_CRCL = 1800.0f, DLON = 3.75f, DLAT = 3.75f;
  const float pi = p.pi, kappa = p.kappa;
  const float deg = 2.f * pi * 6.371e6f / 360.f;
  const float dyy = DLAT * deg;
  mc.ccy_diff = kappa * DT_CRCL / (dyy * dyy);
  mc.ccy_adv = DT_CRCL / dyy / 2.f;
  for (int k = 1; k <= GY; ++k) {
    const float lat = DLAT * (float)k - DLAT / 2.f - 90.f;
    const float dxlat = DLON * deg * cosf(2.f * pi / 360.f * lat);
    mc.ccx_diff[k - 1] = kappa * DT_CRCL / (dxlat * dxlat);
    mc.ccx_adv[k - 1] = DT_CRCL / dxlat / 2.f;
    mc.polar[k - 1] = !(dxlat > 2.5e5f);  // f:592, 799
    {                                     // f:652-654
      const int n = f_nint(DT_CRCL / (1.f * (dxlat * dxlat) / kappa));
      const float dd = (float)std::max(1, n);
      const int dtdff2 = (int)(DT_CRCL / dd);
      mc.time2_diff[k - 1] = std::max(1, f_nint(DT_CRCL / (float)dtdff2));
      mc.ccx2_diff[k - 1] = kappa * (float)dtdff2 / (dxlat * dxlat);
    }
    {  // f:838-840
      const int n = f_nint(DT_CRCL / (dxlat / 10.0f / 1.f));
      const float dd = (float)std::max(1, n);
      const int dtdff2 = (int)(DT_CRCL / dd);
      mc.time2_adv[k - 1] = std::max(1, f_nint(DT_CRCL / (float)dtdff2));
      mc.ccx2_adv[k - 1] = (float)dtdff2 / dxlat / 2.f;
    }
  }
  return greb_assign_rows(mc.polar, mc.time2_diff, mc.time2_adv, mc.row_of_group, mc.hslot_of_row, mc.helper_row,
                          &mc.n_hslots, warp_order);
}

void greb_b200_wz(const float* z_topo, float h_scale, float* out, long n) {  // f:201-202 for any number of cells
  for (long c = 0; c < n; ++c) out[c] = expf(-z_topo[c] / h_scale);
}

void greb_build_wz(float* out, const GrebHostForcing& F, const greb_physics_par& p) {
  for (int c = 0; c < GNC; ++c) {
    out[c] = expf(-F.z_topo[c] / p.z_air);          // f:201 (== exp(-z_topo/z_air) of f:420, 458)
    out[GNC + c] = expf(-F.z_topo[c] / p.z_vapor);  // f:202
  }
}

void greb_build_initial_state(float* out, const GrebHostForcing& F, const GrebMemberConst& mc) {
  const float* t_last = &F.tclim[(size_t)(GNT - 1) * GNC];
  const float* q_last = &F.qclim[(size_t)(GNT - 1) * GNC];
  for (int c = 0; c < GNC; ++c) {
    out[GS_TS * GNC + c] = t_last[c];       // f:194
    out[GS_TA * GNC + c] = t_last[c];       // f:195
    out[GS_TO * GNC + c] = F.toclim[c];     // f:196
    out[GS_Q * GNC + c] = q_last[c];        // f:197
    float cap = 0.f;                        // f:190-191
    if (F.z_topo[c] > 0.f) cap = mc.cap_land;
    if (F.z_topo[c] <= 0.f) cap = mc.cap_ocean * F.mld0[c];
    out[GS_CAP * GNC + c] = cap;
  }
}
