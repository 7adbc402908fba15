/*
 * greb_oracle.c — CPU restatement of the reference GREB time-stepping core (see greb_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).  The reference ships no golden vectors
 * and the image has no Fortran compiler; this restatement is PINNED bit for bit against the
 * reference's own source, machine-translated statement by statement and compiled here
 * (oracle/f90_to_cpp.py -> oracle/_ref/, driven by oracle/ref.py), and against the golden vectors
 * generated from it (tests/golden/, tests/test_golden.py, tests/test_ref_pin.py): every kernel
 * routine, the default 3+50-year run (all 3,000 records), a perturbed member, the greb-original
 * control + scenario run.
 *
 * Conventions: Fortran X(i,k) (i = longitude 1..96 fastest, k = latitude 1..48) is C x[k-1][i-1].
 * All comments "f:NNN" cite /root/reference/src/greb.f90 line numbers.
 * Every expression keeps the reference's operand order and parentheses; integer literals that
 * the Fortran mixes into real expressions become the float they convert to.
 */
#include "greb_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define XD GO_XDIM
#define YD GO_YDIM
#define NT GO_NSTEP_YR
#define NC GO_NCELL

typedef float field[YD][XD];

static const int jday_mon[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31}; /* f:42 */
static const float DT = 43200.0f;      /* f:38  integer dt = 12*3600 used in real expressions */
static const float DT_CRCL = 1800.0f;  /* f:39  integer dt_crcl = 0.5*3600 */
static const int NDT_DAYS = 2;         /* f:40 */
static const float DLON = 3.75f, DLAT = 3.75f; /* f:43-44: 360./96, 180./48 (exact) */

struct go_model {
  go_physics p;
  /* mo_physics fields, f:108-120 */
  field z_topo, glacier, z_ocean, cap_surf, wz_air, wz_vapor;
  field *Tclim, *uclim, *vclim, *qclim, *mldclim, *Toclim, *cldclim, *swetclim, *dTrad;
  field *TF_correct, *qF_correct, *ToF_correct;
  field *uclim_m, *uclim_p, *vclim_m, *vclim_p;
  float sw_solar[NT][YD];
  float cap_ocean, cap_land, cap_air;
  int jday, ityr; /* 1-based like the Fortran module variables */
  /* model state (greb_model locals Ts_ini.. / Ts1.., f:171-172) */
  field Ts, Ta, To, q;
  /* mo_diagnostics, f:145-149 */
  field Tmm, Tamm, Tomm, qmm, apmm, tsmn;
  int mon, irec;
  float year;
  go_geometry geo;
  int geo_valid;
};

/* ------------------------------------------------------------------------------------------ */

void go_physics_defaults(go_physics *p) { /* f:68-104 */
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

void go_physics_original(go_physics *p) { /* src/greb.original.model.f90:63-101, :178 */
  go_physics_defaults(p);
  p->cp_land = p->cp_ocean / 4.5f;
  p->co2_flux = 340.f; /* CO2_ctrl */
}

static field *alloc_clim(void) {
  field *f = (field *)calloc(NT, sizeof(field));
  if (!f) {
    fprintf(stderr, "greb_oracle: out of memory\n");
    abort();
  }
  return f;
}

go_model *go_create(void) {
  go_model *m = (go_model *)calloc(1, sizeof(go_model));
  if (!m) return NULL;
  go_physics_defaults(&m->p);
  m->Tclim = alloc_clim();
  m->uclim = alloc_clim();
  m->vclim = alloc_clim();
  m->qclim = alloc_clim();
  m->mldclim = alloc_clim();
  m->Toclim = alloc_clim();
  m->cldclim = alloc_clim();
  m->swetclim = alloc_clim();
  m->dTrad = alloc_clim();
  m->TF_correct = alloc_clim();
  m->qF_correct = alloc_clim();
  m->ToF_correct = alloc_clim();
  m->uclim_m = alloc_clim();
  m->uclim_p = alloc_clim();
  m->vclim_m = alloc_clim();
  m->vclim_p = alloc_clim();
  m->ityr = 1;
  m->jday = 1;
  m->mon = 1;
  return m;
}

void go_destroy(go_model *m) {
  if (!m) return;
  free(m->Tclim); free(m->uclim); free(m->vclim); free(m->qclim); free(m->mldclim);
  free(m->Toclim); free(m->cldclim); free(m->swetclim); free(m->dTrad);
  free(m->TF_correct); free(m->qF_correct); free(m->ToF_correct);
  free(m->uclim_m); free(m->uclim_p); free(m->vclim_m); free(m->vclim_p);
  free(m);
}

void go_set_physics(go_model *m, const go_physics *p) {
  m->p = *p;
  m->geo_valid = 0;
}
void go_get_physics(const go_model *m, go_physics *p) { *p = m->p; }

void go_set_forcing(go_model *m, const float *z_topo, const float *glacier, const float *sw_solar,
                    const float *tclim, const float *qclim, const float *swetclim, const float *uclim,
                    const float *vclim, const float *mldclim, const float *cldclim) {
  /* f:1073-1085 */
  memcpy(m->z_topo, z_topo, sizeof(field));
  memcpy(m->glacier, glacier, sizeof(field));
  memcpy(m->sw_solar, sw_solar, sizeof m->sw_solar);
  memcpy(m->Tclim, tclim, NT * sizeof(field));
  memcpy(m->qclim, qclim, NT * sizeof(field));
  memcpy(m->swetclim, swetclim, NT * sizeof(field));
  memcpy(m->uclim, uclim, NT * sizeof(field));
  memcpy(m->vclim, vclim, NT * sizeof(field));
  memcpy(m->mldclim, mldclim, NT * sizeof(field));
  memcpy(m->cldclim, cldclim, NT * sizeof(field));
  /* f:1087-1094: Toclim = min over time of Tclim, floored at -1.7 C */
  for (int k = 0; k < YD; ++k)
    for (int i = 0; i < XD; ++i) {
      float mn = m->Tclim[0][k][i];
      for (int n = 1; n < NT; ++n)
        if (m->Tclim[n][k][i] < mn) mn = m->Tclim[n][k][i];
      if (mn - 273.15f < -1.7f) mn = -1.7f + 273.15f; /* f:1091 */
      for (int n = 0; n < NT; ++n) m->Toclim[n][k][i] = mn;
    }
}

/* ---- geometry shared by diffusion and advection ------------------------------------------ */

static int f_nint(float x) { return (int)lroundf(x); } /* Fortran NINT: half away from zero */

void go_geometry_compute(float pi, float kappa, go_geometry *g) {
  /* f:578-582 and f:749-753 (identical text in both routines) */
  float deg = 2.f * pi * 6.371e6f / 360.f;
  float dx = DLON, dy = DLAT;
  float dyy = dy * deg;
  g->deg = deg;
  g->dyy = dyy;
  g->ccy_diff = kappa * DT_CRCL / (dyy * dyy); /* f:581 */
  g->ccy_adv = DT_CRCL / dyy / 2.f;            /* f:752 */
  for (int k = 1; k <= YD; ++k) {
    float lat = DLAT * (float)k - DLAT / 2.f - 90.f;        /* f:580 */
    float dxlat = dx * deg * cosf(2.f * pi / 360.f * lat);   /* f:580 */
    g->dxlat[k - 1] = dxlat;
    g->ccx_diff[k - 1] = kappa * DT_CRCL / (dxlat * dxlat);  /* f:582 */
    g->ccx_adv[k - 1] = DT_CRCL / dxlat / 2.f;              /* f:753 */
    g->polar[k - 1] = !(dxlat > 2.5e5f);                    /* f:592, f:799 */
    {                                                       /* f:652-654 */
      int n = f_nint(DT_CRCL / (1.f * (dxlat * dxlat) / kappa));
      float dd = (float)(n > 1 ? n : 1);
      int dtdff2 = (int)(DT_CRCL / dd);
      int t2 = f_nint(DT_CRCL / (float)dtdff2);
      g->time2_diff[k - 1] = t2 > 1 ? t2 : 1;
      g->ccx2_diff[k - 1] = kappa * (float)dtdff2 / (dxlat * dxlat);
    }
    {                                                       /* f:838-840 */
      int n = f_nint(DT_CRCL / (dxlat / 10.0f / 1.f));
      float dd = (float)(n > 1 ? n : 1);
      int dtdff2 = (int)(DT_CRCL / dd);
      int t2 = f_nint(DT_CRCL / (float)dtdff2);
      g->time2_adv[k - 1] = t2 > 1 ? t2 : 1;
      g->ccx2_adv[k - 1] = (float)dtdff2 / dxlat / 2.f;
    }
  }
}

static const go_geometry *geom(const go_model *m) {
  go_model *mm = (go_model *)m;
  if (!mm->geo_valid) {
    go_geometry_compute(m->p.pi, m->p.kappa, &mm->geo);
    mm->geo_valid = 1;
  }
  return &m->geo;
}

/* padded periodic copy of one row: dst[-3..XD+2] */
static inline void pad_row(float *dst, const float *src) {
  memcpy(dst, src, XD * sizeof(float));
  dst[-3] = src[XD - 3];
  dst[-2] = src[XD - 2];
  dst[-1] = src[XD - 1];
  dst[XD] = src[0];
  dst[XD + 1] = src[1];
  dst[XD + 2] = src[2];
}

/* ---- diffusion, f:556-723 ----------------------------------------------------------------- */

/* the 7-point expression of f:595-650 / f:659-714 for all 96 longitudes of one row.
 * T and w point at element 0 of padded rows, so T[j-3..j+3] are the periodic neighbours:
 * every explicit wrap case of the reference (j=1,2,3,xdim-2,xdim-1,xdim) is this same formula
 * with the wrapped index. */
static inline void diff_x_row(float *out, const float *T, const float *w, float cc) {
  for (int j = 0; j < XD; ++j) {
    out[j] = cc * (10.f * (w[j - 1] * (T[j - 1] - T[j]) + w[j + 1] * (T[j + 1] - T[j])) +
                   4.f * (w[j - 2] * (T[j - 2] - T[j - 1]) + w[j - 1] * (T[j] - T[j - 1])) +
                   4.f * (w[j + 1] * (T[j] - T[j + 1]) + w[j + 2] * (T[j + 2] - T[j + 1])) +
                   1.f * (w[j - 3] * (T[j - 3] - T[j - 2]) + w[j - 2] * (T[j - 1] - T[j - 2])) +
                   1.f * (w[j + 2] * (T[j + 1] - T[j + 2]) + w[j + 3] * (T[j + 3] - T[j + 2]))) /
             20.f;
  }
}

void go_diffusion(const go_model *m, const float *T1_, float *dX_, const float *wz_) {
  const go_geometry *g = geom(m);
  const field *T1 = (const field *)T1_;
  const field *wz = (const field *)wz_;
  field *dX = (field *)dX_;
  float dTx[XD], dTy[XD];
  float Tp_[XD + 6], wp_[XD + 6], dTxh[XD];
  float *Tp = Tp_ + 3, *wp = wp_ + 3;
  const float ccy = g->ccy_diff;

  for (int k = 0; k < YD; ++k) { /* k is the 0-based latitude row */
    /* latitudinal, f:587-590 */
    if (k >= 1 && k <= YD - 2) {
      for (int j = 0; j < XD; ++j)
        dTy[j] = ccy * ((*wz)[k - 1][j] * ((*T1)[k - 1][j] - (*T1)[k][j]) +
                        (*wz)[k + 1][j] * ((*T1)[k + 1][j] - (*T1)[k][j]));
    } else if (k == 0) {
      for (int j = 0; j < XD; ++j) dTy[j] = ccy * (*wz)[k + 1][j] * (-(*T1)[k][j] + (*T1)[k + 1][j]);
    } else {
      for (int j = 0; j < XD; ++j) dTy[j] = ccy * (*wz)[k - 1][j] * ((*T1)[k - 1][j] - (*T1)[k][j]);
    }
    /* longitudinal */
    pad_row(wp, (*wz)[k]);
    if (!g->polar[k]) { /* f:592-650 */
      pad_row(Tp, (*T1)[k]);
      diff_x_row(dTx, Tp, wp, g->ccx_diff[k]);
    } else { /* f:651-718 */
      const int time2 = g->time2_diff[k];
      const float ccx2 = g->ccx2_diff[k];
      pad_row(Tp, (*T1)[k]); /* T1h = T1(:,k) */
      for (int tt2 = 0; tt2 < time2; ++tt2) {
        diff_x_row(dTxh, Tp, wp, ccx2);
        for (int j = 0; j < XD; ++j) { /* f:715-716 */
          float d = dTxh[j];
          if (d <= -Tp[j]) d = -0.9f * Tp[j];
          dTxh[j] = Tp[j] + d;
        }
        pad_row(Tp, dTxh);
      }
      for (int j = 0; j < XD; ++j) dTx[j] = Tp[j] - (*T1)[k][j]; /* f:718 */
    }
    for (int j = 0; j < XD; ++j) (*dX)[k][j] = (*wz)[k][j] * (dTx[j] + dTy[j]); /* f:721 */
  }
}

/* ---- advection, f:726-915 ----------------------------------------------------------------- */

void go_advection(const go_model *m, const float *T1_, float *dX_, const float *wz_) {
  const go_geometry *g = geom(m);
  const field *T1 = (const field *)T1_;
  const field *wz = (const field *)wz_;
  field *dX = (field *)dX_;
  const int it = m->ityr - 1;
  const field *um = &m->uclim_m[it], *up = &m->uclim_p[it];
  const field *vm = &m->vclim_m[it], *vp = &m->vclim_p[it];
  const float ccy = g->ccy_adv;
  float dTx[XD], dTy[XD];
  float Tp_[XD + 6], wp_[XD + 6], dTxh[XD];
  float *Tp = Tp_ + 3, *wp = wp_ + 3;

  for (int k = 0; k < YD; ++k) {
    const float *T = (*T1)[k];
    /* latitudinal, f:756-795; the five row cases differ in parenthesisation */
    if (k == 0) { /* f:759-761 */
      for (int j = 0; j < XD; ++j)
        dTy[j] = ccy * ((*vp)[k][j] * ((*wz)[k + 1][j] * (T[j] - (*T1)[k + 1][j]) +
                                       (*wz)[k + 2][j] * (T[j] - (*T1)[k + 2][j]))) / 3.f;
    } else if (k == 1) { /* f:766-769 */
      for (int j = 0; j < XD; ++j)
        dTy[j] = ccy * (-(*vm)[k][j] * ((*wz)[k - 1][j] * (T[j] - (*T1)[k - 1][j])) +
                        (*vp)[k][j] * ((*wz)[k + 1][j] * (T[j] - (*T1)[k + 1][j]) +
                                       (*wz)[k + 2][j] * (T[j] - (*T1)[k + 2][j])) / 3.f);
    } else if (k <= YD - 3) { /* f:774-778 */
      for (int j = 0; j < XD; ++j)
        dTy[j] = ccy * (-(*vm)[k][j] * ((*wz)[k - 1][j] * (T[j] - (*T1)[k - 1][j]) +
                                        (*wz)[k - 2][j] * (T[j] - (*T1)[k - 2][j])) +
                        (*vp)[k][j] * ((*wz)[k + 1][j] * (T[j] - (*T1)[k + 1][j]) +
                                       (*wz)[k + 2][j] * (T[j] - (*T1)[k + 2][j]))) / 3.f;
    } else if (k == YD - 2) { /* f:784-787 */
      for (int j = 0; j < XD; ++j)
        dTy[j] = ccy * (-(*vm)[k][j] * ((*wz)[k - 1][j] * (T[j] - (*T1)[k - 1][j]) +
                                        (*wz)[k - 2][j] * (T[j] - (*T1)[k - 2][j])) / 3.f +
                        (*vp)[k][j] * ((*wz)[k + 1][j] * (T[j] - (*T1)[k + 1][j])));
    } else { /* f:792-794 */
      for (int j = 0; j < XD; ++j)
        dTy[j] = ccy * (-(*vm)[k][j] * ((*wz)[k - 1][j] * (T[j] - (*T1)[k - 1][j]) +
                                        (*wz)[k - 2][j] * (T[j] - (*T1)[k - 2][j]))) / 3.f;
    }

    /* longitudinal, f:798-911 */
    pad_row(wp, (*wz)[k]);
    pad_row(Tp, T);
    if (!g->polar[k]) { /* f:799-835 */
      const float ccx = g->ccx_adv[k];
      for (int j = 0; j < XD; ++j)
        dTx[j] = ccx * (-(*um)[k][j] * (wp[j - 1] * (Tp[j] - Tp[j - 1]) + wp[j - 2] * (Tp[j] - Tp[j - 2])) +
                        (*up)[k][j] * (wp[j + 1] * (Tp[j] - Tp[j + 1]) + wp[j + 2] * (Tp[j] - Tp[j + 2]))) / 3.f;
    } else { /* f:837-910 */
      const int time2 = g->time2_adv[k];
      const float ccx2 = g->ccx2_adv[k];
      for (int tt2 = 0; tt2 < time2; ++tt2) {
        for (int j = 0; j < XD; ++j)
          dTxh[j] = ccx2 * (-(*um)[k][j] * (10.f * wp[j - 1] * (Tp[j] - Tp[j - 1]) +
                                            4.f * wp[j - 2] * (Tp[j - 1] - Tp[j - 2]) +
                                            1.f * wp[j - 3] * (Tp[j - 2] - Tp[j - 3])) +
                            (*up)[k][j] * (10.f * wp[j + 1] * (Tp[j] - Tp[j + 1]) +
                                           4.f * wp[j + 2] * (Tp[j + 1] - Tp[j + 2]) +
                                           1.f * wp[j + 3] * (Tp[j + 2] - Tp[j + 3]))) / 20.f;
        { /* f:880-888: at j = xdim-2 the reference sets jp1 = xdim-1, jp2 = xdim-1, jp3 = 1
           * (jp2 should be xdim).  Reproduced verbatim: 1-based j=94, jp1=95, jp2=95, jp3=1. */
          const int j = XD - 3, jm1 = j - 1, jm2 = j - 2, jm3 = j - 3;
          const int jp1 = XD - 2, jp2 = XD - 2, jp3 = 0;
          dTxh[j] = ccx2 * (-(*um)[k][j] * (10.f * wp[jm1] * (Tp[j] - Tp[jm1]) +
                                            4.f * wp[jm2] * (Tp[jm1] - Tp[jm2]) +
                                            1.f * wp[jm3] * (Tp[jm2] - Tp[jm3])) +
                            (*up)[k][j] * (10.f * wp[jp1] * (Tp[j] - Tp[jp1]) +
                                           4.f * wp[jp2] * (Tp[jp1] - Tp[jp2]) +
                                           1.f * wp[jp3] * (Tp[jp2] - Tp[jp3]))) / 20.f;
        }
        for (int j = 0; j < XD; ++j) { /* f:907-908 */
          float d = dTxh[j];
          if (d <= -Tp[j]) d = -0.9f * Tp[j];
          dTxh[j] = Tp[j] + d;
        }
        pad_row(Tp, dTxh);
      }
      for (int j = 0; j < XD; ++j) dTx[j] = Tp[j] - T[j]; /* f:910 */
    }
    for (int j = 0; j < XD; ++j) (*dX)[k][j] = dTx[j] + dTy[j]; /* f:913 */
  }
}

/* ---- circulation, f:528-553 --------------------------------------------------------------- */

void go_circulation(const go_model *m, const float *X_in, float *dX_crcl, const float *wz) {
  field X, dxd, dxa;
  int time = f_nint(DT / DT_CRCL); /* f:543 */
  if (time < 1) time = 1;
  memcpy(X, X_in, sizeof X);
  for (int tt = 0; tt < time; ++tt) {
    go_diffusion(m, &X[0][0], &dxd[0][0], wz);
    go_advection(m, &X[0][0], &dxa[0][0], wz);
    for (int c = 0; c < NC; ++c) (&X[0][0])[c] = (&X[0][0])[c] + (&dxd[0][0])[c] + (&dxa[0][0])[c]; /* f:549 */
  }
  for (int c = 0; c < NC; ++c) dX_crcl[c] = (&X[0][0])[c] - X_in[c]; /* f:551 */
}

/* ---- column physics ----------------------------------------------------------------------- */

void go_SWradiation(const go_model *m, const float *Tsurf, float *sw, float *albedo) { /* f:367-403 */
  const go_physics *p = &m->p;
  const int it = m->ityr - 1;
  for (int k = 0; k < YD; ++k)
    for (int i = 0; i < XD; ++i) {
      const int c = k * XD + i;
      const float T = Tsurf[c], z = m->z_topo[k][i];
      float a_atmos = m->cldclim[it][k][i] * p->a_cloud; /* f:380 */
      float a_surf = 0.f;
      if (z >= 0.f && T <= p->Tl_ice1) a_surf = p->a_no_ice + p->da_ice; /* f:384 */
      if (z >= 0.f && T >= p->Tl_ice2) a_surf = p->a_no_ice;             /* f:385 */
      if (z >= 0.f && T > p->Tl_ice1 && T < p->Tl_ice2)                  /* f:386-387 */
        a_surf = p->a_no_ice + p->da_ice * (1.f - (T - p->Tl_ice1) / (p->Tl_ice2 - p->Tl_ice1));
      if (z < 0.f && T <= p->To_ice1) a_surf = p->a_no_ice + p->da_ice;  /* f:389 */
      if (z < 0.f && T >= p->To_ice2) a_surf = p->a_no_ice;              /* f:390 */
      if (z < 0.f && T > p->To_ice1 && T < p->To_ice2)                   /* f:391-392 */
        a_surf = p->a_no_ice + p->da_ice * (1.f - (T - p->To_ice1) / (p->To_ice2 - p->To_ice1));
      if (m->glacier[k][i] > 0.5f) a_surf = p->a_no_ice + p->da_ice;     /* f:395 */
      albedo[c] = a_surf + a_atmos - a_surf * a_atmos;                   /* f:398 */
      sw[c] = m->sw_solar[it][k] * (1.f - albedo[c]);                    /* f:400 */
    }
}

void go_LWradiation(const go_model *m, const float *Tsurf, const float *Tair, const float *q, float CO2,
                    float *LWsurf, float *LWair_up, float *LWair_down, float *em_out) { /* f:407-434 */
  const go_physics *p = &m->p;
  const float *pe = p->p_emi; /* pe[n-1] = p_emi(n) */
  const int it = m->ityr - 1;
  for (int k = 0; k < YD; ++k)
    for (int i = 0; i < XD; ++i) {
      const int c = k * XD + i;
      const float ez = expf(-m->z_topo[k][i] / p->z_air);
      const float e_co2 = ez * CO2;               /* f:420 */
      const float e_vapor = ez * p->r_qviwv * q[c]; /* f:421 */
      const float e_cloud = m->cldclim[it][k][i]; /* f:422 */
      float em = pe[3] * logf(pe[0] * e_co2 + pe[1] * e_vapor + pe[2]) + pe[6] +
                 pe[4] * logf(pe[0] * e_co2 + pe[2]) + pe[5] * logf(pe[1] * e_vapor + pe[2]); /* f:425-427 */
      em = (pe[7] - e_cloud) / pe[8] * (em - pe[9]) + pe[9]; /* f:428 */
      const float Ts = Tsurf[c];
      const float Tr = Tair[c] + m->dTrad[it][k][i];
      LWsurf[c] = -p->sig * ((Ts * Ts) * (Ts * Ts));             /* f:430: x**4 = (x*x)*(x*x) */
      LWair_down[c] = -em * p->sig * ((Tr * Tr) * (Tr * Tr));    /* f:431 */
      LWair_up[c] = LWair_down[c];                               /* f:432 */
      em_out[c] = em;
    }
}

void go_hydro(const go_model *m, const float *Tsurf, const float *q, float *Qlat, float *Qlat_air,
              float *dq_eva, float *dq_rain) { /* f:438-469 */
  const go_physics *p = &m->p;
  const int it = m->ityr - 1;
  for (int k = 0; k < YD; ++k)
    for (int i = 0; i < XD; ++i) {
      const int c = k * XD + i;
      const float z = m->z_topo[k][i];
      const float u = m->uclim[it][k][i], v = m->vclim[it][k][i];
      float abswind = sqrtf(u * u + v * v);                         /* f:452 */
      if (z > 0.f) abswind = sqrtf(abswind * abswind + 2.0f * 2.0f); /* f:453 */
      if (z < 0.f) abswind = sqrtf(abswind * abswind + 3.0f * 3.0f); /* f:454 */
      const float T = Tsurf[c];
      float qs = 3.75e-3f * expf(17.08085f * (T - 273.15f) / (T - 273.15f + 234.175f)); /* f:457 */
      qs = qs * expf(-z / p->z_air);                                                   /* f:458 */
      Qlat[c] = (q[c] - qs) * abswind * p->cq_latent * p->rho_air * p->ce * m->swetclim[it][k][i]; /* f:460 */
      dq_eva[c] = -Qlat[c] / p->cq_latent / p->r_qviwv;      /* f:463 */
      dq_rain[c] = p->cq_rain * q[c];                        /* f:464 */
      Qlat_air[c] = -dq_rain[c] * p->cq_latent * p->r_qviwv; /* f:467 */
    }
}

void go_seaice(go_model *m, const float *Tsurf) { /* f:472-492 */
  const go_physics *p = &m->p;
  const int it = m->ityr - 1;
  for (int k = 0; k < YD; ++k)
    for (int i = 0; i < XD; ++i) {
      const float z = m->z_topo[k][i], T = Tsurf[k * XD + i];
      const float mld = m->mldclim[it][k][i];
      if (z < 0.f && T <= p->To_ice1) m->cap_surf[k][i] = m->cap_land;        /* f:483 */
      if (z < 0.f && T >= p->To_ice2) m->cap_surf[k][i] = m->cap_ocean * mld; /* f:484 */
      if (z < 0.f && T > p->To_ice1 && T < p->To_ice2)                        /* f:485-487 */
        m->cap_surf[k][i] =
            m->cap_land + (m->cap_ocean * mld - m->cap_land) / (p->To_ice2 - p->To_ice1) * (T - p->To_ice1);
      if (m->glacier[k][i] > 0.5f) m->cap_surf[k][i] = m->cap_land; /* f:490 */
    }
}

void go_deep_ocean(const go_model *m, const float *Ts, const float *To, float *dT_ocean, float *dTo) { /* f:495-525 */
  const go_physics *p = &m->p;
  const int it = m->ityr - 1;
  const int itm = (m->ityr > 1) ? it - 1 : NT - 1; /* f:507-508 */
  const float c_effmix = 0.5f;                     /* f:516 */
  for (int k = 0; k < YD; ++k)
    for (int i = 0; i < XD; ++i) {
      const int c = k * XD + i;
      const float z = m->z_topo[k][i];
      const float mld = m->mldclim[it][k][i];
      const float dmld = mld - m->mldclim[itm][k][i];
      float dto = 0.f, dtoc = 0.f; /* f:505 */
      if (z < 0.f && Ts[c] >= p->To_ice2 && dmld < 0.f) dto = -dmld / (m->z_ocean[k][i] - mld) * (Ts[c] - To[c]); /* f:511-512 */
      if (z < 0.f && Ts[c] >= p->To_ice2 && dmld > 0.f) dtoc = dmld / mld * (To[c] - Ts[c]);                      /* f:513-514 */
      dto = c_effmix * dto;   /* f:517 */
      dtoc = c_effmix * dtoc; /* f:518 */
      const float Tx = (p->To_ice2 > Ts[c]) ? p->To_ice2 : Ts[c]; /* f:521 */
      dto = dto + DT * p->co_turb * (Tx - To[c]) / (m->cap_ocean * (m->z_ocean[k][i] - mld)); /* f:522 */
      dtoc = dtoc + DT * p->co_turb * (To[c] - Tx) / (m->cap_ocean * mld);                    /* f:523 */
      dTo[c] = dto;
      dT_ocean[c] = dtoc;
    }
}

/* ---- tendencies, f:277-308 ---------------------------------------------------------------- */

typedef struct {
  field albedo, SW, LW_surf, Q_lat, Q_sens, Q_lat_air, dq_eva, dq_rain, dq_crcl, dTa_crcl, dT_ocean, dTo,
      LWair_down, LWair_up, em;
} tend_t;

static void tendencies(go_model *m, float CO2, const field Ts1, const field Ta1, const field To1,
                       const field q1, tend_t *t) {
  go_SWradiation(m, &Ts1[0][0], &t->SW[0][0], &t->albedo[0][0]);                               /* f:291 */
  go_LWradiation(m, &Ts1[0][0], &Ta1[0][0], &q1[0][0], CO2, &t->LW_surf[0][0], &t->LWair_up[0][0],
                 &t->LWair_down[0][0], &t->em[0][0]);                                          /* f:293 */
  for (int c = 0; c < NC; ++c)
    (&t->Q_sens[0][0])[c] = m->p.ct_sens * ((&Ta1[0][0])[c] - (&Ts1[0][0])[c]);                /* f:295 */
  go_hydro(m, &Ts1[0][0], &q1[0][0], &t->Q_lat[0][0], &t->Q_lat_air[0][0], &t->dq_eva[0][0],
           &t->dq_rain[0][0]);                                                                 /* f:297 */
  go_circulation(m, &Ta1[0][0], &t->dTa_crcl[0][0], &m->wz_air[0][0]);                         /* f:301 */
  go_circulation(m, &q1[0][0], &t->dq_crcl[0][0], &m->wz_vapor[0][0]);                         /* f:303 */
  go_deep_ocean(m, &Ts1[0][0], &To1[0][0], &t->dT_ocean[0][0], &t->dTo[0][0]);                 /* f:306 */
}

/* ---- greb_model preamble, f:176-216 ------------------------------------------------------- */

void go_setup(go_model *m) {
  const go_physics *p = &m->p;
  for (int n = 0; n < NT; ++n)
    for (int c = 0; c < NC; ++c) (&m->dTrad[n][0][0])[c] = -0.16f * (&m->Tclim[n][0][0])[c] - 5.f; /* f:176 */
  memset(m->z_ocean, 0, sizeof(field)); /* f:179-183 */
  for (int n = 0; n < NT; ++n)
    for (int c = 0; c < NC; ++c)
      if ((&m->mldclim[n][0][0])[c] > (&m->z_ocean[0][0])[c]) (&m->z_ocean[0][0])[c] = (&m->mldclim[n][0][0])[c];
  for (int c = 0; c < NC; ++c) (&m->z_ocean[0][0])[c] = 3.0f * (&m->z_ocean[0][0])[c];
  m->cap_ocean = p->cp_ocean * p->rho_ocean;           /* f:186 */
  m->cap_land = p->cp_land * p->rho_land * p->d_land;  /* f:187 */
  m->cap_air = p->cp_air * p->rho_air * p->d_air;      /* f:188 */
  for (int c = 0; c < NC; ++c) {                       /* f:190-191 */
    const float z = (&m->z_topo[0][0])[c];
    if (z > 0.f) (&m->cap_surf[0][0])[c] = m->cap_land;
    if (z <= 0.f) (&m->cap_surf[0][0])[c] = m->cap_ocean * (&m->mldclim[0][0][0])[c];
  }
  memcpy(m->Ts, m->Tclim[NT - 1], sizeof(field));  /* f:194 */
  memcpy(m->Ta, m->Ts, sizeof(field));             /* f:195 */
  memcpy(m->To, m->Toclim[NT - 1], sizeof(field)); /* f:196 */
  memcpy(m->q, m->qclim[NT - 1], sizeof(field));   /* f:197 */
  for (int c = 0; c < NC; ++c) {                   /* f:201-202 */
    (&m->wz_air[0][0])[c] = expf(-(&m->z_topo[0][0])[c] / p->z_air);
    (&m->wz_vapor[0][0])[c] = expf(-(&m->z_topo[0][0])[c] / p->z_vapor);
  }
  for (int n = 0; n < NT; ++n) /* f:203-216 */
    for (int c = 0; c < NC; ++c) {
      const float u = (&m->uclim[n][0][0])[c], v = (&m->vclim[n][0][0])[c];
      (&m->uclim_m[n][0][0])[c] = (u >= 0.0f) ? u : 0.0f;
      (&m->uclim_p[n][0][0])[c] = (u >= 0.0f) ? 0.0f : u;
      (&m->vclim_m[n][0][0])[c] = (v >= 0.0f) ? v : 0.0f;
      (&m->vclim_p[n][0][0])[c] = (v >= 0.0f) ? 0.0f : v;
    }
  /* static storage of the Fortran modules starts zeroed */
  memset(m->Tmm, 0, sizeof(field)); memset(m->Tamm, 0, sizeof(field)); memset(m->Tomm, 0, sizeof(field));
  memset(m->qmm, 0, sizeof(field)); memset(m->apmm, 0, sizeof(field)); memset(m->tsmn, 0, sizeof(field));
  memset(m->TF_correct, 0, NT * sizeof(field));
  memset(m->qF_correct, 0, NT * sizeof(field));
  memset(m->ToF_correct, 0, NT * sizeof(field));
  m->mon = 1; m->irec = 0; m->year = 0.f;
  m->geo_valid = 0;
}

/* ---- diagnostics, f:929-959 (only the part that is ever output) --------------------------- */

static int diagnostics(go_model *m, const field ts0, float *gmean_out) {
  for (int c = 0; c < NC; ++c) (&m->tsmn[0][0])[c] = (&m->tsmn[0][0])[c] + (&ts0[0][0])[c]; /* f:945 */
  if (m->ityr == NT) {                                                                      /* f:948 */
    float s = 0.f;
    for (int c = 0; c < NC; ++c) {
      (&m->tsmn[0][0])[c] = (&m->tsmn[0][0])[c] / (float)NT; /* f:949 */
      s = s + (&m->tsmn[0][0])[c];                           /* f:954 sum(), array element order */
    }
    if (gmean_out) *gmean_out = s / (float)(XD * YD) - 273.15f;
    memset(m->tsmn, 0, sizeof(field)); /* f:955 */
    return 1;
  }
  return 0;
}

/* ---- output, f:962-987 -------------------------------------------------------------------- */

static int output(go_model *m, int it, const field ts0, const field ta0, const field to0, const field q0,
                  const field albedo, float *out5) {
  for (int c = 0; c < NC; ++c) { /* f:974 */
    (&m->Tmm[0][0])[c] = (&m->Tmm[0][0])[c] + (&ts0[0][0])[c];
    (&m->Tamm[0][0])[c] = (&m->Tamm[0][0])[c] + (&ta0[0][0])[c];
    (&m->Tomm[0][0])[c] = (&m->Tomm[0][0])[c] + (&to0[0][0])[c];
    (&m->qmm[0][0])[c] = (&m->qmm[0][0])[c] + (&q0[0][0])[c];
    (&m->apmm[0][0])[c] = (&m->apmm[0][0])[c] + (&albedo[0][0])[c];
  }
  int sumdays = 0;
  for (int i = 0; i < m->mon; ++i) sumdays += jday_mon[i];
  const float r = (float)it / (float)NDT_DAYS;
  if (m->jday == sumdays && r == (float)f_nint(r)) { /* f:975-976 */
    const float ndm = (float)(jday_mon[m->mon - 1] * NDT_DAYS); /* f:977 */
    if (out5) {
      for (int c = 0; c < NC; ++c) { /* f:978-982 */
        out5[0 * NC + c] = (&m->Tmm[0][0])[c] / ndm;
        out5[1 * NC + c] = (&m->Tamm[0][0])[c] / ndm;
        out5[2 * NC + c] = (&m->Tomm[0][0])[c] / ndm;
        out5[3 * NC + c] = (&m->qmm[0][0])[c] / ndm;
        out5[4 * NC + c] = (&m->apmm[0][0])[c] / ndm;
      }
    }
    m->irec += 5;
    memset(m->Tmm, 0, sizeof(field)); memset(m->Tamm, 0, sizeof(field)); memset(m->Tomm, 0, sizeof(field));
    memset(m->qmm, 0, sizeof(field)); memset(m->apmm, 0, sizeof(field)); /* f:983 */
    m->mon = m->mon + 1;
    if (m->mon == 13) m->mon = 1; /* f:984 */
    return 1;
  }
  return 0;
}

/* ---- time_loop, f:239-274 ----------------------------------------------------------------- */

static int time_loop_impl(go_model *m, int it, float CO2, float *out5, float *gmean, int *year_end) {
  tend_t *t = (tend_t *)malloc(sizeof(tend_t));
  field Ts0, Ta0, To0, q0;
  m->jday = ((it - 1) / NDT_DAYS) % 365 + 1; /* f:251 */
  m->ityr = (it - 1) % NT + 1;               /* f:252 */
  const int iy = m->ityr - 1;
  tendencies(m, CO2, m->Ts, m->Ta, m->To, m->q, t);
  for (int c = 0; c < NC; ++c) {
    const float Ts1 = (&m->Ts[0][0])[c], Ta1 = (&m->Ta[0][0])[c], To1 = (&m->To[0][0])[c], q1 = (&m->q[0][0])[c];
    /* f:258 */
    (&Ts0[0][0])[c] = Ts1 + (&t->dT_ocean[0][0])[c] +
                      DT * ((&t->SW[0][0])[c] + (&t->LW_surf[0][0])[c] - (&t->LWair_down[0][0])[c] +
                            (&t->Q_lat[0][0])[c] + (&t->Q_sens[0][0])[c] + (&m->TF_correct[iy][0][0])[c]) /
                          (&m->cap_surf[0][0])[c];
    /* f:260 */
    (&Ta0[0][0])[c] = Ta1 + (&t->dTa_crcl[0][0])[c] +
                      DT * ((&t->LWair_up[0][0])[c] + (&t->LWair_down[0][0])[c] -
                            (&t->em[0][0])[c] * (&t->LW_surf[0][0])[c] + (&t->Q_lat_air[0][0])[c] -
                            (&t->Q_sens[0][0])[c]) / m->cap_air;
    /* f:262 */
    (&To0[0][0])[c] = To1 + (&t->dTo[0][0])[c] + (&m->ToF_correct[iy][0][0])[c];
    /* f:264-266 */
    float dq = DT * ((&t->dq_eva[0][0])[c] + (&t->dq_rain[0][0])[c]) + (&t->dq_crcl[0][0])[c] +
               (&m->qF_correct[iy][0][0])[c];
    if (dq <= -q1) dq = -0.9f * q1;
    (&q0[0][0])[c] = q1 + dq;
  }
  go_seaice(m, &Ts0[0][0]);                                            /* f:268 */
  int wrote = output(m, it, Ts0, Ta0, To0, q0, t->albedo, out5);       /* f:270 */
  int ye = diagnostics(m, Ts0, gmean);                                 /* f:272 */
  if (year_end) *year_end = ye;
  memcpy(m->Ts, Ts0, sizeof(field)); memcpy(m->Ta, Ta0, sizeof(field)); /* f:232 */
  memcpy(m->To, To0, sizeof(field)); memcpy(m->q, q0, sizeof(field));
  free(t);
  return wrote;
}

int go_time_loop(go_model *m, int it, float co2, float *out5) {
  return time_loop_impl(m, it, co2, out5, NULL, NULL);
}

/* ---- qflux_correction, f:311-364 ---------------------------------------------------------- */

void go_qflux_correction(go_model *m, int years, float co2) {
  tend_t *t = (tend_t *)malloc(sizeof(tend_t));
  field Ts0, Ta0, To0, q0;
  for (int it = 1; it <= years * NDT_DAYS * 365; ++it) { /* f:325 */
    m->jday = ((it - 1) / NDT_DAYS) % 365 + 1;           /* f:326 */
    m->ityr = (it - 1) % NT + 1;                         /* f:327 */
    const int iy = m->ityr - 1;
    tendencies(m, co2, m->Ts, m->Ta, m->To, m->q, t);    /* f:328-330 */
    for (int c = 0; c < NC; ++c) {
      const float Ts1 = (&m->Ts[0][0])[c], Ta1 = (&m->Ta[0][0])[c], To1 = (&m->To[0][0])[c], q1 = (&m->q[0][0])[c];
      const float cap = (&m->cap_surf[0][0])[c];
      const float dTs = DT * ((&t->SW[0][0])[c] + (&t->LW_surf[0][0])[c] - (&t->LWair_down[0][0])[c] +
                              (&t->Q_lat[0][0])[c] + (&t->Q_sens[0][0])[c]) / cap; /* f:333 */
      float ts0 = Ts1 + dTs + (&t->dT_ocean[0][0])[c];                              /* f:334 */
      const float dTa = DT * ((&t->LWair_up[0][0])[c] + (&t->LWair_down[0][0])[c] -
                              (&t->em[0][0])[c] * (&t->LW_surf[0][0])[c] + (&t->Q_lat_air[0][0])[c] -
                              (&t->Q_sens[0][0])[c]) / m->cap_air;                  /* f:336 */
      const float ta0 = Ta1 + dTa + (&t->dTa_crcl[0][0])[c];                        /* f:337 */
      float to0 = To1 + (&t->dTo[0][0])[c];                                         /* f:339 */
      const float dq = DT * ((&t->dq_eva[0][0])[c] + (&t->dq_rain[0][0])[c]);       /* f:341 */
      float qq0 = q1 + dq + (&t->dq_crcl[0][0])[c];                                 /* f:342 */
      const float T_error = (&m->Tclim[iy][0][0])[c] - ts0;                         /* f:344 */
      const float tf = T_error * cap / DT;                                          /* f:345 */
      (&m->TF_correct[iy][0][0])[c] = tf;
      ts0 = Ts1 + dTs + (&t->dT_ocean[0][0])[c] + tf * DT / cap;                    /* f:347 */
      const float tof = (&m->Toclim[iy][0][0])[c] - to0;                            /* f:349 */
      (&m->ToF_correct[iy][0][0])[c] = tof;
      to0 = To1 + (&t->dTo[0][0])[c] + tof;                                         /* f:351 */
      const float qf = (&m->qclim[iy][0][0])[c] - qq0;                              /* f:353 */
      (&m->qF_correct[iy][0][0])[c] = qf;
      qq0 = q1 + dq + (&t->dq_crcl[0][0])[c] + qf;                                  /* f:355 */
      (&Ts0[0][0])[c] = ts0; (&Ta0[0][0])[c] = ta0; (&To0[0][0])[c] = to0; (&q0[0][0])[c] = qq0;
    }
    go_seaice(m, &Ts0[0][0]);      /* f:357 */
    diagnostics(m, Ts0, NULL);     /* f:359 */
    memcpy(m->Ts, Ts0, sizeof(field)); memcpy(m->Ta, Ta0, sizeof(field)); /* f:361 */
    memcpy(m->To, To0, sizeof(field)); memcpy(m->q, q0, sizeof(field));
  }
  free(t);
}

/* ---- scenario loop, f:226-234 ------------------------------------------------------------- */

void go_run_scenario(go_model *m, int years, const float *co2_ppm, int year0, float *out, float *gmean,
                     int continue_run) {
  if (!continue_run) { /* f:227 (Tomm is NOT reset there) */
    m->year = (float)year0;
    m->mon = 1;
    m->irec = 0;
    memset(m->Tmm, 0, sizeof(field)); memset(m->Tamm, 0, sizeof(field));
    memset(m->qmm, 0, sizeof(field)); memset(m->apmm, 0, sizeof(field));
  }
  float year = (float)year0;
  size_t nrec = 0;
  for (int it = 1; it <= years * NT; ++it) { /* f:228 */
    const int idx = (int)(year - (float)year0 + 1.f); /* f:924 co2_level */
    const float CO2 = co2_ppm[idx - 1];
    float gm = 0.f;
    int ye = 0;
    int wrote = time_loop_impl(m, it, CO2, out ? out + nrec * NC : NULL, &gm, &ye); /* f:231 */
    if (wrote) nrec += 5;
    if (ye && gmean) gmean[(it - 1) / NT] = gm;
    if (it % NT == 0) year = year + 1.f; /* f:233 */
  }
  m->year = year;
}

/* ---- accessors ---------------------------------------------------------------------------- */

static float *state_ptr(go_model *m, int which) {
  switch (which) {
    case GO_TS: return &m->Ts[0][0];
    case GO_TA: return &m->Ta[0][0];
    case GO_TO: return &m->To[0][0];
    case GO_Q: return &m->q[0][0];
    case GO_CAP: return &m->cap_surf[0][0];
  }
  return NULL;
}
void go_get_state(const go_model *m, int which, float *out) {
  memcpy(out, state_ptr((go_model *)m, which), sizeof(field));
}
void go_set_state(go_model *m, int which, const float *in) { memcpy(state_ptr(m, which), in, sizeof(field)); }
const float *go_fluxcorr(const go_model *m, int which) {
  return which == 0 ? &m->TF_correct[0][0][0] : which == 1 ? &m->qF_correct[0][0][0] : &m->ToF_correct[0][0][0];
}
void go_get_derived(const go_model *m, int which, float *out) {
  const void *src = which == 0 ? (const void *)m->wz_air
                    : which == 1 ? (const void *)m->wz_vapor
                    : which == 2 ? (const void *)m->z_ocean
                                 : (const void *)m->Toclim[0];
  memcpy(out, src, sizeof(field));
}
void go_set_ityr(go_model *m, int ityr) { m->ityr = ityr; }

/* the host libm's expf / logf on arrays: what the reference's exp() / log() of f:422-424, f:457 call when it
 * is built with gfortran + glibc; the checker of the device restatement (greb_b200_device_libm) */
void go_libm_array(int which, const float *x, float *y, int n) {
    for (int i = 0; i < n; ++i) y[i] = which == 0 ? expf(x[i]) : logf(x[i]);
}
