/*
 * grid_oracle.c — CPU restatement of the reference's circulation (diffusion + advection sub-steps,
 * /root/reference/src/greb.f90:528-915) for an ARBITRARY xdim x ydim grid.
 *
 * TEST INFRASTRUCTURE ONLY (parity oracle of the big-grid path, BASELINE.json configs[4]).
 * The reference formulas are kept literally — same operand order, true divisions, the f:881 index
 * slip at j = xdim-2 — with xdim/ydim as run-time values.  They cannot run at 0.25 degrees as
 * written (explicit y-diffusion unstable at dt_crcl = 1800 s, dtdff2 truncates to
 * 0), so this file DECLARES two changes, both of which vanish on the reference's 96x48 grid:
 *
 *   R1  dt_crcl = 1800 * (48/ydim)^2 seconds  (keeps ccy_diff; 1800 at ydim = 48, 8 at ydim = 720),
 *       sub-steps per 12-hour step = nint(43200/dt_crcl) as in f:543;
 *   R2  the latitude that enters the zonal spacing dxlat is clamped to +-88.125 degrees, the
 *       outermost latitude of the reference grid (f:580 otherwise unchanged), and dtdff2 is
 *       floored at 1 second (f:653, f:839).
 *
 * PINNED: at xdim = 96, ydim = 48 the results are bit-identical to oracle/greb_oracle.c
 * (go_circulation), which is itself pinned to the reference source (tests/test_grid_path.py).
 * Array layout: X[k][j], k = latitude row 0..ydim-1 (south to north), j = longitude, j fastest.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int nx, ny, nsub;
  float dt_crcl, ccy_diff, ccy_adv;
  float *dxlat, *ccx_diff, *ccx_adv, *ccx2_diff, *ccx2_adv; /* [ny] */
  int *polar, *time2_diff, *time2_adv;                      /* [ny] */
} gg_geom;

static int f_nint(float x) { return (int)lroundf(x); }

/* f:543, f:578-582, f:652-654, f:749-753, f:838-840 with rules R1, R2 */
int gg_geometry(int nx, int ny, float pi, float kappa, float *dxlat, float *ccx_diff, float *ccx_adv,
                float *ccx2_diff, float *ccx2_adv, int *polar, int *time2_diff, int *time2_adv, float *scalars) {
  const float dlon = 360.f / (float)nx, dlat = 180.f / (float)ny; /* f:43-44 */
  const float dt_crcl = 1800.f * (48.f * 48.f) / ((float)ny * (float)ny); /* R1 (exact for ny = 48, 96, 720) */
  const float deg = 2.f * pi * 6.371e6f / 360.f;
  const float dyy = dlat * deg;
  scalars[0] = dt_crcl;
  scalars[1] = kappa * dt_crcl / (dyy * dyy); /* ccy, f:581 */
  scalars[2] = dt_crcl / dyy / 2.f;           /* ccy, f:752 */
  int nsub = f_nint(43200.f / dt_crcl);       /* f:543 */
  if (nsub < 1) nsub = 1;
  for (int k = 1; k <= ny; ++k) {
    float lat = dlat * (float)k - dlat / 2.f - 90.f; /* f:580 */
    if (lat > 88.125f) lat = 88.125f;                /* R2 */
    if (lat < -88.125f) lat = -88.125f;
    const float dx = dlon * deg * cosf(2.f * pi / 360.f * lat);
    dxlat[k - 1] = dx;
    ccx_diff[k - 1] = kappa * dt_crcl / (dx * dx);
    ccx_adv[k - 1] = dt_crcl / dx / 2.f;
    polar[k - 1] = !(dx > 2.5e5f); /* f:592, f:799 */
    {
      int n = f_nint(dt_crcl / (1.f * (dx * dx) / kappa));
      float dd = (float)(n > 1 ? n : 1);
      int dtdff2 = (int)(dt_crcl / dd);
      if (dtdff2 < 1) dtdff2 = 1; /* R2 */
      int t2 = f_nint(dt_crcl / (float)dtdff2);
      time2_diff[k - 1] = t2 > 1 ? t2 : 1;
      ccx2_diff[k - 1] = kappa * (float)dtdff2 / (dx * dx);
    }
    {
      int n = f_nint(dt_crcl / (dx / 10.0f / 1.f));
      float dd = (float)(n > 1 ? n : 1);
      int dtdff2 = (int)(dt_crcl / dd);
      if (dtdff2 < 1) dtdff2 = 1; /* R2 */
      int t2 = f_nint(dt_crcl / (float)dtdff2);
      time2_adv[k - 1] = t2 > 1 ? t2 : 1;
      ccx2_adv[k - 1] = (float)dtdff2 / dx / 2.f;
    }
  }
  return nsub;
}

static void pad_row(float *dst, const float *src, int nx) { /* dst[-3 .. nx+2], periodic */
  memcpy(dst, src, (size_t)nx * sizeof(float));
  dst[-3] = src[nx - 3];
  dst[-2] = src[nx - 2];
  dst[-1] = src[nx - 1];
  dst[nx] = src[0];
  dst[nx + 1] = src[1];
  dst[nx + 2] = src[2];
}

static void diff_x_row(float *out, const float *T, const float *w, float cc, int nx) { /* f:595-650 */
  for (int j = 0; j < nx; ++j)
    out[j] = cc * (10.f * (w[j - 1] * (T[j - 1] - T[j]) + w[j + 1] * (T[j + 1] - T[j])) +
                   4.f * (w[j - 2] * (T[j - 2] - T[j - 1]) + w[j - 1] * (T[j] - T[j - 1])) +
                   4.f * (w[j + 1] * (T[j] - T[j + 1]) + w[j + 2] * (T[j + 2] - T[j + 1])) +
                   1.f * (w[j - 3] * (T[j - 3] - T[j - 2]) + w[j - 2] * (T[j - 1] - T[j - 2])) +
                   1.f * (w[j + 2] * (T[j + 1] - T[j + 2]) + w[j + 3] * (T[j + 3] - T[j + 2]))) /
             20.f;
}

/* One circulation sub-step, X = (X + dx_diffuse) + dx_advec (f:546-549), for the rows
 * [r0, r1) of the global grid.  All arrays are FULL global fields [ny][nx]; rows outside
 * [r0-2, r1+2) are not read. */
void gg_substep(int nx, int ny, int r0, int r1, const float *X, const float *wz, const float *u, const float *v,
                float ccy_d, float ccy_a, const float *ccx_diff, const float *ccx_adv, const float *ccx2_diff,
                const float *ccx2_adv, const int *polar, const int *time2_diff, const int *time2_adv,
                float *Xnew) {
  float *buf = (float *)malloc((size_t)(6 * (nx + 6)) * sizeof(float));
  float *Tp = buf + 3, *wp = buf + (nx + 6) + 3, *dTxh = buf + 2 * (nx + 6), *dTx = buf + 3 * (nx + 6);
  float *dTy = buf + 4 * (nx + 6), *dd = buf + 5 * (nx + 6);
#define AT(a, k, j) ((a)[(size_t)(k) * nx + (j)])
  for (int k = r0; k < r1; ++k) {
    const float *T = X + (size_t)k * nx;
    /* ---------------- diffusion, f:556-723 ---------------- */
    if (k >= 1 && k <= ny - 2) {
      for (int j = 0; j < nx; ++j)
        dTy[j] = ccy_d * (AT(wz, k - 1, j) * (AT(X, k - 1, j) - T[j]) + AT(wz, k + 1, j) * (AT(X, k + 1, j) - T[j]));
    } else if (k == 0) {
      for (int j = 0; j < nx; ++j) dTy[j] = ccy_d * AT(wz, k + 1, j) * (-T[j] + AT(X, k + 1, j));
    } else {
      for (int j = 0; j < nx; ++j) dTy[j] = ccy_d * AT(wz, k - 1, j) * (AT(X, k - 1, j) - T[j]);
    }
    pad_row(wp, wz + (size_t)k * nx, nx);
    pad_row(Tp, T, nx);
    if (!polar[k]) {
      diff_x_row(dTx, Tp, wp, ccx_diff[k], nx);
    } else {
      for (int tt2 = 0; tt2 < time2_diff[k]; ++tt2) {
        diff_x_row(dTxh, Tp, wp, ccx2_diff[k], nx);
        for (int j = 0; j < nx; ++j) { /* f:715-716 */
          float d = dTxh[j];
          if (d <= -Tp[j]) d = -0.9f * Tp[j];
          dTxh[j] = Tp[j] + d;
        }
        pad_row(Tp, dTxh, nx);
      }
      for (int j = 0; j < nx; ++j) dTx[j] = Tp[j] - T[j]; /* f:718 */
    }
    for (int j = 0; j < nx; ++j) dd[j] = AT(wz, k, j) * (dTx[j] + dTy[j]); /* f:721 */

    /* ---------------- advection, f:726-915 ---------------- */
#define VM(j) (AT(v, k, j) >= 0.f ? AT(v, k, j) : 0.f) /* f:205-214 sign split */
#define VP(j) (AT(v, k, j) >= 0.f ? 0.f : AT(v, k, j))
#define UM(j) (AT(u, k, j) >= 0.f ? AT(u, k, j) : 0.f)
#define UP(j) (AT(u, k, j) >= 0.f ? 0.f : AT(u, k, j))
    if (k == 0) {
      for (int j = 0; j < nx; ++j)
        dTy[j] = ccy_a * (VP(j) * (AT(wz, k + 1, j) * (T[j] - AT(X, k + 1, j)) +
                                   AT(wz, k + 2, j) * (T[j] - AT(X, k + 2, j)))) / 3.f;
    } else if (k == 1) {
      for (int j = 0; j < nx; ++j)
        dTy[j] = ccy_a * (-VM(j) * (AT(wz, k - 1, j) * (T[j] - AT(X, k - 1, j))) +
                          VP(j) * (AT(wz, k + 1, j) * (T[j] - AT(X, k + 1, j)) +
                                   AT(wz, k + 2, j) * (T[j] - AT(X, k + 2, j))) / 3.f);
    } else if (k <= ny - 3) {
      for (int j = 0; j < nx; ++j)
        dTy[j] = ccy_a * (-VM(j) * (AT(wz, k - 1, j) * (T[j] - AT(X, k - 1, j)) +
                                    AT(wz, k - 2, j) * (T[j] - AT(X, k - 2, j))) +
                          VP(j) * (AT(wz, k + 1, j) * (T[j] - AT(X, k + 1, j)) +
                                   AT(wz, k + 2, j) * (T[j] - AT(X, k + 2, j)))) / 3.f;
    } else if (k == ny - 2) {
      for (int j = 0; j < nx; ++j)
        dTy[j] = ccy_a * (-VM(j) * (AT(wz, k - 1, j) * (T[j] - AT(X, k - 1, j)) +
                                    AT(wz, k - 2, j) * (T[j] - AT(X, k - 2, j))) / 3.f +
                          VP(j) * (AT(wz, k + 1, j) * (T[j] - AT(X, k + 1, j))));
    } else {
      for (int j = 0; j < nx; ++j)
        dTy[j] = ccy_a * (-VM(j) * (AT(wz, k - 1, j) * (T[j] - AT(X, k - 1, j)) +
                                    AT(wz, k - 2, j) * (T[j] - AT(X, k - 2, j)))) / 3.f;
    }
    pad_row(Tp, T, nx);
    if (!polar[k]) {
      const float ccx = ccx_adv[k];
      for (int j = 0; j < nx; ++j)
        dTx[j] = ccx * (-UM(j) * (wp[j - 1] * (Tp[j] - Tp[j - 1]) + wp[j - 2] * (Tp[j] - Tp[j - 2])) +
                        UP(j) * (wp[j + 1] * (Tp[j] - Tp[j + 1]) + wp[j + 2] * (Tp[j] - Tp[j + 2]))) / 3.f;
    } else {
      const float ccx2 = ccx2_adv[k];
      for (int tt2 = 0; tt2 < time2_adv[k]; ++tt2) {
        for (int j = 0; j < nx; ++j)
          dTxh[j] = ccx2 * (-UM(j) * (10.f * wp[j - 1] * (Tp[j] - Tp[j - 1]) + 4.f * wp[j - 2] * (Tp[j - 1] - Tp[j - 2]) +
                                      1.f * wp[j - 3] * (Tp[j - 2] - Tp[j - 3])) +
                            UP(j) * (10.f * wp[j + 1] * (Tp[j] - Tp[j + 1]) + 4.f * wp[j + 2] * (Tp[j + 1] - Tp[j + 2]) +
                                     1.f * wp[j + 3] * (Tp[j + 2] - Tp[j + 3]))) / 20.f;
        { /* f:880-888: at j = xdim-2 the reference uses jp2 = xdim-1 (should be xdim) */
          const int j = nx - 3, jm1 = j - 1, jm2 = j - 2, jm3 = j - 3, jp1 = nx - 2, jp2 = nx - 2, jp3 = 0;
          dTxh[j] = ccx2 * (-UM(j) * (10.f * wp[jm1] * (Tp[j] - Tp[jm1]) + 4.f * wp[jm2] * (Tp[jm1] - Tp[jm2]) +
                                      1.f * wp[jm3] * (Tp[jm2] - Tp[jm3])) +
                            UP(j) * (10.f * wp[jp1] * (Tp[j] - Tp[jp1]) + 4.f * wp[jp2] * (Tp[jp1] - Tp[jp2]) +
                                     1.f * wp[jp3] * (Tp[jp2] - Tp[jp3]))) / 20.f;
        }
        for (int j = 0; j < nx; ++j) { /* f:907-908 */
          float d = dTxh[j];
          if (d <= -Tp[j]) d = -0.9f * Tp[j];
          dTxh[j] = Tp[j] + d;
        }
        pad_row(Tp, dTxh, nx);
      }
      for (int j = 0; j < nx; ++j) dTx[j] = Tp[j] - T[j]; /* f:910 */
    }
    for (int j = 0; j < nx; ++j) AT(Xnew, k, j) = T[j] + dd[j] + (dTx[j] + dTy[j]); /* f:913, f:549 */
  }
#undef AT
#undef VM
#undef VP
#undef UM
#undef UP
  free(buf);
}
