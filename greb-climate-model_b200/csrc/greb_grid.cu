// greb_grid.cu — big-grid circulation path (include/greb_grid.h): one latitude band of one member
// on one GPU, for grids that do not fit one SM (BASELINE.json configs[4], 1440x720).
//
// One CTA integrates one latitude ROW for one sub-step: the row and its wz live in shared memory
// (periodic pads of 3, so every wrap case of the reference is the same expression), the polar
// sub-sub-steps (f:655-718, f:841-910) ping-pong between two shared buffers, the y-direction terms
// read rows k-2..k+2 of the previous sub-step from global memory (L2-resident: a 1440x720 field is
// 4 MB).  Expressions keep the reference's operand order; the file is compiled with -fmad=false and
// IEEE division, so results are bit-identical to the CPU restatement the tests check against (and at 96x48 to the
// reference arithmetic).  No CPU path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/greb_grid.h"

#define GG_THREADS 768
#define GG_MAXC 2  // cells per thread: xdim <= GG_THREADS * GG_MAXC = 1536 (768 threads measured best: 256 -> 2.34, 512 -> 2.91, 768 -> 3.23, 1024 -> 2.97 steps/s at 1440x720 on one B200)

struct GridArgs {
  int nx, ny, r0;                 // rows r0 + blockIdx.x
  const float *X, *wz, *u, *v;    // pointers shifted so that [k * nx + j] addresses GLOBAL row k
  float* Xnew;
  float ccy_d, ccy_a;
  const float *ccx_diff, *ccx_adv, *ccx2_diff, *ccx2_adv;  // [ny]
  const int *polar, *t2d, *t2a;                            // [ny]
};

__device__ __forceinline__ void pad_fix(float* p, int nx) {  // p points at element 0 of a padded row
  const int t = threadIdx.x;
  if (t < 3) p[-3 + t] = p[nx - 3 + t];
  else if (t < 6) p[nx + t - 3] = p[t - 3];
}

// Correctly rounded x/3 and x/20 without the division subroutine: q0 = x*RN(1/d); r = fma(-d,q0,x);
// q = fma(r,RN(1/d),q0) equals RN(x/d) for every float x whose quotient is a normal number (the same
// sequences as greb_core.h, verified exhaustively over all 2^32 inputs: tests/test_divc.py).
__device__ __forceinline__ float div3(float x) {
  const float r = 0.3333333432674407958984375f;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-3.0f, q, x), r, q);
}
__device__ __forceinline__ float div20(float x) {
  const float r = 0.0500000007450580596923828125f;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-20.0f, q, x), r, q);
}

// f:595-650 / f:659-714 at longitude j of padded rows
__device__ __forceinline__ float diff_x(const float* T, const float* w, int j, float cc) {
  return div20(cc * (10.f * (w[j - 1] * (T[j - 1] - T[j]) + w[j + 1] * (T[j + 1] - T[j])) +
               4.f * (w[j - 2] * (T[j - 2] - T[j - 1]) + w[j - 1] * (T[j] - T[j - 1])) +
               4.f * (w[j + 1] * (T[j] - T[j + 1]) + w[j + 2] * (T[j + 2] - T[j + 1])) +
               1.f * (w[j - 3] * (T[j - 3] - T[j - 2]) + w[j - 2] * (T[j - 1] - T[j - 2])) +
               1.f * (w[j + 2] * (T[j + 1] - T[j + 2]) + w[j + 3] * (T[j + 3] - T[j + 2]))));
}

__global__ void __launch_bounds__(GG_THREADS) greb_grid_substep_kernel(const GridArgs a) {
  extern __shared__ float sm[];
  const int nx = a.nx, ny = a.ny, k = a.r0 + blockIdx.x, tid = threadIdx.x;
  const int stride = nx + 6;
  float* T0 = sm + 3;               // the row as it entered the sub-step (padded)
  float* wp = sm + stride + 3;      // wz of the row (padded)
  float* A = sm + 2 * stride + 3;   // ping-pong buffers of the polar sub-sub-steps
  float* B = sm + 3 * stride + 3;
  // row base pointers (rows outside the domain are never dereferenced)
  const float* Xk = a.X + (size_t)k * nx;
  const float *Xm1 = Xk - nx, *Xm2 = Xk - 2 * nx, *Xp1 = Xk + nx, *Xp2 = Xk + 2 * nx;
  const float* Wk = a.wz + (size_t)k * nx;
  const float *Wm1 = Wk - nx, *Wm2 = Wk - 2 * nx, *Wp1 = Wk + nx, *Wp2 = Wk + 2 * nx;
  const float* Uk = a.u + (size_t)k * nx;
  const float* Vk = a.v + (size_t)k * nx;
  float* Ok = a.Xnew + (size_t)k * nx;
  for (int j = tid; j < nx; j += GG_THREADS) {
    T0[j] = Xk[j];
    wp[j] = Wk[j];
  }
  __syncthreads();
  pad_fix(T0, nx);
  pad_fix(wp, nx);
  __syncthreads();

  const int polar = a.polar[k];
  float dd[GG_MAXC], adv[GG_MAXC];  // wz*(dTx+dTy) of the diffusion, dTy then dTx+dTy of the advection
#pragma unroll
  for (int c = 0; c < GG_MAXC; ++c) {
    const int j = tid + c * GG_THREADS;
    dd[c] = 0.f;
    adv[c] = 0.f;
    if (j < nx) {
      const float T = T0[j];
      // ---- y part of the diffusion, f:587-590
      float dTy;
      if (k >= 1 && k <= ny - 2)
        dTy = a.ccy_d * (Wm1[j] * (Xm1[j] - T) + Wp1[j] * (Xp1[j] - T));
      else if (k == 0)
        dTy = a.ccy_d * Wp1[j] * (-T + Xp1[j]);
      else
        dTy = a.ccy_d * Wm1[j] * (Xm1[j] - T);
      dd[c] = dTy;
      // ---- y part of the advection, f:756-795 (five row cases, different parenthesisation)
      const float vv = Vk[j];
      const float vm = vv >= 0.f ? vv : 0.f, vp = vv >= 0.f ? 0.f : vv;  // f:205-214
      float aTy;
      if (k == 0)
        aTy = div3(a.ccy_a * (vp * (Wp1[j] * (T - Xp1[j]) + Wp2[j] * (T - Xp2[j]))));
      else if (k == 1)
        aTy = a.ccy_a * (-vm * (Wm1[j] * (T - Xm1[j])) +
                         div3(vp * (Wp1[j] * (T - Xp1[j]) + Wp2[j] * (T - Xp2[j]))));
      else if (k <= ny - 3)
        aTy = div3(a.ccy_a * (-vm * (Wm1[j] * (T - Xm1[j]) + Wm2[j] * (T - Xm2[j])) +
                              vp * (Wp1[j] * (T - Xp1[j]) + Wp2[j] * (T - Xp2[j]))));
      else if (k == ny - 2)
        aTy = a.ccy_a * (div3(-vm * (Wm1[j] * (T - Xm1[j]) + Wm2[j] * (T - Xm2[j]))) +
                         vp * (Wp1[j] * (T - Xp1[j])));
      else
        aTy = div3(a.ccy_a * (-vm * (Wm1[j] * (T - Xm1[j]) + Wm2[j] * (T - Xm2[j]))));
      adv[c] = aTy;
    }
  }

  // ---- x part of the diffusion
  if (!polar) {  // f:592-650
    const float cc = a.ccx_diff[k];
#pragma unroll
    for (int c = 0; c < GG_MAXC; ++c) {
      const int j = tid + c * GG_THREADS;
      if (j < nx) dd[c] = wp[j] * (diff_x(T0, wp, j, cc) + dd[c]);  // f:721
    }
  } else {  // f:651-718
    const int time2 = a.t2d[k];
    const float cc = a.ccx2_diff[k];
    const float* cur = T0;
    float* nxt = A;
    for (int tt = 0; tt < time2; ++tt) {
      for (int j = tid; j < nx; j += GG_THREADS) {
        float d = diff_x(cur, wp, j, cc);
        if (d <= -cur[j]) d = -0.9f * cur[j];  // f:715
        nxt[j] = cur[j] + d;                   // f:716
      }
      __syncthreads();
      pad_fix(nxt, nx);
      __syncthreads();
      cur = nxt;
      nxt = (nxt == A) ? B : A;
    }
#pragma unroll
    for (int c = 0; c < GG_MAXC; ++c) {
      const int j = tid + c * GG_THREADS;
      if (j < nx) dd[c] = wp[j] * ((cur[j] - T0[j]) + dd[c]);  // f:718, f:721
    }
    __syncthreads();  // A/B are reused below
  }

  // ---- x part of the advection
  if (!polar) {  // f:799-835
    const float cc = a.ccx_adv[k];
#pragma unroll
    for (int c = 0; c < GG_MAXC; ++c) {
      const int j = tid + c * GG_THREADS;
      if (j < nx) {
        const float uu = Uk[j];
        const float um = uu >= 0.f ? uu : 0.f, up = uu >= 0.f ? 0.f : uu;
        const float dTx = div3(cc * (-um * (wp[j - 1] * (T0[j] - T0[j - 1]) + wp[j - 2] * (T0[j] - T0[j - 2])) +
                                     up * (wp[j + 1] * (T0[j] - T0[j + 1]) + wp[j + 2] * (T0[j] - T0[j + 2]))));
        adv[c] = dTx + adv[c];  // f:913
      }
    }
  } else {  // f:837-910
    const int time2 = a.t2a[k];
    const float cc = a.ccx2_adv[k];
    const float* cur = T0;
    float* nxt = A;
    for (int tt = 0; tt < time2; ++tt) {
      for (int j = tid; j < nx; j += GG_THREADS) {
        const float uu = Uk[j];
        const float um = uu >= 0.f ? uu : 0.f, up = uu >= 0.f ? 0.f : uu;
        int jp1 = j + 1, jp2 = j + 2, jp3 = j + 3;
        if (j == nx - 3) {  // f:880-888: the reference sets jp2 = xdim-1 here (should be xdim)
          jp1 = nx - 2;
          jp2 = nx - 2;
          jp3 = 0;
        }
        float d = div20(cc * (-um * (10.f * wp[j - 1] * (cur[j] - cur[j - 1]) + 4.f * wp[j - 2] * (cur[j - 1] - cur[j - 2]) +
                                     1.f * wp[j - 3] * (cur[j - 2] - cur[j - 3])) +
                              up * (10.f * wp[jp1] * (cur[j] - cur[jp1]) + 4.f * wp[jp2] * (cur[jp1] - cur[jp2]) +
                                    1.f * wp[jp3] * (cur[jp2] - cur[jp3]))));
        if (d <= -cur[j]) d = -0.9f * cur[j];  // f:907
        nxt[j] = cur[j] + d;                   // f:908
      }
      __syncthreads();
      pad_fix(nxt, nx);
      __syncthreads();
      cur = nxt;
      nxt = (nxt == A) ? B : A;
    }
#pragma unroll
    for (int c = 0; c < GG_MAXC; ++c) {
      const int j = tid + c * GG_THREADS;
      if (j < nx) adv[c] = (cur[j] - T0[j]) + adv[c];  // f:910, f:913
    }
  }

#pragma unroll
  for (int c = 0; c < GG_MAXC; ++c) {
    const int j = tid + c * GG_THREADS;
    if (j < nx) Ok[j] = T0[j] + dd[c] + adv[c];  // f:549
  }
}

// ------------------------------------------------------------------------------------------------
struct greb_grid_handle_s {
  int nx = 0, ny = 0, k0 = 0, k1 = 0, halo = 0, device = 0;
  int kbase = 0, nrows = 0;        // stored rows [kbase, kbase + nrows)
  int valid_lo = 0, valid_hi = 0;  // rows of the current buffer that hold the current sub-step level
  int nsub = 0;
  float dt_crcl = 0, ccy_d = 0, ccy_a = 0;
  bool have_geo = false, have_fields = false;
  float *d_X[2] = {nullptr, nullptr}, *d_wz = nullptr, *d_u = nullptr, *d_v = nullptr;
  float *d_ccx_diff = nullptr, *d_ccx_adv = nullptr, *d_ccx2_diff = nullptr, *d_ccx2_adv = nullptr;
  int *d_polar = nullptr, *d_t2d = nullptr, *d_t2a = nullptr;
  int cur = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  int last_launches = 0;
  bool pending = false;  // an asynchronous batch whose elapsed time has not been read yet
  std::string err;
};

static std::string g_grid_err;
#define GCK(call)                                                    \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
      return -3;                                                     \
    }                                                                \
  } while (0)

static int gfail(greb_grid_t h, const char* msg) {
  h->err = msg;
  return -1;
}

extern "C" const char* greb_grid_last_error(greb_grid_t h) { return h ? h->err.c_str() : g_grid_err.c_str(); }

extern "C" int greb_grid_create(greb_grid_t* out, int nx, int ny, int k0, int k1, int halo_rows, int device) {
  if (!out || nx < 8 || nx > GG_THREADS * GG_MAXC || ny < 5 || k0 < 0 || k1 > ny || k0 >= k1 || halo_rows < 0 ||
      (halo_rows & 1)) {
    g_grid_err = "greb_grid_create: bad arguments (8 <= xdim <= 1536, ydim >= 5, 0 <= k0 < k1 <= ydim, even halo_rows)";
    return -1;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    g_grid_err = "greb_grid_create: no usable CUDA device; this library has no CPU fallback";
    return -2;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_grid_err = "greb_grid_create: the kernels are built for sm_100a only";
    return -2;
  }
  greb_grid_t h = new greb_grid_handle_s;
  h->nx = nx;
  h->ny = ny;
  h->k0 = k0;
  h->k1 = k1;
  h->halo = halo_rows;
  h->device = device;
  h->kbase = k0 - halo_rows < 0 ? 0 : k0 - halo_rows;
  const int top = k1 + halo_rows > ny ? ny : k1 + halo_rows;
  h->nrows = top - h->kbase;
  cudaSetDevice(device);
  const size_t fb = (size_t)h->nrows * nx * sizeof(float);
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
  float** fp[] = {&h->d_X[0], &h->d_X[1], &h->d_wz, &h->d_u, &h->d_v};
  for (float** p : fp) ok = ok && cudaMalloc((void**)p, fb) == cudaSuccess;
  float** gp[] = {&h->d_ccx_diff, &h->d_ccx_adv, &h->d_ccx2_diff, &h->d_ccx2_adv};
  for (float** p : gp) ok = ok && cudaMalloc((void**)p, (size_t)ny * sizeof(float)) == cudaSuccess;
  int** ip[] = {&h->d_polar, &h->d_t2d, &h->d_t2a};
  for (int** p : ip) ok = ok && cudaMalloc((void**)p, (size_t)ny * sizeof(int)) == cudaSuccess;
  if (!ok) {
    g_grid_err = "greb_grid_create: CUDA allocation failed";
    greb_grid_destroy(h);
    return -3;
  }
  *out = h;
  return 0;
}

extern "C" int greb_grid_destroy(greb_grid_t h) {
  if (!h) return -1;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {h->d_X[0], h->d_X[1], h->d_wz, h->d_u, h->d_v, h->d_ccx_diff, h->d_ccx_adv, h->d_ccx2_diff,
                  h->d_ccx2_adv, h->d_polar, h->d_t2d, h->d_t2a};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return 0;
}

static int f_nint(float x) { return (int)lroundf(x); }  // Fortran NINT

extern "C" int greb_grid_set_geometry(greb_grid_t h, float pi, float kappa, int* nsub_out, float* dt_out) {
  if (!h) return -1;
  const int nx = h->nx, ny = h->ny;
  std::vector<float> ccx_diff(ny), ccx_adv(ny), ccx2_diff(ny), ccx2_adv(ny);
  std::vector<int> polar(ny), t2d(ny), t2a(ny);
  const float dlon = 360.f / (float)nx, dlat = 180.f / (float)ny;                 // f:43-44
  const float dt_crcl = 1800.f * (48.f * 48.f) / ((float)ny * (float)ny);         // rule R1
  const float deg = 2.f * pi * 6.371e6f / 360.f;                                  // f:578
  const float dyy = dlat * deg;
  h->dt_crcl = dt_crcl;
  h->ccy_d = kappa * dt_crcl / (dyy * dyy);                                       // f:581
  h->ccy_a = dt_crcl / dyy / 2.f;                                                 // f:752
  h->nsub = f_nint(43200.f / dt_crcl);                                            // f:543
  if (h->nsub < 1) h->nsub = 1;
  for (int k = 1; k <= ny; ++k) {
    float lat = dlat * (float)k - dlat / 2.f - 90.f;                              // f:580
    if (lat > 88.125f) lat = 88.125f;                                             // rule R2
    if (lat < -88.125f) lat = -88.125f;
    const float dx = dlon * deg * cosf(2.f * pi / 360.f * lat);
    ccx_diff[k - 1] = kappa * dt_crcl / (dx * dx);                                // f:582
    ccx_adv[k - 1] = dt_crcl / dx / 2.f;                                          // f:753
    polar[k - 1] = !(dx > 2.5e5f);                                                // f:592, f:799
    {                                                                             // f:652-654
      const int n = f_nint(dt_crcl / (1.f * (dx * dx) / kappa));
      const float dd = (float)(n > 1 ? n : 1);
      int dtdff2 = (int)(dt_crcl / dd);
      if (dtdff2 < 1) dtdff2 = 1;                                                 // rule R2
      const int t2 = f_nint(dt_crcl / (float)dtdff2);
      t2d[k - 1] = t2 > 1 ? t2 : 1;
      ccx2_diff[k - 1] = kappa * (float)dtdff2 / (dx * dx);
    }
    {                                                                             // f:838-840
      const int n = f_nint(dt_crcl / (dx / 10.0f / 1.f));
      const float dd = (float)(n > 1 ? n : 1);
      int dtdff2 = (int)(dt_crcl / dd);
      if (dtdff2 < 1) dtdff2 = 1;
      const int t2 = f_nint(dt_crcl / (float)dtdff2);
      t2a[k - 1] = t2 > 1 ? t2 : 1;
      ccx2_adv[k - 1] = (float)dtdff2 / dx / 2.f;
    }
  }
  cudaSetDevice(h->device);
  GCK(cudaMemcpy(h->d_ccx_diff, ccx_diff.data(), ny * 4, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_ccx_adv, ccx_adv.data(), ny * 4, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_ccx2_diff, ccx2_diff.data(), ny * 4, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_ccx2_adv, ccx2_adv.data(), ny * 4, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_polar, polar.data(), ny * 4, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_t2d, t2d.data(), ny * 4, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_t2a, t2a.data(), ny * 4, cudaMemcpyHostToDevice));
  h->have_geo = true;
  if (nsub_out) *nsub_out = h->nsub;
  if (dt_out) *dt_out = dt_crcl;
  return 0;
}

extern "C" int greb_grid_set_fields(greb_grid_t h, const float* X, const float* wz, const float* u, const float* v) {
  if (!h || !X || !wz || !u || !v) return h ? gfail(h, "greb_grid_set_fields: null pointer") : -1;
  cudaSetDevice(h->device);
  const size_t off = (size_t)h->kbase * h->nx, fb = (size_t)h->nrows * h->nx * sizeof(float);
  GCK(cudaMemcpy(h->d_X[0], X + off, fb, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_X[1], X + off, fb, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_wz, wz + off, fb, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_u, u + off, fb, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_v, v + off, fb, cudaMemcpyHostToDevice));
  h->cur = 0;
  h->valid_lo = h->kbase;
  h->valid_hi = h->kbase + h->nrows;
  h->have_fields = true;
  return 0;
}

extern "C" int greb_grid_substeps_async(greb_grid_t h, int n);
extern "C" int greb_grid_sync(greb_grid_t h);

extern "C" int greb_grid_substeps(greb_grid_t h, int n) {
  const int rc = greb_grid_substeps_async(h, n);
  if (rc != 0) return rc;
  return greb_grid_sync(h);
}

extern "C" int greb_grid_sync(greb_grid_t h) {
  if (!h) return -1;
  cudaSetDevice(h->device);
  GCK(cudaStreamSynchronize(h->stream));
  if (h->pending) {
    GCK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->pending = false;
  }
  return 0;
}

extern "C" int greb_grid_substeps_async(greb_grid_t h, int n) {
  if (!h) return -1;
  if (!h->have_geo || !h->have_fields) return gfail(h, "greb_grid_substeps: geometry and fields must be set first");
  if (n < 0) return gfail(h, "greb_grid_substeps: n < 0");
  cudaSetDevice(h->device);
  if (h->pending) {
    const int rc = greb_grid_sync(h);
    if (rc != 0) return rc;
  }
  const size_t smem = (size_t)4 * (h->nx + 6) * sizeof(float);
  h->last_launches = 0;
  GCK(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < n; ++i) {
    // a row needs rows k-2..k+2 of the previous level, except across the poles (no cross-pole term)
    const int lo = h->valid_lo > 0 ? h->valid_lo + 2 : 0;
    const int hi = h->valid_hi < h->ny ? h->valid_hi - 2 : h->ny;
    if (lo > h->k0 || hi < h->k1) {
      cudaEventRecord(h->ev1, h->stream);
      cudaStreamSynchronize(h->stream);
      return gfail(h, "greb_grid_substeps: the halo is used up; exchange halos (greb_grid_halo_refreshed) first");
    }
    GridArgs a;
    a.nx = h->nx;
    a.ny = h->ny;
    a.r0 = lo;
    const ptrdiff_t shift = -(ptrdiff_t)h->kbase * h->nx;  // index by GLOBAL row
    a.X = h->d_X[h->cur] + shift;
    a.Xnew = h->d_X[h->cur ^ 1] + shift;
    a.wz = h->d_wz + shift;
    a.u = h->d_u + shift;
    a.v = h->d_v + shift;
    a.ccy_d = h->ccy_d;
    a.ccy_a = h->ccy_a;
    a.ccx_diff = h->d_ccx_diff;
    a.ccx_adv = h->d_ccx_adv;
    a.ccx2_diff = h->d_ccx2_diff;
    a.ccx2_adv = h->d_ccx2_adv;
    a.polar = h->d_polar;
    a.t2d = h->d_t2d;
    a.t2a = h->d_t2a;
    greb_grid_substep_kernel<<<hi - lo, GG_THREADS, smem, h->stream>>>(a);
    h->last_launches++;
    h->cur ^= 1;
    h->valid_lo = lo;
    h->valid_hi = hi;
  }
  GCK(cudaEventRecord(h->ev1, h->stream));
  GCK(cudaGetLastError());
  h->pending = true;
  return 0;
}

extern "C" int greb_grid_view(greb_grid_t h, float** dev_rows, int* kbase, int* nrows, int* valid_lo, int* valid_hi) {
  if (!h) return -1;
  if (dev_rows) *dev_rows = h->d_X[h->cur];
  if (kbase) *kbase = h->kbase;
  if (nrows) *nrows = h->nrows;
  if (valid_lo) *valid_lo = h->valid_lo;
  if (valid_hi) *valid_hi = h->valid_hi;
  return 0;
}

extern "C" int greb_grid_halo_refreshed(greb_grid_t h) {
  if (!h) return -1;
  h->valid_lo = h->kbase;
  h->valid_hi = h->kbase + h->nrows;
  return 0;
}

extern "C" int greb_grid_get(greb_grid_t h, float* out) {
  if (!h || !out) return -1;
  cudaSetDevice(h->device);
  GCK(cudaMemcpy(out, h->d_X[h->cur] + (size_t)(h->k0 - h->kbase) * h->nx, (size_t)(h->k1 - h->k0) * h->nx * sizeof(float),
                 cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int greb_grid_last_ms(greb_grid_t h, float* ms, int* launches) {
  if (!h) return -1;
  if (ms) *ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  return 0;
}
