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
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
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

// one latitude row k, one sub-step: X (level L) -> Xnew (level L+1).  `peer` (may be null) = a second
// destination for the new row, indexed like Xnew: the neighbour band's halo row in ITS buffer (peer memory
// over NVLink; persistent path).  All threads of the CTA call it with the same k.
__device__ __forceinline__ void grid_row_substep(const GridArgs& a, int k, float* sm, float* peer) {
  const int nx = a.nx, ny = a.ny, tid = threadIdx.x;
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
  // The circulating field is read with ld.global.cg (L2 only): in the persistent kernel the rows of level L
  // were written by other CTAs — or by the neighbour GPU — while this SM's L1 may still hold level L-2.
  for (int j = tid; j < nx; j += GG_THREADS) {
    const float t = __ldcg(Xk + j), w = Wk[j];
    T0[j] = t;
    wp[j] = w;
    if (j < 3) {                 // the periodic pads are written by the threads that own the wrapped cells:
      T0[nx + j] = t;            // one barrier per pass instead of store / barrier / pad / barrier
      wp[nx + j] = w;
    }
    if (j >= nx - 3) {
      T0[j - nx] = t;
      wp[j - nx] = w;
    }
  }
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
        dTy = a.ccy_d * (Wm1[j] * (__ldcg(Xm1 + j) - T) + Wp1[j] * (__ldcg(Xp1 + j) - T));
      else if (k == 0)
        dTy = a.ccy_d * Wp1[j] * (-T + __ldcg(Xp1 + j));
      else
        dTy = a.ccy_d * Wm1[j] * (__ldcg(Xm1 + j) - T);
      dd[c] = dTy;
      // ---- y part of the advection, f:756-795 (five row cases, different parenthesisation)
      const float vv = Vk[j];
      const float vm = vv >= 0.f ? vv : 0.f, vp = vv >= 0.f ? 0.f : vv;  // f:205-214
      float aTy;
      if (k == 0)
        aTy = div3(a.ccy_a * (vp * (Wp1[j] * (T - __ldcg(Xp1 + j)) + Wp2[j] * (T - __ldcg(Xp2 + j)))));
      else if (k == 1)
        aTy = a.ccy_a * (-vm * (Wm1[j] * (T - __ldcg(Xm1 + j))) +
                         div3(vp * (Wp1[j] * (T - __ldcg(Xp1 + j)) + Wp2[j] * (T - __ldcg(Xp2 + j)))));
      else if (k <= ny - 3)
        aTy = div3(a.ccy_a * (-vm * (Wm1[j] * (T - __ldcg(Xm1 + j)) + Wm2[j] * (T - __ldcg(Xm2 + j))) +
                              vp * (Wp1[j] * (T - __ldcg(Xp1 + j)) + Wp2[j] * (T - __ldcg(Xp2 + j)))));
      else if (k == ny - 2)
        aTy = a.ccy_a * (div3(-vm * (Wm1[j] * (T - __ldcg(Xm1 + j)) + Wm2[j] * (T - __ldcg(Xm2 + j)))) +
                         vp * (Wp1[j] * (T - __ldcg(Xp1 + j))));
      else
        aTy = div3(a.ccy_a * (-vm * (Wm1[j] * (T - __ldcg(Xm1 + j)) + Wm2[j] * (T - __ldcg(Xm2 + j)))));
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
        const float r = cur[j] + d;            // f:716
        nxt[j] = r;
        if (j < 3) nxt[nx + j] = r;
        if (j >= nx - 3) nxt[j - nx] = r;
      }
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
        const float r = cur[j] + d;            // f:908
        nxt[j] = r;
        if (j < 3) nxt[nx + j] = r;
        if (j >= nx - 3) nxt[j - nx] = r;
      }
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
    if (j < nx) {
      const float r = T0[j] + dd[c] + adv[c];  // f:549
      Ok[j] = r;
      if (peer) peer[(size_t)k * nx + j] = r;
    }
  }
}

__global__ void __launch_bounds__(GG_THREADS) greb_grid_substep_kernel(const GridArgs a) {
  extern __shared__ float sm[];
  grid_row_substep(a, a.r0 + blockIdx.x, sm, nullptr);
}

// ------------------------------------------------------------------------------------------------
//   Persistent path: ONE cooperative launch advances the band (one or two fields) n sub-steps with
//   no barrier at all — a dataflow over (level, field, row) work items.
//   Every row carries a level counter in global memory.  Row k may go from level L to L+1 as soon as the
//   rows k-2..k+2 (those that exist) have reached level L: then they hold the values it reads, and they have
//   finished reading the level L-1 values of row k that it is about to overwrite (double buffering).
//   CTAs take items from ONE monotonic counter, level-major, within a level the longest rows first; an item
//   only waits for items earlier in that order, which are already running on co-resident CTAs, so nothing
//   can deadlock and a slow row (the pole rows iterate 8 times) delays only its neighbourhood, not the band.
//   Halo exchange without the host: the CTA that computes one of the band's two outermost rows stores the
//   new row ALSO into the neighbour band's halo row (peer memory, NVLink) and then releases that row's level
//   counter in the neighbour's memory; the neighbour's rows next to the halo wait on it like on any other row.
// ------------------------------------------------------------------------------------------------
struct PField {
  const float* X[2];        // level L in X[L & 1]; pointers shifted to GLOBAL row indexing
  float* Xw[2];
  const float *wz, *u, *v;
  int* level;               // [stored rows] level of every row incl. the halo rows, global-row shifted
  float* nbX[2][2];         // [side 0 = south, 1 = north][buffer]: neighbour's field buffers (shifted), or null
  int* nblevel[2];          // neighbour's row-level array (shifted), or null
};
struct PArgs {
  GridArgs g;               // geometry (X/Xnew/wz/u/v filled per item)
  PField f[2];
  int nfields, k0, k1, level0, nsub, nitems;
  const int* order;         // [nitems] field * 65536 + row, longest rows first
  unsigned* counter;        // the work counter
  int* error;               // set to 1 if a wait timed out
  long long timeout;        // cycles
};

__global__ void __launch_bounds__(GG_THREADS, 2) greb_grid_persistent_kernel(const PArgs a) {
  extern __shared__ float sm[];
  __shared__ int next_item;
  GridArgs g = a.g;
  const int total = a.nsub * a.nitems;             // < 2^31 (checked by the host)
  int item = blockIdx.x;                           // the first gridDim.x items are taken statically
  // Roles: warp 1 fetches the next item and polls the dependencies, thread 0 publishes the finished row.  The
  // publication (fence + release stores, ~1 us) of item i runs beside warp 1's polling for item i+1 — the
  // other 22 warps only ever wait at the two barriers that the data flow itself needs.
  while (item < total) {
    const int n = item / a.nitems, L = a.level0 + n;
    const int code = a.order[item - n * a.nitems];
    const PField& f = a.f[code >> 16];
    const int k = code & 0xffff;
    int fetched = 0;
    if (threadIdx.x == 32) fetched = (int)atomicAdd(a.counter, 1u);   // the next item: in flight during this one
#ifndef GG_DBG_NOWAIT   // timing experiment only (wrong results): no dependency waits
    // dependencies: rows k-2 .. k+2 at level >= L (five lanes of warp 1 poll one counter each)
    if (threadIdx.x >= 32 && threadIdx.x < 37) {
      const int kk = k - 2 + (int)threadIdx.x - 32;
      if (kk >= 0 && kk < a.g.ny) {   // incl. row k itself: its level-L values come from another CTA's item
        const long long t0 = clock64();
        // relaxed polling (L2): what makes the row's VALUES visible is the producer's fence before its release
        // store plus the fact that the field is read with ld.global.cg (L2), never from this SM's L1
        while (*reinterpret_cast<const volatile int*>(f.level + kk) < L) {
          if (clock64() - t0 > a.timeout) {
            *a.error = 1;
            break;
          }
        }
      }
    }
#endif
    __syncthreads();              // dependencies seen -> everybody; the previous item's shared memory is free
    const int side = (k < a.k0 + 2) ? 0 : (k >= a.k1 - 2 ? 1 : -1);
    float* peer = (side >= 0) ? f.nbX[side][(L + 1) & 1] : nullptr;
    g.X = f.X[L & 1];
    g.Xnew = f.Xw[(L + 1) & 1];
    g.wz = f.wz;
    g.u = f.u;
    g.v = f.v;
    grid_row_substep(g, k, sm, peer);
    if (threadIdx.x == 32) next_item = fetched;
    __syncthreads();              // every thread's stores of the new row precede thread 0's fence (CTA scope) ...
    item = (int)gridDim.x + next_item;
    if (threadIdx.x == 0) {
      // ... which makes them visible GPU-wide — system-wide only for the two rows that also went to the
      // neighbour GPU (a system-scope fence costs microseconds) — before the row's level counter.  The
      // counters themselves are plain volatile stores behind that fence.
      if (peer) __threadfence_system();
      else __threadfence();
      *reinterpret_cast<volatile int*>(f.level + k) = L + 1;
      if (peer) *reinterpret_cast<volatile int*>(f.nblevel[side] + k) = L + 1;
    }
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
  // persistent path
  int* d_level = nullptr;      // [nrows] level of every stored row (halo rows: written by the neighbours)
  bool level_stale = false;    // the launch-per-sub-step path advanced the band since d_level was written
  int* d_flags = nullptr;      // [4] error word
  unsigned* d_bar = nullptr;   // [0] work counter (leader handle of a group)
  int* d_order = nullptr;      // work-item order (leader handle)
  int order_items = 0, order_fields = 0;
  int level = 0;               // sub-steps done since set_fields
  float* nbX[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // neighbour buffers opened through CUDA IPC
  int* nblevel[2] = {nullptr, nullptr};
  int nb_kbase[2] = {0, 0};
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
  ok = ok && cudaMalloc((void**)&h->d_flags, 16 * sizeof(int)) == cudaSuccess &&
       cudaMemset(h->d_flags, 0, 16 * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&h->d_level, (size_t)h->nrows * sizeof(int)) == cudaSuccess &&
       cudaMemset(h->d_level, 0, (size_t)h->nrows * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&h->d_bar, 4 * sizeof(unsigned)) == cudaSuccess &&
       cudaMemset(h->d_bar, 0, 4 * sizeof(unsigned)) == cudaSuccess;
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
  for (int sd = 0; sd < 2; ++sd) {
    for (int b = 0; b < 2; ++b)
      if (h->nbX[sd][b]) cudaIpcCloseMemHandle(h->nbX[sd][b]);
    if (h->nblevel[sd]) cudaIpcCloseMemHandle(h->nblevel[sd]);
  }
  void* ptrs[] = {h->d_X[0], h->d_X[1], h->d_wz, h->d_u, h->d_v, h->d_ccx_diff, h->d_ccx_adv, h->d_ccx2_diff,
                  h->d_ccx2_adv, h->d_polar, h->d_t2d, h->d_t2a, h->d_flags, h->d_bar, h->d_order, h->d_level};
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
  GCK(cudaMemset(h->d_flags, 0, 16 * sizeof(int)));
  GCK(cudaMemset(h->d_level, 0, (size_t)h->nrows * sizeof(int)));
  h->level_stale = false;
  h->cur = 0;
  h->level = 0;
  h->valid_lo = h->kbase;
  h->valid_hi = h->kbase + h->nrows;
  h->have_fields = true;
  return 0;
}

// the winds of another step (full global host fields), levels and buffers untouched
extern "C" int greb_grid_set_winds(greb_grid_t h, const float* u, const float* v) {
  if (!h || !u || !v) return h ? gfail(h, "greb_grid_set_winds: null pointer") : -1;
  if (!h->have_fields) return gfail(h, "greb_grid_set_winds: set_fields first");
  cudaSetDevice(h->device);
  const size_t off = (size_t)h->kbase * h->nx, fb = (size_t)h->nrows * h->nx * sizeof(float);
  GCK(cudaMemcpy(h->d_u, u + off, fb, cudaMemcpyHostToDevice));
  GCK(cudaMemcpy(h->d_v, v + off, fb, cudaMemcpyHostToDevice));
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
    h->level++;          // invariant: the current level lives in d_X[level & 1] (the persistent path relies on it)
    h->level_stale = true;
    h->valid_lo = lo;
    h->valid_hi = hi;
  }
  GCK(cudaEventRecord(h->ev1, h->stream));
  GCK(cudaGetLastError());
  h->pending = true;
  return 0;
}

// ---- persistent path: IPC plumbing + the cooperative launch ----------------------------------------
struct GridIpcBlob {
  cudaIpcMemHandle_t X[2], flags;
  int kbase, nrows, nx, k0, k1, pad[3];
};

extern "C" int greb_grid_ipc_bytes(void) { return (int)sizeof(GridIpcBlob); }

extern "C" int greb_grid_ipc_export(greb_grid_t h, void* out) {
  if (!h || !out) return -1;
  cudaSetDevice(h->device);
  GridIpcBlob b;
  memset(&b, 0, sizeof b);
  GCK(cudaIpcGetMemHandle(&b.X[0], h->d_X[0]));
  GCK(cudaIpcGetMemHandle(&b.X[1], h->d_X[1]));
  GCK(cudaIpcGetMemHandle(&b.flags, h->d_level));
  b.kbase = h->kbase;
  b.nrows = h->nrows;
  b.nx = h->nx;
  b.k0 = h->k0;
  b.k1 = h->k1;
  memcpy(out, &b, sizeof b);
  return 0;
}

extern "C" int greb_grid_ipc_import(greb_grid_t h, int side, const void* in) {
  if (!h || !in || side < 0 || side > 1) return h ? gfail(h, "greb_grid_ipc_import: bad arguments") : -1;
  cudaSetDevice(h->device);
  GridIpcBlob b;
  memcpy(&b, in, sizeof b);
  if (b.nx != h->nx || (side == 0 ? b.k1 != h->k0 : b.k0 != h->k1))
    return gfail(h, "greb_grid_ipc_import: that band is not the neighbour on this side");
  // the neighbour must store my two outermost rows as halo rows
  if (side == 0 ? b.kbase + b.nrows < h->k0 + 2 : b.kbase > h->k1 - 2)
    return gfail(h, "greb_grid_ipc_import: the neighbour's halo is thinner than 2 rows");
  for (int i = 0; i < 2; ++i) {
    void* p = nullptr;
    GCK(cudaIpcOpenMemHandle(&p, b.X[i], cudaIpcMemLazyEnablePeerAccess));
    h->nbX[side][i] = (float*)p;
  }
  void* p = nullptr;
  GCK(cudaIpcOpenMemHandle(&p, b.flags, cudaIpcMemLazyEnablePeerAccess));
  h->nblevel[side] = (int*)p;
  h->nb_kbase[side] = b.kbase;
  return 0;
}

extern "C" int greb_grid_run_persistent(greb_grid_t* hs, int nfields, int n) {
  if (!hs || nfields < 1 || nfields > 2 || !hs[0]) return -1;
  greb_grid_t h = hs[0];
  for (int f = 0; f < nfields; ++f) {
    greb_grid_t g = hs[f];
    if (!g || g->device != h->device || g->nx != h->nx || g->ny != h->ny || g->k0 != h->k0 || g->k1 != h->k1 ||
        g->kbase != h->kbase || g->level != h->level)
      return gfail(h, "greb_grid_run_persistent: the handles of a group must be the same band at the same level");
    if (!g->have_geo || !g->have_fields) return gfail(h, "greb_grid_run_persistent: geometry and fields must be set first");
    if (g->cur != (g->level & 1)) return gfail(h, "greb_grid_run_persistent: internal: buffer parity lost");
    if ((g->k0 > 0 && (g->halo < 2 || !g->nbX[0][0])) || (g->k1 < g->ny && (g->halo < 2 || !g->nbX[1][0])))
      return gfail(h, "greb_grid_run_persistent: an inner band needs halo_rows >= 2 and both neighbours imported");
  }
  if (n < 0 || (long long)n * nfields * (h->k1 - h->k0) > 2000000000LL)
    return gfail(h, "greb_grid_run_persistent: n < 0 or more than 2e9 work items in one call");
  if (h->k1 - h->k0 < 4 && (h->k0 > 0 || h->k1 < h->ny)) return gfail(h, "greb_grid_run_persistent: bands need >= 4 rows");
  cudaSetDevice(h->device);
  if (h->pending) {
    const int rc = greb_grid_sync(h);
    if (rc != 0) return rc;
  }
  // work-item order: boundary rows first (their results travel to the neighbours), then by the number of
  // polar sub-sub-steps (the longest rows start first), then south to north
  if (!h->d_order || h->order_fields != nfields) {
    std::vector<int> t2(h->ny);
    GCK(cudaMemcpy(t2.data(), h->d_t2d, h->ny * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<std::pair<long, int>> keyed;
    for (int f = 0; f < nfields; ++f)
      for (int k = h->k0; k < h->k1; ++k) {
        const bool edge = (h->k0 > 0 && k < h->k0 + 2) || (h->k1 < h->ny && k >= h->k1 - 2);
        const long key = (edge ? 0 : 1) * 1000000L + (1000 - t2[k]) * 1000L + (k - h->k0);
        keyed.push_back({key * 2 + f, f * 65536 + k});
      }
    std::sort(keyed.begin(), keyed.end());
    std::vector<int> order;
    for (auto& kv : keyed) order.push_back(kv.second);
    if (h->d_order) cudaFree(h->d_order);
    h->d_order = nullptr;
    GCK(cudaMalloc((void**)&h->d_order, order.size() * sizeof(int)));
    GCK(cudaMemcpy(h->d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->order_items = (int)order.size();
    h->order_fields = nfields;
  }
  PArgs a;
  memset(&a, 0, sizeof a);
  a.g.nx = h->nx;
  a.g.ny = h->ny;
  a.g.ccy_d = h->ccy_d;
  a.g.ccy_a = h->ccy_a;
  a.g.ccx_diff = h->d_ccx_diff;
  a.g.ccx_adv = h->d_ccx_adv;
  a.g.ccx2_diff = h->d_ccx2_diff;
  a.g.ccx2_adv = h->d_ccx2_adv;
  a.g.polar = h->d_polar;
  a.g.t2d = h->d_t2d;
  a.g.t2a = h->d_t2a;
  for (int f = 0; f < nfields; ++f) {
    greb_grid_t g = hs[f];
    const ptrdiff_t shift = -(ptrdiff_t)g->kbase * g->nx;
    // level L lives in buffer (cur ^ (L - level)) ; with cur the buffer of the current level:
    // X[L & 1] must be the buffer of level L -> order the two buffers by the parity of the current level
    float* cur = g->d_X[g->cur];
    float* oth = g->d_X[g->cur ^ 1];
    float* by_parity[2];
    by_parity[g->level & 1] = cur;
    by_parity[(g->level & 1) ^ 1] = oth;
    for (int b = 0; b < 2; ++b) {
      a.f[f].X[b] = by_parity[b] + shift;
      a.f[f].Xw[b] = by_parity[b] + shift;
    }
    a.f[f].wz = g->d_wz + shift;
    a.f[f].u = g->d_u + shift;
    a.f[f].v = g->d_v + shift;
    if (g->level_stale) {   // the launch-per-sub-step path moved the band: every stored row is at g->level
      std::vector<int> lv(g->nrows, g->level);
      GCK(cudaMemcpy(g->d_level, lv.data(), lv.size() * sizeof(int), cudaMemcpyHostToDevice));
      g->level_stale = false;
    }
    a.f[f].level = g->d_level - g->kbase;
    for (int sd = 0; sd < 2; ++sd) {
      if (!g->nbX[sd][0]) continue;
      // the neighbour keeps level L in ITS buffer L & 1 as well (same invariant, same number of sub-steps
      // since set_fields)
      const ptrdiff_t nshift = -(ptrdiff_t)g->nb_kbase[sd] * g->nx;
      a.f[f].nbX[sd][0] = g->nbX[sd][0] + nshift;
      a.f[f].nbX[sd][1] = g->nbX[sd][1] + nshift;
      a.f[f].nblevel[sd] = g->nblevel[sd] - g->nb_kbase[sd];
    }
  }
  a.nfields = nfields;
  a.k0 = h->k0;
  a.k1 = h->k1;
  a.level0 = h->level;
  a.nsub = n;
  a.nitems = h->order_items;
  a.order = h->d_order;
  a.counter = h->d_bar;
  a.error = h->d_flags + 4;
  a.timeout = 6000000000LL;   // ~3 s at 2 GHz: a rank that never arrives ends the kernel instead of hanging the GPU
  const size_t smem = (size_t)4 * (h->nx + 6) * sizeof(float);
  int per_sm = 0, sms = 0;
  GCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greb_grid_persistent_kernel, GG_THREADS, smem));
  GCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  int grid = per_sm * sms;
  if ((long long)grid > (long long)a.nitems * n) grid = a.nitems * n;
  if (getenv("GREB_GRID_DEBUG"))
    fprintf(stderr, "greb_grid_run_persistent: %d CTAs/SM x %d SMs, grid %d, %d items per level, %d levels\n", per_sm, sms,
            grid, a.nitems, n);
  if (n == 0) return 0;
  if (grid < 1) return gfail(h, "greb_grid_run_persistent: the kernel does not fit an SM");
  GCK(cudaMemsetAsync(h->d_bar, 0, 4 * sizeof(unsigned), h->stream));
  void* params[] = {&a};
  GCK(cudaEventRecord(h->ev0, h->stream));
  GCK(cudaLaunchCooperativeKernel((void*)greb_grid_persistent_kernel, dim3(grid), dim3(GG_THREADS), params, smem, h->stream));
  GCK(cudaEventRecord(h->ev1, h->stream));
  GCK(cudaStreamSynchronize(h->stream));
  GCK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_launches = 1;
  int err = 0;
  GCK(cudaMemcpy(&err, h->d_flags + 4, sizeof(int), cudaMemcpyDeviceToHost));
  for (int f = 0; f < nfields; ++f) {
    greb_grid_t g = hs[f];
    g->level += n;
    if (n & 1) g->cur ^= 1;
    // own rows are current; the halo rows hold the neighbours' rows of the same level (pushed by them)
    g->valid_lo = g->kbase;
    g->valid_hi = g->kbase + g->nrows;
  }
  if (err) return gfail(h, "greb_grid_run_persistent: a wait for a neighbour (or the grid barrier) timed out");
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
