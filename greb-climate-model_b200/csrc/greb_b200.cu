// greb_b200.cu — CUDA runtime + C ABI of the B200-native GREB stepping core (include/greb_b200.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false (exact mode: the
// reference is compiled without FMA contraction, see greb_core.h).
//
// There is NO CPU path in this file: every compute entry point needs an sm_100 device and fails
// with GREB_E_NO_DEVICE otherwise.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/greb_b200.h"
#include "greb_core.h"
#include "greb_core6.h"
#include "greb_setup.h"

// ------------------------------------------------------------------------------------------------
//                                            kernels
// ------------------------------------------------------------------------------------------------

#define GREB_SMEM_BYTES (GSM_FLOATS * (int)sizeof(float))

__device__ __forceinline__ void cta_prologue(GrebMemberConst* dst, const GrebMemberConst* src, float* smem) {
  const int* s = reinterpret_cast<const int*>(src);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(GrebMemberConst) / sizeof(int)); i += blockDim.x) d[i] = s[i];
  if (threadIdx.x == 0) {
    sb_init(reinterpret_cast<SplitBar*>(smem + GSM_SYNC), GREB_NWARP);  // one arrival per warp
    tma_bar_init(reinterpret_cast<unsigned long long*>(smem + GSM_TMA_BAR_A));
    tma_bar_init(reinterpret_cast<unsigned long long*>(smem + GSM_TMA_BAR_Q));
  }
  __syncthreads();
}

// hardware warp id -> logical warp (0..11 main, 12..13 helper, -1 = empty slot); greb_types.h "Warp placement"
__device__ __forceinline__ int greb_logical_warp(unsigned long long warp_map, int hw) {
  const int w = (int)((warp_map >> (4 * hw)) & 15ull);
  return w == GREB_SLOT_EMPTY ? -1 : w;
}

// One CTA integrates one ensemble member for a.nsteps 12-hour steps (time_loop, f:239-274, or
// qflux_correction, f:325-362, selected by a.spinup).
// SW = 1: the build with the process switches of greb.original.model.f90 (GREB_SW_*), launched only
// when a member of the handle has a switch set, so the default path carries none of that code.
template <int MODE, int SW>
__global__ void __launch_bounds__(GREB_NTHREADS, 1) greb_member_kernel(const GrebKernelArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ GrebMemberConst mc_s;
  const int member = a.member_ids[blockIdx.x];
  cta_prologue(&mc_s, a.mc + member, smem);
  SimtCtx ctx;
  ctx.warp = warp_uniform(greb_logical_warp(a.warp_map, threadIdx.x >> 5));
  if (ctx.warp < 0) return;   // an empty warp slot of the placement (exited threads do not count at barriers)
  ctx.late = (a.late_mask >> ctx.warp) & 1;
  ctx.lane_u = threadIdx.x & 31;
  ctx.smem = smem;
  member_run<MODE, SW>(ctx, a, mc_s, member);
}

// circulation(X_in, dX_crcl, h_scl, wz) (f:528-553): one CTA per field
template <int MODE>
__global__ void __launch_bounds__(GREB_NTHREADS, 1) greb_circulation_kernel(const GrebCirculationArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ GrebMemberConst mc_s;
  cta_prologue(&mc_s, a.mc, smem);
  SimtCtx ctx;
  ctx.warp = warp_uniform(greb_logical_warp(a.warp_map, threadIdx.x >> 5));
  if (ctx.warp < 0) return;   // an empty warp slot of the placement (exited threads do not count at barriers)
  ctx.late = (a.late_mask >> ctx.warp) & 1;
  ctx.lane_u = threadIdx.x & 31;
  ctx.smem = smem;
  SyncState ss;
  ss.bar = reinterpret_cast<SplitBar*>(smem + GSM_SYNC);
  ss.hb = smem + GSM_HB;
  ss.smem = smem;
  ss.phase = 0;
  const size_t off = (size_t)blockIdx.x * GNC;
  if (!ctx_is_helper(ctx)) {
    const RowGeom g = row_geom(ctx, mc_s);
    Tile t;
    {
      const FastRow fr = fast_row(g.k, mc_s);
      tile_load_uv(t, g, a.uv, a.uv + GNC, smem, MODE == 1 ? -fr.cadv : 1.0f, MODE == 1 ? fr.cyA : 1.0f,
                   MODE == 1 ? fr.cyB : 1.0f);
    }
    tile_load_wz(t, g, a.wz + off, smem);
#if GREB_YCOEF
    if (MODE == 1) tile_fold_ycoef(t, g, mc_s.ccy_diff, smem);
#endif
    tile_load_field(t, g, a.X_in + off);
    circulation_main<MODE>(ctx, t, g, mc_s, ss);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int idx = g.k * GX + g.col + 4 * q;
      const float4 x = *reinterpret_cast<const float4*>(a.X_in + off + idx);
      *reinterpret_cast<float4*>(a.dX + off + idx) = make_float4(t.T[4 * q] - x.x, t.T[4 * q + 1] - x.y,
                                                                 t.T[4 * q + 2] - x.z, t.T[4 * q + 3] - x.w);  // f:551
    }
  } else {
    const HelperGeom hg = helper_geom(ctx, mc_s);
    HelperRow hr[GREB_HROWS];
    helper_load_uv(hr, hg, a.uv, a.uv + GNC);
    helper_load_wz(hr, hg, a.wz + off);
    circulation_helper<MODE>(ctx, hr, hg, mc_s, a.X_in + off, ss);
  }
}

// Column physics of one 12-h step on TILES (BASELINE.json configs[4]: a whole step of one member on a grid of
// any size).  SW/LW radiation, sensible heat, hydro, deep ocean, the Ts/To/cap_surf/Ta/q updates and sea ice
// (f:277-308 + f:258-272) are cell-local, so a band of a big grid is cut into tiles of 4,608 consecutive cells
// laid out exactly like an ensemble member ([tile][field][4608]); every tile carries its own slice of the
// step's forcing, of the solar radiation (one value per 96-cell segment: 96 divides xdim) and of the flux
// corrections.  One CTA per tile runs THE SAME device functions as the member kernel (column_phase_a/b/c), so
// at 96x48 a step is bit-identical to it.  Phase 0 = A (before the circulations), 1 = B (after circulation of
// Ta: X = the circulated field), 2 = C (after circulation of q).  The circulations themselves are greb_grid's.
struct GrebTileArgs {
  int phase, ntiles;
  float co2;
  const GrebMemberConst* mc;  // one physics for all tiles (device copy; group = 0, no switches)
  const float* forc;          // [ntiles][GF_COUNT][GNC]
  const float* sw_solar;      // [ntiles][GY]
  const int* mask;            // [ntiles][GNC]
  const float* z_ocean;       // [ntiles][GNC]
  const float* wz;            // [ntiles][2][GNC]
  float* corr;                // [ntiles][GC_COUNT][GNC]
  float* state;               // [ntiles][GS_COUNT][GNC]
  float* acc;                 // [ntiles][GA_COUNT][GNC]
  float* stash;               // [ntiles][2][GNC]
  const float* X;             // [ntiles][GNC] (phases 1, 2)
};

template <int MODE>
__global__ void __launch_bounds__(GREB_NMAIN * 32) greb_tile_phase_kernel(const GrebTileArgs ta) {
  __shared__ GrebMemberConst mc_s;
  {
    const int* s = reinterpret_cast<const int*>(ta.mc);
    int* d = reinterpret_cast<int*>(&mc_s);
    for (int i = threadIdx.x; i < (int)(sizeof(GrebMemberConst) / sizeof(int)); i += blockDim.x) d[i] = s[i];
  }
  __syncthreads();
  const int t = blockIdx.x;
  GrebKernelArgs a;
  a.forc = ta.forc + (size_t)t * GF_COUNT * GNC;
  a.sw_solar = ta.sw_solar + (size_t)t * GY;
  a.mask = ta.mask + (size_t)t * GNC;
  a.z_ocean = ta.z_ocean + (size_t)t * GNC;
  a.toclim = nullptr;
  a.tclim = nullptr;
  a.qclim = nullptr;
  a.wz = ta.wz + (size_t)t * 2 * GNC;
  a.corr = ta.corr + (size_t)t * GC_COUNT * GNC;
  a.state = ta.state + (size_t)t * GS_COUNT * GNC;
  a.acc = ta.acc + (size_t)t * GA_COUNT * GNC;
  a.out = nullptr;
  a.spinup = 0;
  StepInfo si;
  si.ityr = 0;
  si.month_end = 0;
  si.ndm = 1.0f;
  si.out_rec = 0;
  si.co2 = ta.co2;
  si.spinup = 0;
  float* stash = ta.stash + (size_t)t * 2 * GNC;
  const int k = threadIdx.x >> 3, col = 12 * (threadIdx.x & 7);
#pragma unroll 1
  for (int q = 0; q < 3; ++q) {
    const int idx0 = k * GX + col + 4 * q;
    if (ta.phase == 0) {
      column_phase_a<MODE, 0>(a, mc_s, 0, si, k, idx0, stash, a.corr);
    } else {
      const float4 x = *reinterpret_cast<const float4*>(ta.X + (size_t)t * GNC + idx0);
      const float X4[4] = {x.x, x.y, x.z, x.w};
      if (ta.phase == 1) column_phase_b(a, 0, si, idx0, X4, stash);
      else column_phase_c(a, mc_s, 0, si, idx0, X4, stash, a.corr);
    }
  }
}

// Ensemble moments of the monthly-mean fields (SURVEY 8d config-4 output policy): for every element e of a
// year's record block [12][5][GNC], sum and sum of squares over the members of the handle, in float64 and
// in member order (deterministic).  Thread = element; consecutive threads read consecutive addresses of one
// member's block, so every load is a coalesced 128-byte line; 1.1 GB per 1,024 members and year.
__global__ void __launch_bounds__(256) greb_ensemble_moments_kernel(const float* __restrict__ out, int n_members,
                                                                    double* __restrict__ sum, double* __restrict__ sq) {
  const size_t E = (size_t)12 * 5 * GNC;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double s = 0.0, q = 0.0;
  int m = 0;
  for (; m + 4 <= n_members; m += 4) {      // four independent loads in flight per thread
    const float a = __ldg(out + (size_t)m * E + e), b = __ldg(out + (size_t)(m + 1) * E + e);
    const float c = __ldg(out + (size_t)(m + 2) * E + e), d = __ldg(out + (size_t)(m + 3) * E + e);
    s += (double)a; q += (double)a * (double)a;
    s += (double)b; q += (double)b * (double)b;
    s += (double)c; q += (double)c * (double)c;
    s += (double)d; q += (double)d * (double)d;
  }
  for (; m < n_members; ++m) {
    const float a = __ldg(out + (size_t)m * E + e);
    s += (double)a; q += (double)a * (double)a;
  }
  sum[e] = s;
  sq[e] = q;
}

// circulation on 6-cell tiles (greb_core6.h): 24 warps, no helper warps; exact mode
#define T6_SMEM_BYTES (T6_FLOATS * (int)sizeof(float))
__global__ void __launch_bounds__(T6_NTHREADS, 1) greb_circulation6_kernel(const GrebCirculationArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ GrebMemberConst mc_s;
  {
    const int* s = reinterpret_cast<const int*>(a.mc);
    int* d = reinterpret_cast<int*>(&mc_s);
    for (int i = threadIdx.x; i < (int)(sizeof(GrebMemberConst) / sizeof(int)); i += blockDim.x) d[i] = s[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Row6 g = row6(warp, lane, mc_s);
  Tile6 t;
  const size_t off = (size_t)blockIdx.x * GNC;
  t6_load_uv(t, g, a.uv, a.uv + GNC, smem);
  t6_load_wz(t, g, a.wz + off, smem);
  t6_load_field(t, g, a.X_in + off);
  int phase = 0;
  t6_circulation(t, g, mc_s, smem + T6_HB, smem, phase, ((a.late_mask >> warp) & 1u) != 0);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int idx = g.k * GX + g.col + 2 * q;
    const float2 x = *reinterpret_cast<const float2*>(a.X_in + off + idx);
    *reinterpret_cast<float2*>(a.dX + off + idx) = make_float2(t.T[2 * q] - x.x, t.T[2 * q + 1] - x.y);   // f:551
  }
}

// expf / logf of the exact mode on n arguments (parity entry greb_b200_device_libm)
__global__ void greb_libm_kernel(int which, const float* x, float* y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = which == 0 ? v_exp(x[i]) : v_log(x[i]);
}

// ------------------------------------------------------------------------------------------------
//                                            runtime
// ------------------------------------------------------------------------------------------------

// Warp placements (greb_types.h).  map[hw] = logical warp of hardware warp slot hw (sub-partition hw % 4),
// late = stagger mask over logical main warps, order = logical main warps in the order in which they take
// quads of rows, the costlier polar-branch rows first (greb_assign_rows).
struct GrebLayout {
  const char* name;
  int map[GREB_NSLOTS];
  unsigned late;
  int order[GREB_NMAIN];
};
static const GrebLayout greb_layouts[] = {
    // 0: the 14-warp placement of round 1: helpers share sub-partitions 0 and 1 with three main warps each
    {"shared", {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 15}, 0x0f0u, {2, 3, 6, 7, 10, 11, 0, 1, 4, 5, 8, 9}},
    // 1: sub-partition 3 hosts only the two helpers, sub-partitions 0-2 four main warps each
    {"helpers-alone", {0, 1, 2, 12, 3, 4, 5, 13, 6, 7, 8, 15, 9, 10, 11, 15}, 0xe38u, {0, 3, 1, 4, 2, 5, 6, 9, 7, 10, 8, 11}},
    // 2: sub-partition 3 = two helpers + two main warps; 0 = four non-polar main warps; 1, 2 = three main warps
    {"helpers+2", {0, 1, 2, 12, 3, 4, 5, 13, 6, 7, 8, 9, 10, 15, 15, 11}, 0xd38u, {1, 4, 2, 5, 9, 7, 8, 11, 0, 3, 6, 10}},
    // 3: sub-partition 3 = two helpers + one main warp; 0, 1 = four main warps; 2 = three
    {"helpers+1", {0, 1, 2, 12, 3, 4, 5, 13, 6, 7, 8, 9, 10, 11, 15, 15}, 0xc38u, {2, 5, 8, 9, 1, 4, 7, 11, 0, 3, 6, 10}},
};
#define GREB_NLAYOUTS ((int)(sizeof(greb_layouts) / sizeof(greb_layouts[0])))

static unsigned long long pack_map(const int* map) {
  unsigned long long m = 0;
  for (int i = 0; i < GREB_NSLOTS; ++i) m |= (unsigned long long)(map[i] & 15) << (4 * i);
  return m;
}

struct greb_b200_handle_s {
  int device = 0;
  int n_members = 0;
  std::string err;
  bool have_forcing = false, inited = false;
  GrebHostForcing F;
  std::vector<greb_physics_par> phys;
  std::vector<char> have_member;
  std::vector<std::vector<float>> co2;
  std::vector<int> year0;
  std::vector<unsigned> switches;  // GREB_SW_* per member
  std::vector<int> group_of;   // member -> group
  std::vector<int> group_rep;  // group -> representative member
  int co2_stride = 0;
  // device
  float *d_forc = nullptr, *d_sw = nullptr, *d_zoc = nullptr, *d_toclim = nullptr, *d_tclim = nullptr,
        *d_qclim = nullptr, *d_wz = nullptr, *d_corr = nullptr, *d_state = nullptr, *d_acc = nullptr,
        *d_co2 = nullptr, *d_diag = nullptr, *d_coslat = nullptr;
  float* d_out[2] = {nullptr, nullptr};
  int *d_mask = nullptr, *d_flags = nullptr, *d_ids_all = nullptr, *d_ids_rep = nullptr;
  GrebMemberConst* d_mc = nullptr;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_k[2] = {nullptr, nullptr}, ev_c[2] = {nullptr, nullptr};
  int arith = GREB_ARITH_EXACT;
  GrebLayout layout = greb_layouts[0];   // placement in force (choose_layout)
  int layout_built = -1;                 // arithmetic mode the device row tables were built for
  std::vector<GrebMemberConst> mc_host;
  int it_next = 1;   // step counter `it` of the next scenario step
  int last_out = 0;  // d_out buffer holding the last completed year
  float last_ms = 0.f;
  int last_launches = 0;
  // asynchronous run (greb_b200_run_async / greb_b200_wait)
  long long year_seq = 0;              // years launched since greb_b200_init: d_out slot = year_seq & 1
  bool copy_pending[2] = {false, false};  // ev_c[b] recorded and not yet known to be complete
  bool pending = false;                // a run_async has not been waited for
  int pend_years = 0;
  float *pend_gmean = nullptr, *pend_gcos = nullptr;
  double* d_ens = nullptr;             // [2][12][5][GNC] ensemble sum / sum of squares of the last year's records
  float* d_diag_hist = nullptr;        // [hist_years][n_members][2] annual diagnostics of the pending call
  float* h_diag_hist = nullptr;        // pinned mirror
  int hist_years = 0;
};

static std::string g_create_err;

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
      return GREB_E_CUDA;                                                                         \
    }                                                                                             \
  } while (0)

static int finish_pending(greb_b200_t h);
// completes an asynchronous run before an entry point touches device state or the calendar
#define FIN(h)                                  \
  do {                                          \
    const int rc_ = finish_pending(h);          \
    if (rc_ != GREB_OK) return rc_;             \
  } while (0)

static int fail(greb_b200_t h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

extern "C" const char* greb_b200_last_error(greb_b200_t h) { return h ? h->err.c_str() : g_create_err.c_str(); }
extern "C" int greb_b200_n_members(greb_b200_t h) { return h ? h->n_members : GREB_E_INVALID; }

extern "C" int greb_b200_create(greb_b200_t* out, int n_members, int device) {
  if (!out || n_members < 1) {
    g_create_err = "greb_b200_create: bad arguments";
    return GREB_E_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_create_err = std::string("greb_b200_create: no usable CUDA device (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                   "); this library has no CPU fallback";
    return GREB_E_NO_DEVICE;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_err = std::string("greb_b200_create: device '") + prop.name +
                   "' is not sm_100 (the kernels are built for sm_100a only)";
    return GREB_E_NO_DEVICE;
  }
  greb_b200_t h = new greb_b200_handle_s;
  h->device = device;
  h->n_members = n_members;
  h->phys.resize(n_members);
  h->have_member.assign(n_members, 0);
  h->co2.resize(n_members);
  h->year0.assign(n_members, 1940);
  h->switches.assign(n_members, 0u);
  for (auto& p : h->phys) greb_b200_physics_defaults(&p);
  cudaSetDevice(device);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "greb_b200_create: cudaStreamCreate failed";
    delete h;
    return GREB_E_CUDA;
  }
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_c[i], cudaEventDisableTiming);
  }
  cudaFuncSetAttribute(greb_member_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GREB_SMEM_BYTES);
  cudaFuncSetAttribute(greb_member_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GREB_SMEM_BYTES);
  cudaFuncSetAttribute(greb_member_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GREB_SMEM_BYTES);
  cudaFuncSetAttribute(greb_member_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GREB_SMEM_BYTES);
  cudaFuncSetAttribute(greb_circulation_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GREB_SMEM_BYTES);
  cudaFuncSetAttribute(greb_circulation_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GREB_SMEM_BYTES);
  *out = h;
  return GREB_OK;
}

static void free_device(greb_b200_t h) {
  float** fp[] = {&h->d_forc, &h->d_sw,  &h->d_zoc,   &h->d_toclim, &h->d_tclim,  &h->d_qclim,  &h->d_wz,
                  &h->d_corr, &h->d_state, &h->d_acc, &h->d_co2,    &h->d_diag,   &h->d_coslat, &h->d_out[0],
                  &h->d_out[1]};
  for (float** p : fp) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  int** ip[] = {&h->d_mask, &h->d_flags, &h->d_ids_all, &h->d_ids_rep};
  for (int** p : ip) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (h->d_mc) cudaFree(h->d_mc);
  h->d_mc = nullptr;
  if (h->d_ens) cudaFree(h->d_ens);
  h->d_ens = nullptr;
  if (h->d_diag_hist) cudaFree(h->d_diag_hist);
  h->d_diag_hist = nullptr;
  if (h->h_diag_hist) cudaFreeHost(h->h_diag_hist);
  h->h_diag_hist = nullptr;
  h->hist_years = 0;
  h->pending = false;
  h->copy_pending[0] = h->copy_pending[1] = false;
}

extern "C" int greb_b200_destroy(greb_b200_t h) {
  if (!h) return GREB_E_INVALID;
  cudaSetDevice(h->device);
  finish_pending(h);
  cudaDeviceSynchronize();
  free_device(h);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
    if (h->ev_c[i]) cudaEventDestroy(h->ev_c[i]);
  }
  delete h;
  return GREB_OK;
}

// The placement of an arithmetic mode (measured).  GREB_B200_LAYOUT=<index> or
// GREB_B200_LAYOUT_EXACT / _FAST override it for experiments; "m0,..,m15/late/o0,..,o11" gives a custom
// table (map, hexadecimal late mask, row order).
static bool parse_layout(const char* txt, GrebLayout* out) {
  if (!txt || !*txt) return false;
  int id = -1;
  char tail = 0;
  if (sscanf(txt, "%d%c", &id, &tail) == 1) {
    if (id < 0 || id >= GREB_NLAYOUTS) return false;
    *out = greb_layouts[id];
    return true;
  }
  GrebLayout L = greb_layouts[0];
  L.name = "custom";
  int n = 0, pos = 0;
  for (int i = 0; i < GREB_NSLOTS; ++i) {
    if (sscanf(txt + pos, "%d%n", &L.map[i], &n) != 1) return false;
    pos += n;
    if (txt[pos] == ',') ++pos;
  }
  if (txt[pos++] != '/') return false;
  if (sscanf(txt + pos, "%x%n", &L.late, &n) != 1) return false;
  pos += n;
  if (txt[pos++] != '/') return false;
  for (int i = 0; i < GREB_NMAIN; ++i) {
    if (sscanf(txt + pos, "%d%n", &L.order[i], &n) != 1) return false;
    pos += n;
    if (txt[pos] == ',') ++pos;
  }
  // every logical warp exactly once
  int seen[16] = {0}, seen_o[GREB_NMAIN] = {0};
  for (int i = 0; i < GREB_NSLOTS; ++i) {
    if (L.map[i] < 0 || L.map[i] > 15) return false;
    seen[L.map[i]]++;
  }
  for (int w = 0; w < GREB_NWARP; ++w)
    if (seen[w] != 1) return false;
  for (int i = 0; i < GREB_NMAIN; ++i) {
    if (L.order[i] < 0 || L.order[i] >= GREB_NMAIN || seen_o[L.order[i]]++) return false;
  }
  *out = L;
  return true;
}

static GrebLayout choose_layout(int arith) {
  GrebLayout L = greb_layouts[arith == GREB_ARITH_FAST ? 1 : 0];
  GrebLayout tmp;
  if (parse_layout(getenv("GREB_B200_LAYOUT"), &tmp)) L = tmp;
  if (parse_layout(getenv(arith == GREB_ARITH_FAST ? "GREB_B200_LAYOUT_FAST" : "GREB_B200_LAYOUT_EXACT"), &tmp)) L = tmp;
  return L;
}

// (re)builds the lane-group -> row tables of every member for the placement of the current arithmetic mode
static int apply_layout(greb_b200_t h) {
  if (!h->inited || h->layout_built == h->arith) return GREB_OK;
  h->layout = choose_layout(h->arith);
  for (auto& mc : h->mc_host)
    greb_assign_rows(mc.polar, mc.time2_diff, mc.time2_adv, mc.row_of_group, mc.hslot_of_row, mc.helper_row,
                     &mc.n_hslots, h->layout.order);
  CK(cudaMemcpyAsync(h->d_mc, h->mc_host.data(), h->mc_host.size() * sizeof(GrebMemberConst), cudaMemcpyHostToDevice,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->layout_built = h->arith;
  return GREB_OK;
}

extern "C" int greb_b200_set_arithmetic(greb_b200_t h, int mode) {
  if (!h) return GREB_E_INVALID;
  if (mode != GREB_ARITH_EXACT && mode != GREB_ARITH_FAST)
    return fail(h, GREB_E_INVALID, "greb_b200_set_arithmetic: mode must be GREB_ARITH_EXACT or GREB_ARITH_FAST");
  if (h->inited && mode != h->arith) {
    cudaSetDevice(h->device);
    FIN(h);
  }
  h->arith = mode;
  return apply_layout(h);
}

static void launch_member(greb_b200_t h, int grid, GrebKernelArgs a) {
  a.warp_map = pack_map(h->layout.map);
  a.late_mask = h->layout.late;
  bool sw = false;
  for (unsigned m : h->switches) sw = sw || m != 0;
  const bool fast = h->arith == GREB_ARITH_FAST;
  if (!sw && fast) greb_member_kernel<1, 0><<<grid, GREB_NTHREADS, GREB_SMEM_BYTES, h->stream>>>(a);
  else if (!sw) greb_member_kernel<0, 0><<<grid, GREB_NTHREADS, GREB_SMEM_BYTES, h->stream>>>(a);
  else if (fast) greb_member_kernel<1, 1><<<grid, GREB_NTHREADS, GREB_SMEM_BYTES, h->stream>>>(a);
  else greb_member_kernel<0, 1><<<grid, GREB_NTHREADS, GREB_SMEM_BYTES, h->stream>>>(a);
}

extern "C" int greb_b200_set_forcing(greb_b200_t h, const float* z_topo, const float* glacier, const float* sw_solar,
                                     const float* tclim, const float* qclim, const float* swetclim,
                                     const float* uclim, const float* vclim, const float* mldclim,
                                     const float* cldclim) {
  if (!h || !z_topo || !glacier || !sw_solar || !tclim || !qclim || !swetclim || !uclim || !vclim || !mldclim ||
      !cldclim)
    return h ? fail(h, GREB_E_INVALID, "greb_b200_set_forcing: null pointer") : GREB_E_INVALID;
  greb_build_forcing(h->F, z_topo, glacier, sw_solar, tclim, qclim, swetclim, uclim, vclim, mldclim, cldclim);
  h->have_forcing = true;
  h->inited = false;
  return GREB_OK;
}

extern "C" int greb_b200_set_member(greb_b200_t h, int member, const greb_physics_par* p, const float* co2_ppm,
                                    int n_years, int year0) {
  if (!h) return GREB_E_INVALID;
  if (member < 0 || member >= h->n_members || !p || n_years < 0 || (n_years > 0 && !co2_ppm))
    return fail(h, GREB_E_INVALID, "greb_b200_set_member: bad arguments");
  h->phys[member] = *p;
  h->co2[member].assign(co2_ppm, co2_ppm + n_years);
  h->year0[member] = year0;
  h->have_member[member] = 1;
  h->inited = false;
  return GREB_OK;
}

// switches that change the spin-up (everything but the scenario-only SST forcing)
static unsigned spinup_switches(unsigned mask) { return mask & ~(unsigned)GREB_SW_SST_PLUS_1K; }

extern "C" int greb_b200_set_switches(greb_b200_t h, int member, unsigned mask) {
  if (!h) return GREB_E_INVALID;
  if (member < 0 || member >= h->n_members || (mask & ~(unsigned)GREB_SW_ALL))
    return fail(h, GREB_E_INVALID, "greb_b200_set_switches: bad member index or unknown switch bits");
  if (h->inited) {
    // only the scenario-only bit may change once the spin-up groups exist
    if (spinup_switches(mask) != spinup_switches(h->switches[member]))
      return fail(h, GREB_E_INVALID,
                  "greb_b200_set_switches: after greb_b200_init only GREB_SW_SST_PLUS_1K may be toggled");
    cudaSetDevice(h->device);
    FIN(h);
    h->switches[member] = mask;
    h->mc_host[member].switches = (int)mask;
    const int sw = (int)mask;
    CK(cudaMemcpyAsync(reinterpret_cast<char*>(h->d_mc + member) + offsetof(GrebMemberConst, switches), &sw,
                       sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GREB_OK;
  }
  h->switches[member] = mask;
  return GREB_OK;
}

template <class T>
static cudaError_t upload(T** dptr, const std::vector<T>& v) {
  cudaError_t e = cudaMalloc((void**)dptr, v.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" int greb_b200_init(greb_b200_t h) {
  if (!h) return GREB_E_INVALID;
  if (!h->have_forcing) return fail(h, GREB_E_INVALID, "greb_b200_init: set_forcing has not been called");
  cudaSetDevice(h->device);
  finish_pending(h);
  h->inited = false;  // stays false if anything below fails: no entry point may touch half-built device state
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->copy_stream);
  free_device(h);
  const int N = h->n_members;
  // physics groups: members with identical physics_par share wz fields, spin-up and corrections
  h->group_of.assign(N, -1);
  h->group_rep.clear();
  for (int m = 0; m < N; ++m) {
    int g = -1;
    // compare against group representatives (linear scan with a cheap hash of the bytes)
    for (int gi = (int)h->group_rep.size() - 1; gi >= 0; --gi)
      if (greb_physics_equal(h->phys[m], h->phys[h->group_rep[gi]]) &&
          spinup_switches(h->switches[m]) == spinup_switches(h->switches[h->group_rep[gi]])) {
        g = gi;
        break;
      }
    if (g < 0) {
      g = (int)h->group_rep.size();
      h->group_rep.push_back(m);
    }
    h->group_of[m] = g;
  }
  const int G = (int)h->group_rep.size();
  // memory check before allocating the big per-group correction arrays
  size_t need = (size_t)G * GNT * GC_COUNT * GNC * 4 + (size_t)N * (GS_COUNT + GA_COUNT + 2 * 12 * 5) * GNC * 4 +
                (size_t)GNT * (GF_COUNT + 2) * GNC * 4 + (size_t)G * 2 * GNC * 4;
  size_t freeb = 0, totalb = 0;
  cudaMemGetInfo(&freeb, &totalb);
  if (need + (512u << 20) > freeb) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "greb_b200_init: %d members in %d physics groups need %.1f GB of device memory, %.1f GB free; "
             "run the ensemble in smaller batches",
             N, G, need / 1e9, freeb / 1e9);
    return fail(h, GREB_E_NOMEM, buf);
  }
  std::vector<GrebMemberConst> mc(N);
  std::vector<float> state((size_t)N * GS_COUNT * GNC), wz((size_t)G * 2 * GNC);
  for (int g = 0; g < G; ++g) greb_build_wz(&wz[(size_t)g * 2 * GNC], h->F, h->phys[h->group_rep[g]]);
  for (int m = 0; m < N; ++m) {
    if (m == h->group_rep[h->group_of[m]]) {
      const int rc = greb_build_member_const(mc[m], h->phys[m], h->group_of[m]);
      if (rc != 0) {
        char buf[256];
        snprintf(buf, sizeof buf,
                 "greb_b200_init: member %d: kappa = %g / pi = %g need %s, which this kernel does not support", m,
                 h->phys[m].kappa, h->phys[m].pi,
                 rc == -1 ? "more than 2 latitude rows besides the poles with several polar diffusion sub-steps"
                          : "sub-stepped polar advection or a non-polar pole row");
        return fail(h, GREB_E_INVALID, buf);
      }
      greb_build_initial_state(&state[(size_t)m * GS_COUNT * GNC], h->F, mc[m]);
    } else {
      const int r = h->group_rep[h->group_of[m]];
      mc[m] = mc[r];
      memcpy(&state[(size_t)m * GS_COUNT * GNC], &state[(size_t)r * GS_COUNT * GNC], GS_COUNT * GNC * 4);
    }
  }
  for (int m = 0; m < N; ++m) mc[m].switches = (int)h->switches[m];
  h->layout = choose_layout(h->arith);
  for (auto& c : mc)
    greb_assign_rows(c.polar, c.time2_diff, c.time2_adv, c.row_of_group, c.hslot_of_row, c.helper_row, &c.n_hslots,
                     h->layout.order);
  h->layout_built = h->arith;
  h->co2_stride = 1;
  for (int m = 0; m < N; ++m) h->co2_stride = std::max(h->co2_stride, (int)h->co2[m].size());
  std::vector<float> co2((size_t)N * h->co2_stride, 680.f);
  for (int m = 0; m < N; ++m)
    for (size_t y = 0; y < h->co2[m].size(); ++y) co2[(size_t)m * h->co2_stride + y] = h->co2[m][y];
  std::vector<int> ids_all(N);
  for (int m = 0; m < N; ++m) ids_all[m] = m;

  CK(upload(&h->d_forc, h->F.forc));
  CK(upload(&h->d_sw, h->F.sw_solar));
  CK(upload(&h->d_mask, h->F.mask));
  CK(upload(&h->d_zoc, h->F.z_ocean));
  CK(upload(&h->d_toclim, h->F.toclim));
  CK(upload(&h->d_tclim, h->F.tclim));
  CK(upload(&h->d_qclim, h->F.qclim));
  CK(upload(&h->d_coslat, h->F.coslat_w));
  CK(upload(&h->d_wz, wz));
  CK(upload(&h->d_state, state));
  CK(upload(&h->d_co2, co2));
  CK(upload(&h->d_mc, mc));
  h->mc_host = mc;
  CK(upload(&h->d_ids_all, ids_all));
  CK(upload(&h->d_ids_rep, h->group_rep));
  CK(cudaMalloc((void**)&h->d_corr, (size_t)G * GNT * GC_COUNT * GNC * 4));
  CK(cudaMemset(h->d_corr, 0, (size_t)G * GNT * GC_COUNT * GNC * 4));
  CK(cudaMalloc((void**)&h->d_acc, (size_t)N * GA_COUNT * GNC * 4));
  CK(cudaMemset(h->d_acc, 0, (size_t)N * GA_COUNT * GNC * 4));
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc((void**)&h->d_out[i], (size_t)N * 12 * 5 * GNC * 4));
    CK(cudaMemset(h->d_out[i], 0, (size_t)N * 12 * 5 * GNC * 4));
  }
  CK(cudaMalloc((void**)&h->d_diag, (size_t)N * 2 * 4));
  CK(cudaMemset(h->d_diag, 0, (size_t)N * 2 * 4));
  CK(cudaMalloc((void**)&h->d_flags, (size_t)N * 4));
  CK(cudaMemset(h->d_flags, 0, (size_t)N * 4));
  h->it_next = 1;
  h->year_seq = 0;
  h->inited = true;
  return GREB_OK;
}

static GrebKernelArgs base_args(greb_b200_t h) {
  GrebKernelArgs a;
  memset(&a, 0, sizeof a);
  a.mc = h->d_mc;
  a.member_ids = h->d_ids_all;
  a.forc = h->d_forc;
  a.sw_solar = h->d_sw;
  a.mask = h->d_mask;
  a.z_ocean = h->d_zoc;
  a.toclim = h->d_toclim;
  a.tclim = h->d_tclim;
  a.qclim = h->d_qclim;
  a.wz = h->d_wz;
  a.corr = h->d_corr;
  a.state = h->d_state;
  a.acc = h->d_acc;
  a.out = nullptr;
  a.co2 = h->d_co2;
  a.diag = h->d_diag;
  a.coslat_w = h->d_coslat;
  a.flags = h->d_flags;
  a.co2_stride = h->co2_stride;
  a.out_months = 12;
  return a;
}

extern "C" int greb_b200_spinup(greb_b200_t h, int years) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited) return fail(h, GREB_E_INVALID, "greb_b200_spinup: greb_b200_init has not been called");
  if (years < 0) return fail(h, GREB_E_INVALID, "greb_b200_spinup: years < 0");
  cudaSetDevice(h->device);
  FIN(h);
  const int G = (int)h->group_rep.size();
  GrebKernelArgs a = base_args(h);
  a.member_ids = h->d_ids_rep;
  a.spinup = 1;
  h->last_launches = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int y = 0; y < years; ++y) {
    a.it0 = 1 + y * GNT;
    a.nsteps = GNT;
    launch_member(h, G, a);
    h->last_launches++;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaGetLastError());
  // the spin-up end state of a group's representative is the start state of all its members (f:361, f:226)
  for (int m = 0; m < h->n_members; ++m) {
    const int r = h->group_rep[h->group_of[m]];
    if (r != m)
      CK(cudaMemcpyAsync(h->d_state + (size_t)m * GS_COUNT * GNC, h->d_state + (size_t)r * GS_COUNT * GNC,
                         GS_COUNT * GNC * 4, cudaMemcpyDeviceToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  return GREB_OK;
}

extern "C" int greb_b200_reset_scenario(greb_b200_t h) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited) return fail(h, GREB_E_INVALID, "greb_b200_reset_scenario: not initialised");
  cudaSetDevice(h->device);
  FIN(h);
  // f:227: year=year0, mon=1, irec=0, monthly accumulators zero (tsmn is zero after whole years)
  CK(cudaMemsetAsync(h->d_acc, 0, (size_t)h->n_members * GA_COUNT * GNC * 4, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->it_next = 1;
  return GREB_OK;
}

// Completes a greb_b200_run_async: both streams drained, the annual diagnostics scattered from the pinned
// mirror into the caller's arrays.  Every entry point that touches device state calls it first, so a
// forgotten greb_b200_wait cannot race with anything.
static int finish_pending(greb_b200_t h) {
  if (!h->pending) return GREB_OK;
  h->pending = false;
  cudaError_t e1 = cudaStreamSynchronize(h->stream);
  cudaError_t e2 = cudaStreamSynchronize(h->copy_stream);
  h->copy_pending[0] = h->copy_pending[1] = false;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    h->err = std::string("greb_b200_wait: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
    return GREB_E_CUDA;
  }
  cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  const int N = h->n_members, Y = h->pend_years;
  if (h->pend_gmean || h->pend_gcos)
    for (int y = 0; y < Y; ++y)
      for (int m = 0; m < N; ++m) {
        const float* d = h->h_diag_hist + ((size_t)y * N + m) * 2;
        if (h->pend_gmean) h->pend_gmean[(size_t)m * Y + y] = d[0];
        if (h->pend_gcos) h->pend_gcos[(size_t)m * Y + y] = d[1];
      }
  h->pend_gmean = h->pend_gcos = nullptr;
  return GREB_OK;
}

extern "C" int greb_b200_wait(greb_b200_t h) {
  if (!h) return GREB_E_INVALID;
  cudaSetDevice(h->device);
  return finish_pending(h);
}

extern "C" int greb_b200_run_async(greb_b200_t h, int years, float* out, const int* out_members, int n_out,
                                   float* gmean, float* gmean_coslat) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited) return fail(h, GREB_E_INVALID, "greb_b200_run: greb_b200_init has not been called");
  if (years < 0) return fail(h, GREB_E_INVALID, "greb_b200_run: years < 0");
  if ((h->it_next - 1) % GNT != 0) return fail(h, GREB_E_INVALID, "greb_b200_run: not at a year boundary");
  const int N = h->n_members;
  const int y_first = (h->it_next - 1) / GNT;
  if (y_first + years > h->co2_stride)
    return fail(h, GREB_E_INVALID, "greb_b200_run: the CO2 paths given to set_member are shorter than the run");
  if (!out_members) n_out = N;
  // every argument is checked BEFORE the first launch: a failed call leaves the calendar and the device untouched
  if (out_members) {
    if (n_out < 0) return fail(h, GREB_E_INVALID, "greb_b200_run: n_out < 0");
    for (int i = 0; i < n_out; ++i)
      if (out_members[i] < 0 || out_members[i] >= N)
        return fail(h, GREB_E_INVALID, "greb_b200_run: out_members entry out of range");
  }
  cudaSetDevice(h->device);
  // A call still in flight keeps its kernels and copies; only its host-side completion (diagnostics
  // scatter) must happen before the pinned mirror is reused.  The copies of its last years go on
  // overlapping this call's kernels: that is the point of the asynchronous entry.
  const bool want_diag = gmean || gmean_coslat;
  if (h->pending && (want_diag || h->pend_gmean || h->pend_gcos)) {
    const int rc = finish_pending(h);
    if (rc != GREB_OK) return rc;
  }
  const bool chained = h->pending;   // true: the previous async call's copies may still be running
  if (want_diag && years > h->hist_years) {
    if (h->d_diag_hist) cudaFree(h->d_diag_hist);
    if (h->h_diag_hist) cudaFreeHost(h->h_diag_hist);
    h->d_diag_hist = h->h_diag_hist = nullptr;
    h->hist_years = 0;
    CK(cudaMalloc((void**)&h->d_diag_hist, (size_t)years * N * 2 * 4));
    CK(cudaMallocHost((void**)&h->h_diag_hist, (size_t)years * N * 2 * 4));
    h->hist_years = years;
  }
  GrebKernelArgs a = base_args(h);
  a.spinup = 0;
  const size_t year_floats = (size_t)12 * 5 * GNC;
  h->last_launches = 0;
  if (!chained) CK(cudaEventRecord(h->ev0, h->stream));
  for (int y = 0; y < years; ++y) {
    const int b = (int)(h->year_seq & 1);
    a.it0 = h->it_next;
    a.nsteps = GNT;
    a.out = h->d_out[b];
    if (h->copy_pending[b]) CK(cudaStreamWaitEvent(h->stream, h->ev_c[b], 0));  // buffer b free again
    launch_member(h, N, a);
    h->last_launches++;
    CK(cudaGetLastError());
    h->it_next += GNT;
    h->year_seq++;
    h->last_out = b;
    if (want_diag)   // the year's diagnostics stay on the device; ONE pinned copy at the end of the call
      CK(cudaMemcpyAsync(h->d_diag_hist + (size_t)y * N * 2, h->d_diag, (size_t)N * 2 * 4, cudaMemcpyDeviceToDevice,
                         h->stream));
    if (out) {
      CK(cudaEventRecord(h->ev_k[b], h->stream));
      CK(cudaStreamWaitEvent(h->copy_stream, h->ev_k[b], 0));
      if (!out_members) {
        CK(cudaMemcpy2DAsync(out + (size_t)y * year_floats, (size_t)years * year_floats * 4, h->d_out[b],
                             year_floats * 4, year_floats * 4, N, cudaMemcpyDeviceToHost, h->copy_stream));
      } else {
        for (int i = 0; i < n_out; ++i)
          CK(cudaMemcpyAsync(out + ((size_t)i * years + y) * year_floats,
                             h->d_out[b] + (size_t)out_members[i] * year_floats, year_floats * 4,
                             cudaMemcpyDeviceToHost, h->copy_stream));
      }
      CK(cudaEventRecord(h->ev_c[b], h->copy_stream));
      h->copy_pending[b] = true;
    }
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  if (want_diag && years > 0)
    CK(cudaMemcpyAsync(h->h_diag_hist, h->d_diag_hist, (size_t)years * N * 2 * 4, cudaMemcpyDeviceToHost, h->stream));
  h->pending = true;
  h->pend_years = years;
  h->pend_gmean = gmean;
  h->pend_gcos = gmean_coslat;
  return GREB_OK;
}

// The last completed year's records -> host, on the copy stream, ordered behind everything enqueued on the
// compute stream SO FAR (so a small state transfer issued before this call is not stuck behind 1 GB of
// records on the same DMA engine).  Completed by greb_b200_wait.
extern "C" int greb_b200_fetch_monthly_async(greb_b200_t h, float* out, const int* out_members, int n_out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !out) return fail(h, GREB_E_INVALID, "greb_b200_fetch_monthly_async: bad arguments");
  const int N = h->n_members;
  if (!out_members) n_out = N;
  if (out_members)
    for (int i = 0; i < n_out; ++i)
      if (out_members[i] < 0 || out_members[i] >= N)
        return fail(h, GREB_E_INVALID, "greb_b200_fetch_monthly_async: out_members entry out of range");
  cudaSetDevice(h->device);
  const int b = h->last_out;
  const size_t year_floats = (size_t)12 * 5 * GNC;
  CK(cudaEventRecord(h->ev_k[b], h->stream));
  CK(cudaStreamWaitEvent(h->copy_stream, h->ev_k[b], 0));
  if (!out_members) {
    CK(cudaMemcpyAsync(out, h->d_out[b], (size_t)N * year_floats * 4, cudaMemcpyDeviceToHost, h->copy_stream));
  } else {
    for (int i = 0; i < n_out; ++i)
      CK(cudaMemcpyAsync(out + (size_t)i * year_floats, h->d_out[b] + (size_t)out_members[i] * year_floats,
                         year_floats * 4, cudaMemcpyDeviceToHost, h->copy_stream));
  }
  CK(cudaEventRecord(h->ev_c[b], h->copy_stream));
  h->copy_pending[b] = true;
  h->pending = true;   // greb_b200_wait (or any other entry point) drains the copy stream
  return GREB_OK;
}

extern "C" int greb_b200_run(greb_b200_t h, int years, float* out, const int* out_members, int n_out, float* gmean,
                             float* gmean_coslat) {
  if (!h) return GREB_E_INVALID;
  cudaSetDevice(h->device);
  int rc = finish_pending(h);
  if (rc != GREB_OK) return rc;
  rc = greb_b200_run_async(h, years, out, out_members, n_out, gmean, gmean_coslat);
  if (rc != GREB_OK) return rc;
  return finish_pending(h);
}

extern "C" int greb_b200_time_steps(greb_b200_t h, int it0, int nsteps) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited) return fail(h, GREB_E_INVALID, "greb_b200_time_steps: not initialised");
  if (it0 < 1 || nsteps < 1 || nsteps > GNT || (it0 + nsteps - 2) / GNT >= h->co2_stride)
    return fail(h, GREB_E_INVALID, "greb_b200_time_steps: bad step range (1 <= nsteps <= 730, inside the CO2 path)");
  cudaSetDevice(h->device);
  FIN(h);
  GrebKernelArgs a = base_args(h);
  a.spinup = 0;
  a.it0 = it0;
  a.nsteps = nsteps;
  a.out = h->d_out[0];   // at most 12 month ends in 730 steps: they fill slots 0.. of the year buffer
  h->last_out = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  launch_member(h, h->n_members, a);
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_launches = 1;
  h->it_next = it0 + nsteps;
  return GREB_OK;
}

extern "C" int greb_b200_time_loop(greb_b200_t h, int it) { return greb_b200_time_steps(h, it, 1); }

extern "C" int greb_b200_get_state(greb_b200_t h, int member, int which, float* out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || member < 0 || member >= h->n_members || which < 0 || which >= GS_COUNT || !out)
    return fail(h, GREB_E_INVALID, "greb_b200_get_state: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpy(out, h->d_state + ((size_t)member * GS_COUNT + which) * GNC, GNC * 4, cudaMemcpyDeviceToHost));
  return GREB_OK;
}

extern "C" int greb_b200_set_state(greb_b200_t h, int member, int which, const float* in) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || member < 0 || member >= h->n_members || which < 0 || which >= GS_COUNT || !in)
    return fail(h, GREB_E_INVALID, "greb_b200_set_state: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpy(h->d_state + ((size_t)member * GS_COUNT + which) * GNC, in, GNC * 4, cudaMemcpyHostToDevice));
  return GREB_OK;
}

extern "C" int greb_b200_get_states(greb_b200_t h, float* out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !out) return fail(h, GREB_E_INVALID, "greb_b200_get_states: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpyAsync(out, h->d_state, (size_t)h->n_members * GS_COUNT * GNC * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return GREB_OK;
}

extern "C" int greb_b200_set_states(greb_b200_t h, const float* in) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !in) return fail(h, GREB_E_INVALID, "greb_b200_set_states: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpyAsync(h->d_state, in, (size_t)h->n_members * GS_COUNT * GNC * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return GREB_OK;
}

extern "C" int greb_b200_get_fluxcorr(greb_b200_t h, int member, int which, float* out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || member < 0 || member >= h->n_members || which < 0 || which > 2 || !out)
    return fail(h, GREB_E_INVALID, "greb_b200_get_fluxcorr: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  static const int sel[3] = {GC_TF, GC_QF, GC_TOF};  // ABI order: TF, qF, ToF
  const float* base = h->d_corr + (size_t)h->group_of[member] * GNT * GC_COUNT * GNC + (size_t)sel[which] * GNC;
  CK(cudaMemcpy2D(out, GNC * 4, base, (size_t)GC_COUNT * GNC * 4, GNC * 4, GNT, cudaMemcpyDeviceToHost));
  return GREB_OK;
}

extern "C" int greb_b200_set_fluxcorr(greb_b200_t h, int member, int which, const float* in) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || member < 0 || member >= h->n_members || which < 0 || which > 2 || !in)
    return fail(h, GREB_E_INVALID, "greb_b200_set_fluxcorr: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  static const int sel[3] = {GC_TF, GC_QF, GC_TOF};  // ABI order: TF, qF, ToF
  float* base = h->d_corr + (size_t)h->group_of[member] * GNT * GC_COUNT * GNC + (size_t)sel[which] * GNC;
  CK(cudaMemcpy2D(base, (size_t)GC_COUNT * GNC * 4, in, GNC * 4, GNC * 4, GNT, cudaMemcpyHostToDevice));
  return GREB_OK;
}

extern "C" int greb_b200_get_monthly(greb_b200_t h, int member, float* out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || member < 0 || member >= h->n_members || !out)
    return fail(h, GREB_E_INVALID, "greb_b200_get_monthly: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpy(out, h->d_out[h->last_out] + (size_t)member * 12 * 5 * GNC, (size_t)12 * 5 * GNC * 4,
                cudaMemcpyDeviceToHost));
  return GREB_OK;
}

extern "C" int greb_b200_diag_device(greb_b200_t h, const float** dev_ptr, int* n_floats) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !dev_ptr || !n_floats) return fail(h, GREB_E_INVALID, "greb_b200_diag_device: bad arguments");
  *dev_ptr = h->d_diag;
  *n_floats = h->n_members * 2;
  return GREB_OK;
}

extern "C" int greb_b200_get_flags(greb_b200_t h, int* flags) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !flags) return fail(h, GREB_E_INVALID, "greb_b200_get_flags: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpy(flags, h->d_flags, (size_t)h->n_members * 4, cudaMemcpyDeviceToHost));
  return GREB_OK;
}

extern "C" int greb_b200_circulation(greb_b200_t h, int member, int ityr, const float* X_in, const float* wz,
                                     float* dX_crcl, int n) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || member < 0 || member >= h->n_members || ityr < 1 || ityr > GNT || !X_in || !wz || !dX_crcl ||
      n < 1)
    return fail(h, GREB_E_INVALID, "greb_b200_circulation: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  float *dX = nullptr, *dW = nullptr, *dO = nullptr;
  const size_t bytes = (size_t)n * GNC * 4;
  CK(cudaMalloc((void**)&dX, bytes));
  CK(cudaMalloc((void**)&dW, bytes));
  CK(cudaMalloc((void**)&dO, bytes));
  CK(cudaMemcpy(dX, X_in, bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, wz, bytes, cudaMemcpyHostToDevice));
  GrebCirculationArgs a;
  a.mc = h->d_mc + member;
  a.uv = h->d_forc + (size_t)(ityr - 1) * GF_COUNT * GNC + GF_U * GNC;  // GF_U, GF_V are adjacent
  a.X_in = dX;
  a.wz = dW;
  a.dX = dO;
  a.warp_map = pack_map(h->layout.map);
  a.late_mask = h->layout.late;
  const char* t6 = getenv("GREB_B200_TILE6");
  if (t6 && h->arith == GREB_ARITH_EXACT) {   // experiment: 6-cell tiles, 24 warps (greb_core6.h)
    unsigned late = 0x0f0f0fu;                // every other warp of a sub-partition does its x part after the barrier
    if (t6[0] == '0') late = 0;
    sscanf(t6, "%x", &late);
    a.late_mask = late & 0xffffffu;
    cudaFuncSetAttribute(greb_circulation6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T6_SMEM_BYTES);
  }
  CK(cudaEventRecord(h->ev0, h->stream));
  if (t6 && h->arith == GREB_ARITH_EXACT) greb_circulation6_kernel<<<n, T6_NTHREADS, T6_SMEM_BYTES, h->stream>>>(a);
  else if (h->arith == GREB_ARITH_FAST) greb_circulation_kernel<1><<<n, GREB_NTHREADS, GREB_SMEM_BYTES, h->stream>>>(a);
  else greb_circulation_kernel<0><<<n, GREB_NTHREADS, GREB_SMEM_BYTES, h->stream>>>(a);
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_launches = 1;
  CK(cudaMemcpy(dX_crcl, dO, bytes, cudaMemcpyDeviceToHost));
  cudaFree(dX);
  cudaFree(dW);
  cudaFree(dO);
  return GREB_OK;
}

// ---- asynchronous state transfers + compute-stream sync (pipelined host loops, bench.py e2e) -----
extern "C" int greb_b200_set_states_async(greb_b200_t h, const float* in) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !in) return fail(h, GREB_E_INVALID, "greb_b200_set_states_async: bad arguments");
  cudaSetDevice(h->device);
  CK(cudaMemcpyAsync(h->d_state, in, (size_t)h->n_members * GS_COUNT * GNC * 4, cudaMemcpyHostToDevice, h->stream));
  return GREB_OK;
}

extern "C" int greb_b200_get_states_async(greb_b200_t h, float* out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !out) return fail(h, GREB_E_INVALID, "greb_b200_get_states_async: bad arguments");
  cudaSetDevice(h->device);
  CK(cudaMemcpyAsync(out, h->d_state, (size_t)h->n_members * GS_COUNT * GNC * 4, cudaMemcpyDeviceToHost, h->stream));
  return GREB_OK;
}

extern "C" int greb_b200_sync_compute(greb_b200_t h) {
  if (!h) return GREB_E_INVALID;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return GREB_OK;
}

// ---- checkpoint / resume of a scenario (src/greb.f90:226-234 loop state) ------------
// The loop state of the reference is Ts1,Ta1,To1,q1 + cap_surf (get/set_states), the calendar
// (it -> ityr, jday, mon, year, irec are all functions of `it`, src/greb.f90:241-252, 975-985) and the
// accumulators Tmm,Tamm,Tomm,qmm,apmm (:149) + tsmn (:145).
extern "C" int greb_b200_get_calendar(greb_b200_t h, int* it_next) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !it_next) return fail(h, GREB_E_INVALID, "greb_b200_get_calendar: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  *it_next = h->it_next;
  return GREB_OK;
}

extern "C" int greb_b200_set_calendar(greb_b200_t h, int it_next) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || it_next < 1) return fail(h, GREB_E_INVALID, "greb_b200_set_calendar: bad arguments");
  if ((it_next - 1) / GNT > h->co2_stride)
    return fail(h, GREB_E_INVALID, "greb_b200_set_calendar: beyond the CO2 paths given to set_member");
  cudaSetDevice(h->device);
  FIN(h);
  h->it_next = it_next;
  return GREB_OK;
}

extern "C" int greb_b200_get_accumulators(greb_b200_t h, float* out) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !out) return fail(h, GREB_E_INVALID, "greb_b200_get_accumulators: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpyAsync(out, h->d_acc, (size_t)h->n_members * GA_COUNT * GNC * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return GREB_OK;
}

extern "C" int greb_b200_set_accumulators(greb_b200_t h, const float* in) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !in) return fail(h, GREB_E_INVALID, "greb_b200_set_accumulators: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  CK(cudaMemcpyAsync(h->d_acc, in, (size_t)h->n_members * GA_COUNT * GNC * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return GREB_OK;
}


// ---- ensemble moments of the last completed year's monthly means --------------------------------
static int ensemble_moments(greb_b200_t h) {
  const size_t E = (size_t)12 * 5 * GNC;
  if (!h->d_ens) CK(cudaMalloc((void**)&h->d_ens, 2 * E * sizeof(double)));
  greb_ensemble_moments_kernel<<<(unsigned)((E + 255) / 256), 256, 0, h->stream>>>(h->d_out[h->last_out], h->n_members,
                                                                                    h->d_ens, h->d_ens + E);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return GREB_OK;
}

extern "C" int greb_b200_ensemble_moments_device(greb_b200_t h, const double** dev_sum, const double** dev_sumsq,
                                                 int* n_elements) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !dev_sum || !dev_sumsq || !n_elements)
    return fail(h, GREB_E_INVALID, "greb_b200_ensemble_moments_device: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  const int rc = ensemble_moments(h);
  if (rc != GREB_OK) return rc;
  *dev_sum = h->d_ens;
  *dev_sumsq = h->d_ens + (size_t)12 * 5 * GNC;
  *n_elements = 12 * 5 * GNC;
  return GREB_OK;
}

extern "C" int greb_b200_ensemble_moments(greb_b200_t h, double* sum, double* sumsq) {
  if (!h) return GREB_E_INVALID;
  if (!h->inited || !sum || !sumsq) return fail(h, GREB_E_INVALID, "greb_b200_ensemble_moments: bad arguments");
  cudaSetDevice(h->device);
  FIN(h);
  const int rc = ensemble_moments(h);
  if (rc != GREB_OK) return rc;
  const size_t E = (size_t)12 * 5 * GNC;
  CK(cudaMemcpy(sum, h->d_ens, E * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sumsq, h->d_ens + E, E * sizeof(double), cudaMemcpyDeviceToHost));
  return GREB_OK;
}

// ---- column physics of a step on tiles (include/greb_b200.h) ------------------------------------------
extern "C" int greb_b200_tile_phase(int device, int arith, int phase, int ntiles, const greb_physics_par* p, float co2,
                                    const float* forc, const float* sw_solar, const int* mask, const float* z_ocean,
                                    const float* wz, float* corr, float* state, float* acc, float* stash,
                                    const float* X) {
  if (phase < 0 || phase > 2 || ntiles < 1 || !p || !forc || !sw_solar || !mask || !z_ocean || !wz || !corr || !state ||
      !acc || !stash || (phase > 0 && !X) || (arith != GREB_ARITH_EXACT && arith != GREB_ARITH_FAST)) {
    g_create_err = "greb_b200_tile_phase: bad arguments";
    return GREB_E_INVALID;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_err = "greb_b200_tile_phase: no such CUDA device; this library has no CPU fallback";
    return GREB_E_NO_DEVICE;
  }
  GrebMemberConst mc;
  greb_build_member_const(mc, *p, 0);   // the row tables are not used here
  GrebMemberConst* d_mc = nullptr;
  if (cudaMalloc((void**)&d_mc, sizeof mc) != cudaSuccess ||
      cudaMemcpy(d_mc, &mc, sizeof mc, cudaMemcpyHostToDevice) != cudaSuccess) {
    g_create_err = "greb_b200_tile_phase: CUDA allocation failed";
    return GREB_E_CUDA;
  }
  GrebTileArgs ta;
  ta.phase = phase;
  ta.ntiles = ntiles;
  ta.co2 = co2;
  ta.mc = d_mc;
  ta.forc = forc;
  ta.sw_solar = sw_solar;
  ta.mask = mask;
  ta.z_ocean = z_ocean;
  ta.wz = wz;
  ta.corr = corr;
  ta.state = state;
  ta.acc = acc;
  ta.stash = stash;
  ta.X = X;
  // the legacy default stream: ordered with the caller's own work on it (torch's default stream)
  if (arith == GREB_ARITH_FAST) greb_tile_phase_kernel<1><<<ntiles, GREB_NMAIN * 32>>>(ta);
  else greb_tile_phase_kernel<0><<<ntiles, GREB_NMAIN * 32>>>(ta);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(d_mc);
  if (e != cudaSuccess) {
    g_create_err = std::string("greb_b200_tile_phase: ") + cudaGetErrorString(e);
    return GREB_E_CUDA;
  }
  return GREB_OK;
}

extern "C" int greb_b200_device_libm(greb_b200_t h, int which, const float* x, float* y, int n) {
  if (!h) return GREB_E_INVALID;
  if ((which != 0 && which != 1) || !x || !y || n < 1)
    return fail(h, GREB_E_INVALID, "greb_b200_device_libm: bad arguments");
  cudaSetDevice(h->device);
  float *dx = nullptr, *dy = nullptr;
  CK(cudaMalloc((void**)&dx, (size_t)n * 4));
  CK(cudaMalloc((void**)&dy, (size_t)n * 4));
  CK(cudaMemcpy(dx, x, (size_t)n * 4, cudaMemcpyHostToDevice));
  greb_libm_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(which, dx, dy, n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(y, dy, (size_t)n * 4, cudaMemcpyDeviceToHost));
  cudaFree(dx);
  cudaFree(dy);
  return GREB_OK;
}

extern "C" int greb_b200_last_kernel_ms(greb_b200_t h, float* ms, int* launches) {
  if (!h) return GREB_E_INVALID;
  if (ms) *ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  return GREB_OK;
}
