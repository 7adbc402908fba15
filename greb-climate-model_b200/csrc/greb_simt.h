// greb_simt.h — lane-vector abstraction the GREB kernels are written against.
//
// Product build (nvcc, sm_100a): vf/vi/vb are plain float/int/bool, every helper is a
// single native instruction (SHFL, FSEL, LDS, ...) and nothing of the emulation exists.
//
// Test build (-DGREB_EMU, g++ only, used by tests/emu): vf/vi/vb are 32-wide arrays and
// the helpers loop over lanes, so the *same* warp-level kernel source runs lane-exact on the
// CPU where it can be compared bit-for-bit with the oracle without a GPU.  The emulation is
// test infrastructure; the shipped library never contains it and has no CPU path.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(GREB_EMU)
// ------------------------------------------------------------------------------------------
//                                   device (product)
// ------------------------------------------------------------------------------------------
#define GREB_DEVICE 1
#define GDEV __device__ __forceinline__
#define GUNROLL _Pragma("unroll")
#define GNOUNROLL _Pragma("unroll 1")

typedef float vf;
typedef int vi;
typedef bool vb;

struct SimtCtx {
  int warp;     // warp index in the CTA (uniform)
  int lane_u;   // this thread's lane
  float* smem;  // dynamic shared memory base
};

GDEV vi ctx_lane(const SimtCtx& c) { return c.lane_u; }
GDEV void cta_sync(const SimtCtx&) { __syncthreads(); }
GDEV void warp_sync(const SimtCtx&) { __syncwarp(); }
GDEV bool lane0(const SimtCtx& c) { return c.lane_u == 0; }
// tells the compiler a value is warp-uniform (so branches on it need no divergence handling)
GDEV int warp_uniform(int x) { return __shfl_sync(0xffffffffu, x, 0); }

GDEV vf v_bcast(float x) { return x; }
GDEV vf v_shfl(vf x, vi src) { return __shfl_sync(0xffffffffu, x, src); }
GDEV vf v_sel(vb p, vf a, vf b) { return p ? a : b; }
GDEV vi v_seli(vb p, vi a, vi b) { return p ? a : b; }
GDEV vf v_ld(const float* p, vi idx) { return p[idx]; }
GDEV vf v_ldg(const float* p, vi idx) { return __ldg(p + idx); }
GDEV vi v_ldgi(const int* p, vi idx) { return __ldg(p + idx); }
GDEV void v_st(float* p, vi idx, vf v) { p[idx] = v; }
GDEV vf v_fma(vf a, vf b, vf c) { return __fmaf_rn(a, b, c); }
GDEV vf v_add(vf a, vf b) { return __fadd_rn(a, b); }  // never contracted into an FMA
GDEV vf v_sub(vf a, vf b) { return __fsub_rn(a, b); }
GDEV vf v_mul(vf a, vf b) { return __fmul_rn(a, b); }
GDEV vf v_div(vf a, vf b) { return __fdiv_rn(a, b); }
GDEV vf v_abs(vf a) { return fabsf(a); }
GDEV vf v_max(vf a, vf b) { return fmaxf(a, b); }
GDEV vf v_exp(vf a) { return expf(a); }
GDEV vf v_log(vf a) { return logf(a); }
GDEV vb v_any_true(vb p) { return __any_sync(0xffffffffu, p); }  // uniform result

#else
// ------------------------------------------------------------------------------------------
//                                emulation (tests only)
// ------------------------------------------------------------------------------------------
#define GREB_DEVICE 0
#define GDEV static inline
#define GUNROLL
#define GNOUNROLL
#include <pthread.h>

#define GW 32
struct vb {
  bool v[GW];
};
struct vi {
  int v[GW];
  vi() {}
  vi(int x) { for (int l = 0; l < GW; ++l) v[l] = x; }
};
struct vf {
  float v[GW];
  vf() {}
  vf(float x) { for (int l = 0; l < GW; ++l) v[l] = x; }
};

struct SimtCtx {
  int warp;
  vi lane_v;
  float* smem;
  pthread_barrier_t* bar;
};
GDEV vi ctx_lane(const SimtCtx& c) { return c.lane_v; }
GDEV void cta_sync(const SimtCtx& c) { pthread_barrier_wait(c.bar); }
GDEV void warp_sync(const SimtCtx&) {}
GDEV bool lane0(const SimtCtx&) { return true; }  // "one elected lane" work runs once per warp
GDEV int warp_uniform(int x) { return x; }

#define VOP2(name, expr)                                  \
  GDEV vf name(vf a, vf b) {                              \
    vf r;                                                 \
    for (int l = 0; l < GW; ++l) { float x = a.v[l], y = b.v[l]; r.v[l] = (expr); } \
    return r;                                             \
  }
VOP2(v_add, x + y)
VOP2(v_sub, x - y)
VOP2(v_mul, x* y)
VOP2(v_div, x / y)
VOP2(v_max, fmaxf(x, y))
#undef VOP2
GDEV vf v_fma(vf a, vf b, vf c) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = fmaf(a.v[l], b.v[l], c.v[l]); return r; }
GDEV vf v_abs(vf a) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = fabsf(a.v[l]); return r; }
GDEV vf v_exp(vf a) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = expf(a.v[l]); return r; }
GDEV vf v_log(vf a) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = logf(a.v[l]); return r; }
GDEV vf v_bcast(float x) { return vf(x); }
GDEV vf v_shfl(vf x, vi src) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = x.v[src.v[l] & 31]; return r; }
GDEV vf v_sel(vb p, vf a, vf b) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
GDEV vi v_seli(vb p, vi a, vi b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
GDEV vf v_ld(const float* p, vi idx) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = p[idx.v[l]]; return r; }
GDEV vf v_ldg(const float* p, vi idx) { return v_ld(p, idx); }
GDEV vi v_ldgi(const int* p, vi idx) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = p[idx.v[l]]; return r; }
GDEV void v_st(float* p, vi idx, vf v) { for (int l = 0; l < GW; ++l) p[idx.v[l]] = v.v[l]; }
GDEV bool v_any_true(vb p) { bool a = false; for (int l = 0; l < GW; ++l) a = a || p.v[l]; return a; }

// operators so that expression code reads the same in both builds
GDEV vf operator+(vf a, vf b) { return v_add(a, b); }
GDEV vf operator-(vf a, vf b) { return v_sub(a, b); }
GDEV vf operator*(vf a, vf b) { return v_mul(a, b); }
GDEV vf operator/(vf a, vf b) { return v_div(a, b); }
GDEV vf operator+(vf a, float b) { return v_add(a, vf(b)); }
GDEV vf operator-(vf a, float b) { return v_sub(a, vf(b)); }
GDEV vf operator*(vf a, float b) { return v_mul(a, vf(b)); }
GDEV vf operator/(vf a, float b) { return v_div(a, vf(b)); }
GDEV vf operator+(float a, vf b) { return v_add(vf(a), b); }
GDEV vf operator-(float a, vf b) { return v_sub(vf(a), b); }
GDEV vf operator*(float a, vf b) { return v_mul(vf(a), b); }
GDEV vf operator/(float a, vf b) { return v_div(vf(a), b); }
GDEV vf operator-(vf a) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = -a.v[l]; return r; }
#define VCMP(op)                                                                                  \
  GDEV vb operator op(vf a, vf b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] op b.v[l]; return r; } \
  GDEV vb operator op(vf a, float b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] op b; return r; }
VCMP(<) VCMP(<=) VCMP(>) VCMP(>=) VCMP(==)
#undef VCMP
GDEV vb operator&&(vb a, vb b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] && b.v[l]; return r; }
GDEV vb operator||(vb a, vb b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] || b.v[l]; return r; }
GDEV vb operator!(vb a) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = !a.v[l]; return r; }
GDEV vb operator&&(vb a, bool b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] && b; return r; }
GDEV vi operator+(vi a, vi b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] + b.v[l]; return r; }
GDEV vi operator+(vi a, int b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] + b; return r; }
GDEV vi operator+(int a, vi b) { return b + a; }
GDEV vi operator-(vi a, int b) { return a + (-b); }
GDEV vi operator*(vi a, int b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] * b; return r; }
GDEV vi operator*(int a, vi b) { return b * a; }
GDEV vi operator&(vi a, int b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] & b; return r; }
GDEV vb operator==(vi a, int b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] == b; return r; }
GDEV vb operator!=(vi a, int b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] != b; return r; }
GDEV vb v_bit(vi a, int bit) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = (a.v[l] >> bit) & 1; return r; }
#endif

#if GREB_DEVICE
GDEV vb v_bit(vi a, int bit) { return (a >> bit) & 1; }
#endif
