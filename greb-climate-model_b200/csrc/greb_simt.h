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
#ifdef GREB_DBG_FASTMATH   // timing experiment only: approximate transcendentals
GDEV vf v_exp(vf a) { return __expf(a); }
GDEV vf v_log(vf a) { return __logf(a); }
#else
GDEV vf v_exp(vf a) { return expf(a); }
GDEV vf v_log(vf a) { return logf(a); }
#endif
// approximate forms of the fast arithmetic mode (max error 2 ulp; never used in the exact mode)
GDEV vf v_div_fast(vf a, vf b) { return __fdividef(a, b); }
GDEV vf v_log_fast(vf a) { return __logf(a); }
GDEV vf v_exp_fast(vf a) { return __expf(a); }
GDEV vb v_any_true(vb p) { return __any_sync(0xffffffffu, p); }  // uniform result
// 4 consecutive floats, 16-byte aligned (LDG.128 / LDS.128 / STG.128 / STS.128)
GDEV void v_ld4(vf (&o)[4], const float* p, vi idx) {
  const float4 t = *reinterpret_cast<const float4*>(p + idx);
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
GDEV void v_ldg4(vf (&o)[4], const float* p, vi idx) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p + idx));
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
GDEV void v_ldgi4(vi (&o)[4], const int* p, vi idx) {
  const int4 t = __ldg(reinterpret_cast<const int4*>(p + idx));
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
GDEV void v_st4(float* p, vi idx, vf a, vf b, vf c, vf d) {
  *reinterpret_cast<float4*>(p + idx) = make_float4(a, b, c, d);
}

// ---- split-phase CTA barrier (mbarrier): arrive now, wait later --------------------------------
struct SplitBar { unsigned long long mbar; };
GDEV void sb_init(SplitBar* b, int count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(&b->mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
#if !defined(GREB_BAR_MBARRIER)
// Default: ONE hardware CTA barrier per sub-step, placed after the x-direction part (which needs only
// the thread's own row): publish -> [x part] -> BAR.SYNC -> [y part].  A blocked warp costs no issue
// slots, whereas the split-phase mbarrier below is polled (SYNCS + NANOSLEEP + BRA were 18 % of the
// issued instructions); measured +6 % (1,497 -> 1,591 member-years/s, 148 members).  Shared-memory
// stores before the barrier are visible after it; the field is double-buffered, so the next publish
// never overwrites rows that are still being read.
GDEV void sb_arrive(const SimtCtx&, SplitBar*) {}
GDEV void sb_wait(const SimtCtx&, SplitBar*, int) { __syncthreads(); }
#else
// one arrival per warp: the warp's shared-memory stores are ordered before it
GDEV void sb_arrive(const SimtCtx& c, SplitBar* b) {
  __syncwarp();
  if (c.lane_u == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(&b->mbar);
    unsigned long long st;
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(a) : "memory");
    (void)st;
  }
}
GDEV void sb_wait(const SimtCtx&, SplitBar* b, int phase) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(&b->mbar);
  unsigned ok;
  // try_wait with a suspend-time hint: a waiting warp sleeps in hardware instead of burning the
  // issue slots of the warps it is waiting for
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(a), "r"((unsigned)(phase & 1)), "r"(2000u)
        : "memory");
  } while (!ok);
}
#endif
// ---- release/acquire flags in shared memory (helper warp -> owner warps) -------------------------
GDEV void flag_set(const SimtCtx& c, int* f, int v) {
  __syncwarp();
  if (c.lane_u == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(f);
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
  }
}
GDEV void flag_wait(const SimtCtx&, const int* f, int v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(f);
  int cur;
  do {
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(cur) : "r"(a) : "memory");
  } while (cur != v);
}

#else
// ------------------------------------------------------------------------------------------
//                                emulation (tests only)
// ------------------------------------------------------------------------------------------
#define GREB_DEVICE 0
#define GDEV static inline
#define GUNROLL
#define GNOUNROLL
#include <pthread.h>
#include <sched.h>

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
GDEV vf v_div_fast(vf a, vf b) { return v_div(a, b); }   // the emulator has no approximate forms
GDEV vf v_log_fast(vf a) { return v_log(a); }
GDEV vf v_exp_fast(vf a) { return v_exp(a); }
GDEV vf v_shfl(vf x, vi src) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = x.v[src.v[l] & 31]; return r; }
GDEV vf v_sel(vb p, vf a, vf b) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
GDEV vi v_seli(vb p, vi a, vi b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
GDEV vf v_ld(const float* p, vi idx) { vf r; for (int l = 0; l < GW; ++l) r.v[l] = p[idx.v[l]]; return r; }
GDEV vf v_ldg(const float* p, vi idx) { return v_ld(p, idx); }
GDEV vi v_ldgi(const int* p, vi idx) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = p[idx.v[l]]; return r; }
GDEV void v_st(float* p, vi idx, vf v) { for (int l = 0; l < GW; ++l) p[idx.v[l]] = v.v[l]; }
GDEV bool v_any_true(vb p) { bool a = false; for (int l = 0; l < GW; ++l) a = a || p.v[l]; return a; }
GDEV void v_ld4(vf (&o)[4], const float* p, vi idx) {
  for (int q = 0; q < 4; ++q) for (int l = 0; l < GW; ++l) o[q].v[l] = p[idx.v[l] + q];
}
GDEV void v_ldg4(vf (&o)[4], const float* p, vi idx) { v_ld4(o, p, idx); }
GDEV void v_ldgi4(vi (&o)[4], const int* p, vi idx) {
  for (int q = 0; q < 4; ++q) for (int l = 0; l < GW; ++l) o[q].v[l] = p[idx.v[l] + q];
}
GDEV void v_st4(float* p, vi idx, vf a, vf b, vf c, vf d) {
  for (int l = 0; l < GW; ++l) { p[idx.v[l]] = a.v[l]; p[idx.v[l] + 1] = b.v[l]; p[idx.v[l] + 2] = c.v[l]; p[idx.v[l] + 3] = d.v[l]; }
}
// split-phase barrier: counting barrier with generations; every emulated thread group arrives once
struct EmuBar { pthread_mutex_t mu; pthread_cond_t cv; int count, arrived; long gen; };
struct SplitBar { EmuBar* p; };  // pointer-sized so that it fits the slot reserved in shared memory
GDEV void sb_init(SplitBar* b, int count) {
  b->p = new EmuBar;
  pthread_mutex_init(&b->p->mu, nullptr); pthread_cond_init(&b->p->cv, nullptr);
  b->p->count = count; b->p->arrived = 0; b->p->gen = 0;
}
GDEV void sb_destroy(SplitBar* b) { delete b->p; }
GDEV void sb_arrive(const SimtCtx&, SplitBar* b) {
  EmuBar* e = b->p;
  pthread_mutex_lock(&e->mu);
  if (++e->arrived == e->count) { e->arrived = 0; ++e->gen; pthread_cond_broadcast(&e->cv); }
  pthread_mutex_unlock(&e->mu);
}
// waits until phase number `phase` (0-based) is complete, i.e. phase+1 generations have finished
GDEV void sb_wait(const SimtCtx&, SplitBar* b, int phase) {
  EmuBar* e = b->p;
  pthread_mutex_lock(&e->mu);
  while (e->gen < (long)phase + 1) pthread_cond_wait(&e->cv, &e->mu);
  pthread_mutex_unlock(&e->mu);
}
GDEV void flag_set(const SimtCtx&, int* f, int v) { __atomic_store_n(f, v, __ATOMIC_RELEASE); }
GDEV void flag_wait(const SimtCtx&, const int* f, int v) {
  while (__atomic_load_n(f, __ATOMIC_ACQUIRE) != v) sched_yield();
}

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
