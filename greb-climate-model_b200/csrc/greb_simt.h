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
#include <string.h>

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
  int warp;     // logical warp index in the CTA (uniform): 0..11 main, 12.. helper
  int lane_u;   // this thread's lane
  float* smem;  // dynamic shared memory base
  int late;     // stagger: this (main) warp does its x part after the barrier
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
// ---- expf / logf of the exact mode: glibc's algorithm, bit for bit -------------------------------------
// The reference is linked against glibc's libm; the CUDA libm differs from it in the last ulp of a few
// per cent of the arguments, and the ice-albedo / To_ice2 switches of the column physics amplify that
// over decades (a perturbed member at 280 ppm drifted by 0.02 K in 50 years).  So the exact mode
// evaluates exp and log the way glibc >= 2.27 does (sysdeps/ieee754/flt-32/e_expf.c, e_logf.c — the ARM
// optimized-routines algorithms): double-precision arithmetic, a 32-entry 2^(i/32) table / a 16-entry
// (1/c, log c) table, a cubic.  Constants are those of glibc 2.39's __exp2f_data / __logf_data.
// tools/glibc_libm_check.c compares this restatement with the host's expf/logf over ALL 2^32 inputs:
// logf identical for every positive finite float, expf identical for every |x| < 88 except x = 0x1.04845ep+5
// and x = -0x1.f8cbb2p+5 (1 ulp; the model's arguments lie in [-20, 5]).  Outside those ranges the CUDA
// libm is used (the special values agree).  B200 executes FP64 at half the FP32 rate: ~10 DP operations
// per call cost less than the CUDA libm's single-precision sequences.
__device__ const unsigned long long greb_exp2f_tab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};
__device__ const double greb_logf_tab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

GDEV float greb_expf_glibc(float x) {
  if (!(fabsf(x) <= 32.0f)) return expf(x);
  const double z = __dmul_rn(0x1.71547652b82fep+5, (double)x);       // x * N/ln2, N = 32
  double kd = __dadd_rn(z, 0x1.8p+52);                               // round to integer (ties to even)
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dsub_rn(kd, 0x1.8p+52);
  const double r = __dsub_rn(z, kd);
  const unsigned long long t = __ldg(&greb_exp2f_tab[ki & 31]) + (ki << 47);
  const double s = __longlong_as_double((long long)t);               // 2^(k/N)
  const double p = __fma_rn(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(0x1.62e42ff0c52d6p-6, r, 1.0);
  y = __fma_rn(p, r2, y);
  return __double2float_rn(__dmul_rn(y, s));
}

GDEV float greb_logf_glibc(float x) {
  unsigned ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix == 0u || ix >= 0x7f800000u) return logf(x);                 // +-0, negative, inf, nan: -inf / nan / inf
    ix = __float_as_uint(__fmul_rn(x, 0x1p23f)) - (23u << 23);         // subnormal: normalise (e_logf.c)
  }
  const unsigned tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int)tmp >> 23;
  const unsigned iz = ix - (tmp & 0xff800000u);
  const double invc = __ldg(&greb_logf_tab[2 * i]), logc = __ldg(&greb_logf_tab[2 * i + 1]);
  const double z = (double)__uint_as_float(iz);
  const double r = __fma_rn(z, invc, -1.0);
  const double y0 = __fma_rn((double)k, 0x1.62e42fefa39efp-1, logc);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
  y = __fma_rn(-0x1.00ea348b88334p-2, r2, y);
  y = __fma_rn(y, r2, __dadd_rn(y0, r));
  return __double2float_rn(y);
}
#ifdef GREB_DBG_FASTMATH   // timing experiment only: approximate transcendentals
GDEV vf v_exp(vf a) { return __expf(a); }
GDEV vf v_log(vf a) { return __logf(a); }
#elif defined(GREB_CUDA_LIBM)   // the CUDA libm (differs from glibc in the last ulp)
GDEV vf v_exp(vf a) { return expf(a); }
GDEV vf v_log(vf a) { return logf(a); }
#else
GDEV vf v_exp(vf a) { return greb_expf_glibc(a); }
GDEV vf v_log(vf a) { return greb_logf_glibc(a); }
#endif
// x / c for a constant c whose double-precision reciprocal rc = RN_f64(1/c) is known: correctly rounded
// for ALL fp32 x and c.  A quotient of two 24-bit significands that is not a float lies at least 2^-49
// (relative) away from every float and every midpoint between floats, and x * rc is within 2^-52 of it, so
// rounding the double product to fp32 rounds the exact quotient.  Three instructions without the FCHK / slow
// path of the IEEE fp32 division sequence.
typedef double vd;
GDEV vd v_seld(vb p, vd a, vd b) { return p ? a : b; }
GDEV vf v_divc(vf x, vf c, vd rc) { (void)c; return __double2float_rn(__dmul_rn((double)x, rc)); }
// approximate forms of the fast arithmetic mode (never used in the exact mode)
GDEV vf v_div_fast(vf a, vf b) { return __fdividef(a, b); }
GDEV vf v_log_fast(vf a) { return __logf(a); }
GDEV vf v_exp_fast(vf a) { return __expf(a); }
GDEV vb v_any_true(vb p) { return __any_sync(0xffffffffu, p); }  // uniform result
// ---- packed pairs of fp32 (Blackwell FADD2 / FMUL2 / FFMA2) ---------------------------------------
// A vf2 holds two independent fp32 values in an aligned 64-bit register pair; add / sub / mul / fma issue
// as ONE instruction for both halves (half the issue slots of the scalar forms; tools/microbench: twice
// the scalar FADD/FMUL rate).  Every operation is the IEEE round-to-nearest operation on each half, so
// packing never changes results — with one caveat: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even with --fmad=false (the scalar forms are left alone; tools/microbench/packed_exact.cu).  A
// multiply and an add whose flush-to-zero modes differ cannot be contracted, so the packed add/sub carry
// .ftz: they differ from IEEE only when an operand or the result is subnormal (< 1.2e-38), which no
// temperature, humidity or increment of the model is (the whole-run bit-identity tests watch over it).
struct vf2 { unsigned long long v; };
GDEV vf2 p_pack(vf lo, vf hi) { vf2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
GDEV vf p_lo(vf2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return lo; }
GDEV vf p_hi(vf2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return hi; }
GDEV vf2 p_bcast(float x) { return p_pack(x, x); }
GDEV vf2 p_add(vf2 a, vf2 b) { vf2 r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
GDEV vf2 p_sub(vf2 a, vf2 b) { vf2 r; asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
GDEV vf2 p_mul(vf2 a, vf2 b) { vf2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
GDEV vf2 p_fma(vf2 a, vf2 b, vf2 c) {
  vf2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
// 4 consecutive floats = 2 packed pairs, 16-byte aligned (LDS.128)
GDEV void p_ld2(vf2 (&o)[2], const float* p, vi idx) {
  const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p + idx);
  o[0].v = t.x; o[1].v = t.y;
}
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
// ---- 1-D bulk asynchronous copy global -> shared (TMA engine: cp.async.bulk, SASS UBLKCP) -------------
// One thread arms the transaction barrier with the byte count and issues the copy; the data lands in shared
// memory without passing through any register, and every consumer thread waits on the barrier's phase
// parity (a hardware sleep, not a poll).  Addresses and sizes are multiples of 16 bytes.
GDEV void tma_bar_init(unsigned long long* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
GDEV void tma_load(float* dst_smem, const float* src_gmem, unsigned bytes, unsigned long long* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(src_gmem), "r"(bytes), "r"(b)
               : "memory");
}
// asks the bulk-copy engine to bring `bytes` at src into L2 (no destination, no completion to wait for)
GDEV void tma_prefetch_l2(const float* src_gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
GDEV void tma_wait(unsigned long long* bar, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(a), "r"((unsigned)(parity & 1))
        : "memory");
  } while (!ok);
}
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
  int late;
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
struct vd { double v[GW]; vd() {} vd(double x) { for (int l = 0; l < GW; ++l) v[l] = x; } };
GDEV vd v_seld(vb p, vd a, vd b) { vd r; for (int l = 0; l < GW; ++l) r.v[l] = p.v[l] ? a.v[l] : b.v[l]; return r; }
GDEV vf v_divc(vf x, vf c, vd) { return v_div(x, c); }   // the emulator divides (it IS the definition)
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

// packed pairs: two lane vectors, every operation applied to each half
struct vf2 { vf lo, hi; };
GDEV vf2 p_pack(vf lo, vf hi) { vf2 r; r.lo = lo; r.hi = hi; return r; }
GDEV vf p_lo(vf2 a) { return a.lo; }
GDEV vf p_hi(vf2 a) { return a.hi; }
GDEV vf2 p_bcast(float x) { return p_pack(vf(x), vf(x)); }
GDEV vf2 p_add(vf2 a, vf2 b) { return p_pack(v_add(a.lo, b.lo), v_add(a.hi, b.hi)); }
GDEV vf2 p_sub(vf2 a, vf2 b) { return p_pack(v_sub(a.lo, b.lo), v_sub(a.hi, b.hi)); }
GDEV vf2 p_mul(vf2 a, vf2 b) { return p_pack(v_mul(a.lo, b.lo), v_mul(a.hi, b.hi)); }
GDEV vf2 p_fma(vf2 a, vf2 b, vf2 c) { return p_pack(v_fma(a.lo, b.lo, c.lo), v_fma(a.hi, b.hi, c.hi)); }
GDEV void p_ld2(vf2 (&o)[2], const float* p, vi idx) {
  vf t[4];
  v_ld4(t, p, idx);
  o[0] = p_pack(t[0], t[1]);
  o[1] = p_pack(t[2], t[3]);
}

// bulk copy stand-ins: the copy happens at issue time; the CTA barriers between issue and use order it
GDEV void tma_bar_init(unsigned long long*) {}
GDEV void tma_load(float* dst, const float* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
GDEV void tma_wait(unsigned long long*, int) {}
GDEV void tma_prefetch_l2(const float*, unsigned) {}

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
GDEV vi operator|(vi a, int b) { vi r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] | b; return r; }
GDEV vb operator==(vi a, int b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] == b; return r; }
GDEV vb operator!=(vi a, int b) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = a.v[l] != b; return r; }
GDEV vb v_bit(vi a, int bit) { vb r; for (int l = 0; l < GW; ++l) r.v[l] = (a.v[l] >> bit) & 1; return r; }
#endif

#if GREB_DEVICE
GDEV vb v_bit(vi a, int bit) { return (a >> bit) & 1; }
#endif
