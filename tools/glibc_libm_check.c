/* tools/glibc_libm_check.c — the evidence behind greb_expf_glibc / greb_logf_glibc (csrc/greb_simt.h).
 *
 * Restates glibc's expf and logf (sysdeps/ieee754/flt-32/e_expf.c, e_logf.c; constants of glibc 2.39's
 * __exp2f_data / __logf_data) in plain C, once with every a*b+c fused (what the x86_64 FMA ifunc variant
 * and the device code execute) and once unfused, and compares both with the host's expf / logf over ALL
 * 2^32 float inputs (sampled with a stride when argv[1] is given).
 *
 *   gcc -O2 -ffp-contract=off -mfma -o glibc_libm_check tools/glibc_libm_check.c -lm -lpthread && ./glibc_libm_check
 *
 * glibc 2.39, x86_64: logf 0 mismatches over all positive finite floats (both flavours); expf, |x| < 88:
 * 2 mismatches (x = 0x1.04845ep+5, x = -0x1.f8cbb2p+5, 1 ulp, both flavours) — the device code uses the
 * restatement for |x| <= 32 only. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static const uint64_t T[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};
static const double LT[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
static inline double FMA(double a, double b, double c, int f) { return f ? fma(a, b, c) : a * b + c; }
static float my_expf(float x, int f) {
  const double SHIFT = 0x1.8p+52, InvLn2N = 0x1.71547652b82fep+5, C0 = 0x1.c6af84b912394p-20,
               C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;
  double z = InvLn2N * (double)x, kd = z + SHIFT;
  uint64_t ki;
  memcpy(&ki, &kd, 8);
  kd -= SHIFT;
  const double r = z - kd;
  uint64_t t = T[ki % 32] + (ki << 47);
  double s;
  memcpy(&s, &t, 8);
  const double p = FMA(C0, r, C1, f), r2 = r * r;
  double y = FMA(C2, r, 1.0, f);
  y = FMA(p, r2, y, f);
  return (float)(y * s);
}
static float my_logf(float x, int f) {
  const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2,
               A2 = -0x1.ffffef20a4123p-2;
  uint32_t ix;
  memcpy(&ix, &x, 4);
  if (ix == 0x3f800000) return 0;
  if (ix < 0x00800000) { /* subnormal: normalise */
    const float xs = x * 0x1p23f;
    memcpy(&ix, &xs, 4);
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000;
  const int i = (tmp >> 19) % 16, k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  const double r = FMA((double)zf, LT[i][0], -1.0, f), y0 = FMA((double)k, Ln2, LT[i][1], f), r2 = r * r;
  double y = FMA(A1, r, A2, f);
  y = FMA(A0, r2, y, f);
  return (float)FMA(y, r2, y0 + r, f);
}
#define NT 64
static long bad[4][NT], badmodel[NT];
static uint64_t stride = 1;
static void* work(void* p) {
  const long id = (long)p;
  for (uint64_t u = id * stride; u < (1ull << 32); u += NT * stride) {
    const uint32_t ix = (uint32_t)u;
    float x;
    memcpy(&x, &ix, 4);
    if (x == x && fabsf(x) < 88.0f) {
      const float e = expf(x);
      for (int f = 0; f < 2; f++) {
        const float m = my_expf(x, f);
        if (memcmp(&e, &m, 4)) {
          bad[f][id]++;
          if (fabsf(x) <= 32.0f) badmodel[id]++;
          if (f) printf("expf(%a): host %a restated %a\n", x, e, m);
        }
      }
    }
    if (x == x && x > 0.0f && x < INFINITY) {
      const float e = logf(x);
      for (int f = 0; f < 2; f++) {
        const float m = my_logf(x, f);
        if (memcmp(&e, &m, 4)) bad[2 + f][id]++;
      }
    }
  }
  return 0;
}
int main(int argc, char** argv) {
  if (argc > 1) stride = strtoull(argv[1], 0, 10);
  pthread_t th[NT];
  for (long i = 0; i < NT; i++) pthread_create(&th[i], 0, work, (void*)i);
  for (int i = 0; i < NT; i++) pthread_join(th[i], 0);
  long s[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < NT; i++) {
    for (int j = 0; j < 4; j++) s[j] += bad[j][i];
    s[4] += badmodel[i];
  }
  printf("stride %llu: expf mismatches unfused %ld fused %ld (|x| <= 32: %ld); logf mismatches unfused %ld fused %ld\n",
         (unsigned long long)stride, s[0], s[1], s[4], s[2], s[3]);
  return (s[4] || s[2] || s[3]) ? 1 : 0;
}
