// imc_math.h — deterministic elementary functions for host and device.
//
// The tracking loop decides events by exact floating-point comparisons (SURVEY.md §8a,
// steps 3-9), so the CUDA kernels and the CPU oracle must agree bit-for-bit on exp, expm1,
// log, sin/cos and atan.  libm (glibc) and libdevice do not agree in the last ulp, so every
// function here is built only from IEEE add/mul/div/sqrt/fma and integer bit operations,
// which round identically on x86-64 and sm_100a when compiled without contraction
// (nvcc -fmad=false, gcc -ffp-contract=off).  All polynomial steps use explicit fma().
//
// Accuracy (checked against glibc/long double in tests/test_num_math.py): < 1 ulp for exp,
// expm1, log, sin, cos on the ranges the transport uses; <= 2 ulp for atan2.  That is the same
// class as Julia's pure-Julia Base implementations (not correctly rounded either), which the
// reference calls (imc_transport.jl:95,110,534,628) — see DESIGN.md "parity unpinned".
//
// Algorithms are the textbook ones: Cody-Waite additive range reduction, Taylor/atanh series
// in Horner form with enough terms for the target precision, exact power-of-two scaling.
#pragma once
#include "imc_num.h"

namespace imc {
namespace dm {

IMC_HD double fma_d(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return fma(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
IMC_HD float fma_f(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return fmaf(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}
IMC_HD double rint_d(double x) {
#if defined(__CUDA_ARCH__)
  return rint(x);
#else
  return __builtin_rint(x);
#endif
}
IMC_HD float rint_f(float x) {
#if defined(__CUDA_ARCH__)
  return rintf(x);
#else
  return __builtin_rintf(x);
#endif
}
IMC_HD double sqrt_d(double x) {
#if defined(__CUDA_ARCH__)
  return sqrt(x);
#else
  return __builtin_sqrt(x);
#endif
}
IMC_HD float sqrt_f(float x) {
#if defined(__CUDA_ARCH__)
  return sqrtf(x);
#else
  return __builtin_sqrtf(x);
#endif
}

IMC_HD uint64_t d_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
IMC_HD double bits_d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
IMC_HD uint32_t f_bits(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
IMC_HD float bits_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

// 2^n as a double / float, n inside the normal exponent range.
IMC_HD double pow2_d(int n) { return bits_d((uint64_t)(n + 1023) << 52); }
IMC_HD float pow2_f(int n) { return bits_f((uint32_t)(n + 127) << 23); }
// p * 2^n with one rounding at most (p in [0.5, 2), |n| <= 1100): split so both factors are normal.
IMC_HD double scale_d(double p, int n) {
  int n1 = n / 2, n2 = n - n1;
  return (p * pow2_d(n1)) * pow2_d(n2);
}
IMC_HD float scale_f(float p, int n) {
  int n1 = n / 2, n2 = n - n1;
  return (p * pow2_f(n1)) * pow2_f(n2);
}

// Coefficient tables of the double-precision polynomials.  On sm_100a a 64-bit literal operand costs two UMOVs at every use
// (DFMA takes no 64-bit immediate and no constant-bank operand), i.e. a Horner step is three issue slots instead of one;
// from a __constant__ table two coefficients arrive per LDCU.128.  The host side (and the CPU oracle) reads the same
// values from a plain array: the initialisers are the same constant expressions, so both sides hold the same bits.
#if defined(__CUDACC__)
#define IMC_DTAB(name, n, ...) static __constant__ double name##_c[n] = {__VA_ARGS__}; static const double name##_h[n] = {__VA_ARGS__};
#else
#define IMC_DTAB(name, n, ...) static const double name##_h[n] = {__VA_ARGS__};
#endif
#if defined(__CUDA_ARCH__)
#define IMC_DT(name) name##_c
#else
#define IMC_DT(name) name##_h
#endif
// p = t[0]; p = fma(p, x, t[i]) for i = 1 .. N-1
#define IMC_HORNER_D(p, x, name, N) do { p = IMC_DT(name)[0]; _Pragma("unroll") for (int i_ = 1; i_ < (N); ++i_) p = fma_d(p, x, IMC_DT(name)[i_]); } while (0)

// =======================================================================================
// double precision
// =======================================================================================
IMC_DTAB(EXPQ, 12, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5)
IMC_DTAB(LOGP, 12, 1.0 / 25.0, 1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0)
IMC_DTAB(SINP, 9, -1.0 / 121645100408832000.0, 1.0 / 355687428096000.0, -1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0)
IMC_DTAB(COSP, 9, 1.0 / 2432902008176640000.0, -1.0 / 6402373705728000.0, 1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0)
IMC_DTAB(ATANP, 22, 1.0 / 45.0, -1.0 / 43.0, 1.0 / 41.0, -1.0 / 39.0, 1.0 / 37.0, -1.0 / 35.0, 1.0 / 33.0, -1.0 / 31.0, 1.0 / 29.0, -1.0 / 27.0, 1.0 / 25.0, -1.0 / 23.0, 1.0 / 21.0, -1.0 / 19.0, 1.0 / 17.0, -1.0 / 15.0, 1.0 / 13.0, -1.0 / 11.0, 1.0 / 9.0, -1.0 / 7.0, 1.0 / 5.0, -1.0 / 3.0)

// scalar constants of the range reductions, through the same tables (DK: exp / log, DS: sin / cos)
IMC_DTAB(DK, 4, 1.44269504088896338700e+00, 6.93147180369123816490e-01 /* 0x3fe62e42fee00000: ln2 to 32 bits */, 1.90821492927058770002e-10, 0.0)
#define IMC_LOG2E_D IMC_DT(DK)[0]
#define IMC_LN2_HI_D IMC_DT(DK)[1]
#define IMC_LN2_LO_D IMC_DT(DK)[2]

// Shared core of exp/expm1: x = n ln2 + r, |r| <= 0.3466; returns em = expm1(r) (~0.5 ulp).
IMC_HD double exp_core_d(double x, int* n_out) {
  double fn = rint_d(x * IMC_LOG2E_D);
  double hi = fma_d(fn, -IMC_LN2_HI_D, x);  // exact
  double lo = fn * IMC_LN2_LO_D;
  double r = hi - lo;
  double c = (hi - r) - lo;  // rounding error of r
  // q = 1/2 + r/6 + ... + r^11/13!
  double q; IMC_HORNER_D(q, r, EXPQ, 12);
  double em = fma_d(r * r, q, r);
  em = fma_d(c, em, em + c);  // first-order correction: expm1(r + c) ~ em + c (1 + em)
  *n_out = (int)fn;
  return em;
}

IMC_HD double exp_d(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return bits_d(0x7ff0000000000000ull);
  if (x < -745.1332191019412) return 0.0;
  int n;
  double em = exp_core_d(x, &n);
  return scale_d(1.0 + em, n);
}

IMC_HD double expm1_from_core_d(double em, int n) {
  if (n == 0) return em;
  if (n > 53) {
    if (n > 1023) return scale_d(1.0 + em, n);
    double t = pow2_d(n);
    return fma_d(t, em, t) - 1.0;
  }
  if (n < -53) return -1.0;
  double t = pow2_d(n);
  return fma_d(t, em, t - 1.0);  // 2^n em + (2^n - 1); (2^n - 1) exact for |n| <= 53
}

IMC_HD double expm1_d(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return bits_d(0x7ff0000000000000ull);
  if (x < -37.5) return -1.0;
  double ax = x < 0 ? -x : x;
  if (ax < 5.551115123125783e-17) return x;  // 2^-54
  int n;
  double em = exp_core_d(x, &n);
  return expm1_from_core_d(em, n);
}

// exp(x) and expm1(x) from one range reduction; bit-identical to the separate calls.
IMC_HD void exp_expm1_d(double x, double* e, double* em1) {
  if (!(x <= 709.0 && x >= -37.0)) { *e = exp_d(x); *em1 = expm1_d(x); return; }
  double ax = x < 0 ? -x : x;
  if (rint_d(x * IMC_LOG2E_D) == 0.0) {  // n = 0: r = x, c = 0, scale 2^0 (see exp_expm1_f)
    double q; IMC_HORNER_D(q, x, EXPQ, 12);
    double em0 = fma_d(x * x, q, x);
    *e = 1.0 + em0;
    *em1 = (ax < 5.551115123125783e-17) ? x : em0;
    return;
  }
  int n;
  double em = exp_core_d(x, &n);
  *e = scale_d(1.0 + em, n);
  *em1 = (ax < 5.551115123125783e-17) ? x : expm1_from_core_d(em, n);
}

IMC_HD double log_d(double x) {
  uint64_t ux = d_bits(x);
  int e = 0;
  if (ux >= 0x7ff0000000000000ull || ux < 0x0010000000000000ull) {  // negative, nan, inf, zero, subnormal
    if ((ux << 1) == 0) return -bits_d(0x7ff0000000000000ull);      // log(+-0) = -inf
    if (ux >> 63) return bits_d(0x7ff8000000000000ull);             // log(<0) = nan
    if (ux >= 0x7ff0000000000000ull) return x;                      // inf / nan
    x *= 18014398509481984.0;                                       // 2^54
    ux = d_bits(x);
    e = -54;
  }
  // normalise the mantissa to [sqrt(1/2), sqrt(2))
  uint32_t hx = (uint32_t)(ux >> 32);
  hx += 0x3ff00000u - 0x3fe6a09eu;
  e += (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
  double m = bits_d(((uint64_t)hx << 32) | (ux & 0xffffffffull));
  double f = m - 1.0;  // exact
  double s = f / (2.0 + f);
  double z = s * s;
  // R = 2 z (1/3 + z/5 + z^2/7 + ... ) : log(m) = 2 s + s R
  double p; IMC_HORNER_D(p, z, LOGP, 12);
  double R = 2.0 * (z * p);
  double hfsq = 0.5 * f * f;
  double dk = (double)e;
  // log(m) = f - hfsq + s (hfsq + R)   (since 2s = f - s f)
  return fma_d(s, hfsq + R, dk * IMC_LN2_LO_D) - hfsq + f + dk * IMC_LN2_HI_D;
}

IMC_DTAB(DS, 6, 6.36619772367581382433e-01 /* 2/pi */, 1.57079632673412561417e+00 /* first 33 bits of pi/2 */, 6.07710050630396597660e-11 /* next 33 bits */,
         2.02226624871116645580e-21 /* next 33 bits */, 8.47842766036889956997e-32 /* tail */, 0.0)
#define IMC_2OPI_D IMC_DT(DS)[0]
#define IMC_PIO2_1_D IMC_DT(DS)[1]
#define IMC_PIO2_2_D IMC_DT(DS)[2]
#define IMC_PIO2_3_D IMC_DT(DS)[3]
#define IMC_PIO2_4_D IMC_DT(DS)[4]

IMC_HD double sin_kernel_d(double r) {
  double z = r * r;
  double p; IMC_HORNER_D(p, z, SINP, 9);
  return fma_d(r * z, p, r);
}
IMC_HD double cos_kernel_d(double r) {
  double z = r * r;
  double p; IMC_HORNER_D(p, z, COSP, 9);
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * z) * p);
}
// valid for |x| <~ 1e5 (transport angles are within a few multiples of pi)
IMC_HD void sincos_d(double x, double* s, double* c) {
  if (!(x - x == 0.0)) { *s = x - x; *c = x - x; return; }  // inf / nan -> nan
  double fq = rint_d(x * IMC_2OPI_D);
  double t = fma_d(fq, -IMC_PIO2_1_D, x);  // exact: PIO2_1 has 33 significant bits
  double w = fq * IMC_PIO2_2_D;            // exact for |fq| < 2^20
  double r = t - w;
  double lo = (t - r) - w;                 // r + lo == t - w
  lo = fma_d(fq, -IMC_PIO2_3_D, lo);
  lo = fma_d(fq, -IMC_PIO2_4_D, lo);
  int q = (int)fq & 3;
  double sr = sin_kernel_d(r), cr = cos_kernel_d(r);
  // first-order correction for the reduction tail: sin(r+lo) ~ sin r + lo cos r, cos(r+lo) ~ cos r - lo sin r
  sr = fma_d(lo, fma_d(-0.5 * r, r, 1.0), sr);
  cr = fma_d(-lo, r, cr);
  double ss = (q & 1) ? cr : sr;
  double cc = (q & 1) ? sr : cr;
  if (q == 1 || q == 2) cc = -cc;
  if (q >= 2) ss = -ss;
  *s = ss;
  *c = cc;
}

#define IMC_PI_HI_D 3.14159265358979311600e+00
#define IMC_PI_LO_D 1.22464679914735317723e-16
IMC_DTAB(DH, 2, 1.57079632679489655800e+00, 6.12323399573676603587e-17)   // pi/2 = hi + lo
#define IMC_PIO2_HI_D IMC_DT(DH)[0]
#define IMC_PIO2_LO_D IMC_DT(DH)[1]
#define IMC_PIO4_HI_D 7.85398163397448278999e-01
#define IMC_PIO4_LO_D 3.06161699786838301793e-17

// atan on [0, 1] -> hi + lo parts added by the caller
IMC_HD double atan01_d(double t) {
  double base_hi = 0.0, base_lo = 0.0, u = t;
  if (t > 0.41421356237309503) {  // tan(pi/8)
    u = (t - 1.0) / (t + 1.0);
    base_hi = IMC_PIO4_HI_D;
    base_lo = IMC_PIO4_LO_D;
  }
  double z = u * u;
  double p; IMC_HORNER_D(p, z, ATANP, 22);
  double a = fma_d(u * z, p, u);  // atan(u)
  return base_hi + (a + base_lo);
}

// Julia's atan(y, x) (= atan2)
IMC_HD double atan2_d(double y, double x) {
  if (x != x || y != y) return x + y;
  uint64_t sy = d_bits(y) >> 63, sx = d_bits(x) >> 63;
  double ay = y < 0 ? -y : y, ax = x < 0 ? -x : x;
  if (sy && ay == 0) ay = 0.0;
  double r;
  const double inf = bits_d(0x7ff0000000000000ull);
  if (ay == 0.0) {
    r = sx ? IMC_PI_HI_D : 0.0;
  } else if (ax == 0.0) {
    r = IMC_PIO2_HI_D;
  } else if (ax == inf) {
    r = (ay == inf) ? (sx ? 3.0 * IMC_PIO4_HI_D : IMC_PIO4_HI_D) : (sx ? IMC_PI_HI_D : 0.0);
  } else if (ay == inf) {
    r = IMC_PIO2_HI_D;
  } else {
    bool swap = ay > ax;
    double t = swap ? ax / ay : ay / ax;
    double a = atan01_d(t);
    if (swap) a = IMC_PIO2_HI_D - (a - IMC_PIO2_LO_D);
    if (sx) a = IMC_PI_HI_D - (a - IMC_PI_LO_D);
    r = a;
  }
  return sy ? -r : r;
}

// x^y for the material power laws (imc_update.jl:31,35) and matenergydens^(1/4) (imc_tally.jl:72).
// Exact special cases; general case exp(y log x) (error grows with |y log x|; documented).
IMC_HD double pow_d(double x, double y) {
  if (y == 0.0) return 1.0;
  if (y == 1.0) return x;
  if (x != x || y != y) return x + y;
  if (y == 0.25) return x < 0 ? bits_d(0x7ff8000000000000ull) : sqrt_d(sqrt_d(x));
  if (y == 0.5) return x < 0 ? bits_d(0x7ff8000000000000ull) : sqrt_d(x);
  double ay = y < 0 ? -y : y;
  if (ay <= 8.0 && ay == rint_d(ay)) {
    int k = (int)ay;
    double r = x;
    for (int i = 1; i < k; ++i) r *= x;
    return y < 0 ? 1.0 / r : r;
  }
  if (x < 0) return bits_d(0x7ff8000000000000ull);
  if (x == 0) return y > 0 ? 0.0 : bits_d(0x7ff0000000000000ull);
  return exp_d(y * log_d(x));
}

// =======================================================================================
// single precision (native float arithmetic; Float16 ops use these and round, as Julia does)
// =======================================================================================
#define IMC_LN2_HI_F 6.93145751953125e-01f  /* 0x3f317200 */
#define IMC_LN2_LO_F 1.42860676533018712e-06f
#define IMC_LOG2E_F 1.44269502162933349609f

IMC_HD float exp_core_f(float x, int* n_out) {
  float fn = rint_f(x * IMC_LOG2E_F);
  float hi = fma_f(fn, -IMC_LN2_HI_F, x);  // exact for |fn| < 2^9
  float lo = fn * IMC_LN2_LO_F;
  float r = hi - lo;
  float c = (hi - r) - lo;
  float q = 1.0f / 40320.0f;
  q = fma_f(q, r, 1.0f / 5040.0f);
  q = fma_f(q, r, 1.0f / 720.0f);
  q = fma_f(q, r, 1.0f / 120.0f);
  q = fma_f(q, r, 1.0f / 24.0f);
  q = fma_f(q, r, 1.0f / 6.0f);
  q = fma_f(q, r, 0.5f);
  float em = fma_f(r * r, q, r);
  em = fma_f(c, em, em + c);
  *n_out = (int)fn;
  return em;
}
IMC_HD float exp_f(float x) {
  if (x != x) return x;
  if (x > 88.72283935546875f) return bits_f(0x7f800000u);
  if (x < -103.972084045410f) return 0.0f;
  int n;
  float em = exp_core_f(x, &n);
  return scale_f(1.0f + em, n);
}
IMC_HD float expm1_from_core_f(float em, int n) {
  if (n == 0) return em;
  if (n > 24) {
    if (n > 127) return scale_f(1.0f + em, n);
    float t = pow2_f(n);
    return fma_f(t, em, t) - 1.0f;
  }
  if (n < -24) return -1.0f;
  float t = pow2_f(n);
  return fma_f(t, em, t - 1.0f);
}
IMC_HD float expm1_f(float x) {
  if (x != x) return x;
  if (x > 88.72283935546875f) return bits_f(0x7f800000u);
  if (x < -17.5f) return -1.0f;
  float ax = x < 0 ? -x : x;
  if (ax < 2.98023223876953125e-08f) return x;  // 2^-25
  int n;
  float em = exp_core_f(x, &n);
  return expm1_from_core_f(em, n);
}
IMC_HD void exp_expm1_f(float x, float* e, float* em1) {
  float ax = x < 0 ? -x : x;
  // |x| < ln2/2 (the usual case in tracking: one segment attenuates little): n = 0, so the reduction leaves
  // r = x, c = 0 and the scaling is by 2^0 — the same operations as below with the no-ops removed, bit-identical.
  // n = rint(x log2e) == 0  <=>  |RN(x log2e)| <= 1/2 (ties go to the even 0); NaN and the out-of-range arguments fail the
  // test and take the general path below
  const float z = x * IMC_LOG2E_F;
  if ((z < 0 ? -z : z) <= 0.5f) {
    float q = 1.0f / 40320.0f;
    q = fma_f(q, x, 1.0f / 5040.0f);
    q = fma_f(q, x, 1.0f / 720.0f);
    q = fma_f(q, x, 1.0f / 120.0f);
    q = fma_f(q, x, 1.0f / 24.0f);
    q = fma_f(q, x, 1.0f / 6.0f);
    q = fma_f(q, x, 0.5f);
    float em0 = fma_f(x * x, q, x);
    *e = 1.0f + em0;
    *em1 = (ax < 2.98023223876953125e-08f) ? x : em0;
    return;
  }
  if (!(x <= 88.0f && x >= -17.0f)) { *e = exp_f(x); *em1 = expm1_f(x); return; }
  int n;
  float em = exp_core_f(x, &n);
  *e = scale_f(1.0f + em, n);
  *em1 = (ax < 2.98023223876953125e-08f) ? x : expm1_from_core_f(em, n);
}

IMC_HD float log_f(float x) {
  uint32_t ix = f_bits(x);
  int e = 0;
  if (ix >= 0x7f800000u || ix < 0x00800000u) {
    if ((ix << 1) == 0) return -bits_f(0x7f800000u);
    if (ix >> 31) return bits_f(0x7fc00000u);
    if (ix >= 0x7f800000u) return x;
    x *= 33554432.0f;  // 2^25
    ix = f_bits(x);
    e = -25;
  }
  ix += 0x3f800000u - 0x3f3504f3u;
  e += (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  float m = bits_f(ix);
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float z = s * s;
  float p = 1.0f / 13.0f;
  p = fma_f(p, z, 1.0f / 11.0f);
  p = fma_f(p, z, 1.0f / 9.0f);
  p = fma_f(p, z, 1.0f / 7.0f);
  p = fma_f(p, z, 1.0f / 5.0f);
  p = fma_f(p, z, 1.0f / 3.0f);
  float R = 2.0f * (z * p);
  float hfsq = 0.5f * f * f;
  float dk = (float)e;
  const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  return fma_f(s, hfsq + R, dk * ln2_lo) - hfsq + f + dk * ln2_hi;
}

// log_f for a positive, normal, finite argument (the uniform of randexp): the same operations without the
// special-case tests — bit-identical to log_f there
IMC_HD float log_pos_normal_f(float x) {
  uint32_t ix = f_bits(x);
  ix += 0x3f800000u - 0x3f3504f3u;
  int e = (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  float m = bits_f(ix);
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float z = s * s;
  float p = 1.0f / 13.0f;
  p = fma_f(p, z, 1.0f / 11.0f);
  p = fma_f(p, z, 1.0f / 9.0f);
  p = fma_f(p, z, 1.0f / 7.0f);
  p = fma_f(p, z, 1.0f / 5.0f);
  p = fma_f(p, z, 1.0f / 3.0f);
  float R = 2.0f * (z * p);
  float hfsq = 0.5f * f * f;
  float dk = (float)e;
  const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  return fma_f(s, hfsq + R, dk * ln2_lo) - hfsq + f + dk * ln2_hi;
}

// -log(u) for u in (0, 1], positive and normal: the exponential sampler of the Float16 / Float32 tracking loops
// (randexp32_from_word).  Division-free: log(m) = f + f^2 r(f), f = m - 1 in [sqrt(1/2) - 1, sqrt(2) - 1), r a
// degree-8 fit of (log1p(f) - f) / f^2 at Chebyshev nodes (<= 1.3 ulp); -log(u) = (-e) ln2 - log(m), <= 3 ulp
// overall — a sampling transform, not a math-library log (log_f above stays the 1-ulp one).
IMC_HD float neglog_unit_f(float u) {
  uint32_t ix = f_bits(u);
  ix += 0x3f800000u - 0x3f3504f3u;
  int e = (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  float f = bits_f(ix) - 1.0f;  // exact
  float r = bits_f(0xbd9d27fbu);
  r = fma_f(r, f, bits_f(0x3e0247b9u));
  r = fma_f(r, f, bits_f(0xbe066d19u));
  r = fma_f(r, f, bits_f(0x3e11752du));
  r = fma_f(r, f, bits_f(0xbe2a41c1u));
  r = fma_f(r, f, bits_f(0x3e4cd009u));
  r = fma_f(r, f, bits_f(0xbe800106u));
  r = fma_f(r, f, bits_f(0x3eaaaaaau));
  r = fma_f(r, f, bits_f(0xbeffffffu));
  float lg = fma_f(f * f, r, f);
  return fma_f((float)(-e), 6.931471824645996094e-01f, -lg);
}

IMC_HD float sin_kernel_f(float r) {
  float z = r * r;
  float p = -1.0f / 39916800.0f;
  p = fma_f(p, z, 1.0f / 362880.0f);
  p = fma_f(p, z, -1.0f / 5040.0f);
  p = fma_f(p, z, 1.0f / 120.0f);
  p = fma_f(p, z, -1.0f / 6.0f);
  return fma_f(r * z, p, r);
}
IMC_HD float cos_kernel_f(float r) {
  float z = r * r;
  float p = 1.0f / 479001600.0f;
  p = fma_f(p, z, -1.0f / 3628800.0f);
  p = fma_f(p, z, 1.0f / 40320.0f);
  p = fma_f(p, z, -1.0f / 720.0f);
  p = fma_f(p, z, 1.0f / 24.0f);
  float hz = 0.5f * z;
  float w = 1.0f - hz;
  return w + (((1.0f - w) - hz) + (z * z) * p);
}
IMC_HD void sincos_f(float x, float* s, float* c) {
  if (!(x - x == 0.0f)) { *s = x - x; *c = x - x; return; }
  float fq = rint_f(x * 0.636619772367581382433f);
  // reduce in double: exact enough for any float x of moderate size, one rounding to float
  double rd = fma_d((double)fq, -IMC_PIO2_HI_D, (double)x);
  rd = fma_d((double)fq, -IMC_PIO2_LO_D, rd);
  float r = (float)rd;
  float lo = (float)(rd - (double)r);
  int q = (int)fq & 3;
  float sr = sin_kernel_f(r), cr = cos_kernel_f(r);
  sr = fma_f(lo, fma_f(-0.5f * r, r, 1.0f), sr);
  cr = fma_f(-lo, r, cr);
  float ss = (q & 1) ? cr : sr;
  float cc = (q & 1) ? sr : cr;
  if (q == 1 || q == 2) cc = -cc;
  if (q >= 2) ss = -ss;
  *s = ss;
  *c = cc;
}
IMC_HD float atan2_f(float y, float x) { return (float)atan2_d((double)y, (double)x); }
IMC_HD float pow_f(float x, float y) { return (float)pow_d((double)x, (double)y); }

}  // namespace dm

// ---------------------------------------------------------------------------------------
// Math policy used by both kernels and oracle: deterministic functions, results in Num<P>.
// Float16: compute in Float32 and round once (Julia: exp(x::Float16) = Float16(exp(Float32(x)))).
// ---------------------------------------------------------------------------------------
struct MathDet {
  template <class P> static IMC_HD Num<P> exp(Num<P> x) {
    if constexpr (P::id == 2) return Num<P>(dm::exp_d(x.v)); else return Num<P>(P::rnd(dm::exp_f(x.v)));
  }
  template <class P> static IMC_HD Num<P> expm1(Num<P> x) {
    if constexpr (P::id == 2) return Num<P>(dm::expm1_d(x.v)); else return Num<P>(P::rnd(dm::expm1_f(x.v)));
  }
  template <class P> static IMC_HD void exp_expm1(Num<P> x, Num<P>* e, Num<P>* em1) {
    if constexpr (P::id == 2) { double a, b; dm::exp_expm1_d(x.v, &a, &b); *e = Num<P>(a); *em1 = Num<P>(b); }
    else { float a, b; dm::exp_expm1_f(x.v, &a, &b); *e = Num<P>(P::rnd(a)); *em1 = Num<P>(P::rnd(b)); }
  }
  template <class P> static IMC_HD Num<P> log(Num<P> x) {
    if constexpr (P::id == 2) return Num<P>(dm::log_d(x.v)); else return Num<P>(P::rnd(dm::log_f(x.v)));
  }
  template <class P> static IMC_HD Num<P> sqrt(Num<P> x) {
    if constexpr (P::id == 2) return Num<P>(dm::sqrt_d(x.v)); else return Num<P>(P::rnd(dm::sqrt_f(x.v)));
  }
  template <class P> static IMC_HD void sincos(Num<P> x, Num<P>* s, Num<P>* c) {
    if constexpr (P::id == 2) { double a, b; dm::sincos_d(x.v, &a, &b); *s = Num<P>(a); *c = Num<P>(b); }
    else { float a, b; dm::sincos_f(x.v, &a, &b); *s = Num<P>(P::rnd(a)); *c = Num<P>(P::rnd(b)); }
  }
  template <class P> static IMC_HD Num<P> atan2(Num<P> y, Num<P> x) {
    if constexpr (P::id == 2) return Num<P>(dm::atan2_d(y.v, x.v)); else return Num<P>(P::rnd(dm::atan2_f(y.v, x.v)));
  }
  template <class P> static IMC_HD Num<P> pow(Num<P> x, Num<P> y) {
    if constexpr (P::id == 2) return Num<P>(dm::pow_d(x.v, y.v)); else return Num<P>(P::rnd(dm::pow_f(x.v, y.v)));
  }
  // Float64 entry points (for the places the reference computes in Float64)
  static IMC_HD double exp64(double x) { return dm::exp_d(x); }
  static IMC_HD double expm164(double x) { return dm::expm1_d(x); }
  static IMC_HD double log64(double x) { return dm::log_d(x); }
  static IMC_HD double pow64(double x, double y) { return dm::pow_d(x, y); }
  static IMC_HD double sqrt64(double x) { return dm::sqrt_d(x); }
  static IMC_HD double cos64(double x) { double s, c; dm::sincos_d(x, &s, &c); return c; }
};

}  // namespace imc
