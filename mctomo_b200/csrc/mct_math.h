/*
 * mct_math.h -- portable, bit-reproducible double-precision sin/cos/exp/x^(1/4).
 *
 * Why this exists.  The reference's secular functions (surfmodes/surfdisp96.f:1056-1337)
 * call libm's dsin/dcos/dexp.  glibc's and CUDA's libm differ in the last bit for a
 * few percent of arguments, which is invisible after surfdisp96 rounds the root to
 * float32 (surfdisp96.f:313) except when the root sits on a float rounding boundary.
 * To make "GPU == CPU restatement" a bit-exact gate rather than a statistical one,
 * the CUDA kernels use ONLY the functions below, and the oracle can be switched to
 * the same functions ("portable" math mode); every operation in them is an IEEE-754
 * add / mul / fma, so host (gcc, -ffp-contract=off, hardware FMA) and device
 * (nvcc -fmad=false, explicit __fma_rn) produce identical bits.
 *
 * Accuracy (measured against mpmath in tests/test_mct_math.py): < 1.5 ulp for
 * |x| <= 1e4 (sin/cos), < 1 ulp on [-700, 0] (exp).  Domain notes:
 *   mct_exp    : intended for x <= 0 (the reference only ever calls exp(-2p), p<16,
 *                exp(-exa), exa<60); valid for -700 <= x <= 700, returns 0 below.
 *   mct_sincos : 3-term Cody-Waite reduction; accurate for |x| < ~1e5, deterministic
 *                (and identical on both sides) for any finite x with |x| < 2^50.
 */
#ifndef MCT_MATH_H
#define MCT_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define MCT_HD __device__ __forceinline__
#define MCT_FMA(a, b, c) __fma_rn((a), (b), (c))
#define MCT_MUL(a, b) __dmul_rn((a), (b))
#define MCT_ADD(a, b) __dadd_rn((a), (b))
#define MCT_SQRT(a) __dsqrt_rn(a)
#define MCT_DIV(a, b) __ddiv_rn((a), (b))
#elif defined(__CUDACC__)
/* host pass of nvcc: never executed, only has to parse */
#include <math.h>
#define MCT_HD __host__ __device__ inline
#define MCT_FMA(a, b, c) fma((a), (b), (c))
#define MCT_MUL(a, b) ((a) * (b))
#define MCT_ADD(a, b) ((a) + (b))
#define MCT_SQRT(a) sqrt(a)
#define MCT_DIV(a, b) ((a) / (b))
#else
#include <math.h>
#define MCT_HD static inline
#define MCT_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define MCT_MUL(a, b) ((a) * (b))
#define MCT_ADD(a, b) ((a) + (b))
#define MCT_SQRT(a) __builtin_sqrt(a)
#define MCT_DIV(a, b) ((a) / (b))
#endif

MCT_HD double mct_bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}
MCT_HD uint64_t mct_d2bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

/* 1.5 * 2^52: adding it rounds to the nearest integer (ties to even) and leaves
 * that integer in the low mantissa bits. */
#define MCT_MAGIC 6755399441055744.0

/* ---- sin & cos on the reduced argument |r| <= pi/4 (fdlibm-style minimax polynomials) */
MCT_HD double mct_ksin(double r) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = MCT_MUL(r, r);
  double p = MCT_FMA(z, S6, S5);
  p = MCT_FMA(z, p, S4);
  p = MCT_FMA(z, p, S3);
  p = MCT_FMA(z, p, S2);
  p = MCT_FMA(z, p, S1);
  return MCT_FMA(MCT_MUL(z, r), p, r);
}
MCT_HD double mct_kcos(double r) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = MCT_MUL(r, r);
  double p = MCT_FMA(z, C6, C5);
  p = MCT_FMA(z, p, C4);
  p = MCT_FMA(z, p, C3);
  p = MCT_FMA(z, p, C2);
  p = MCT_FMA(z, p, C1);
  /* 1 - z/2 + z^2 p, the -z/2 split off exactly (hz exact, 1-hz rounded once, error recovered) */
  double hz = MCT_MUL(0.5, z);
  double w = MCT_ADD(1.0, -hz);
  double e = MCT_ADD(MCT_ADD(1.0, -w), -hz); /* exact rounding error of w */
  return MCT_ADD(w, MCT_FMA(MCT_MUL(z, z), p, e));
}

/* sin(x) and cos(x) together.  */
MCT_HD void mct_sincos(double x, double* sn, double* cs) {
  const double TWO_OVER_PI = 6.36619772367581382433e-01; /* 0x1.45f306dc9c883p-1 */
  const double P1 = 1.5707963267948966e+00;              /* 0x1.921fb54442d18p+0 */
  const double P2 = 6.123233995736766e-17;               /* 0x1.1a62633145c07p-54 */
  const double P3 = -1.4973849048591698e-33;             /* -0x1.f1976b7ed8fbcp-110 */
  double t = MCT_FMA(x, TWO_OVER_PI, MCT_MAGIC);
  uint32_t q = (uint32_t)mct_d2bits(t);
  double k = MCT_ADD(t, -MCT_MAGIC);
  double r = MCT_FMA(-k, P1, x);
  r = MCT_FMA(-k, P2, r);
  r = MCT_FMA(-k, P3, r);
  double s = mct_ksin(r), c = mct_kcos(r);
  double a = (q & 1u) ? c : s;
  double b = (q & 1u) ? s : c;
  *sn = (q & 2u) ? -a : a;
  *cs = ((q + 1u) & 2u) ? -b : b;
}

/* exp(x).  k = rint(x/ln2), r = x - k ln2 (2-term), degree-13 Taylor, scale by 2^k
 * through the exponent field (result is always normal on the stated domain). */
MCT_HD double mct_exp(double x) {
  const double INV_LN2 = 1.44269504088896338700e+00; /* 0x1.71547652b82fep+0 */
  const double LN2_HI = 6.93147180369123816490e-01;  /* 0x1.62e42fee00000p-1 */
  const double LN2_LO = 1.90821492927058770002e-10;  /* 0x1.a39ef35793c76p-33 */
  if (!(x >= -700.0)) return (x != x) ? x : 0.0;
  if (x > 700.0) x = 700.0;
  double t = MCT_FMA(x, INV_LN2, MCT_MAGIC);
  int32_t ki = (int32_t)(uint32_t)mct_d2bits(t);
  double k = MCT_ADD(t, -MCT_MAGIC);
  double r = MCT_FMA(-k, LN2_HI, x);
  r = MCT_FMA(-k, LN2_LO, r);
  double p = 0x1.6124613a86d09p-33;           /* 1/13! */
  p = MCT_FMA(p, r, 0x1.1eed8eff8d898p-29);   /* 1/12! */
  p = MCT_FMA(p, r, 0x1.ae64567f544e4p-26);   /* 1/11! */
  p = MCT_FMA(p, r, 0x1.27e4fb7789f5cp-22);   /* 1/10! */
  p = MCT_FMA(p, r, 0x1.71de3a556c734p-19);   /* 1/9!  */
  p = MCT_FMA(p, r, 0x1.a01a01a01a01ap-16);   /* 1/8!  */
  p = MCT_FMA(p, r, 0x1.a01a01a01a01ap-13);   /* 1/7!  */
  p = MCT_FMA(p, r, 0x1.6c16c16c16c17p-10);   /* 1/6!  */
  p = MCT_FMA(p, r, 0x1.1111111111111p-7);    /* 1/5!  */
  p = MCT_FMA(p, r, 0x1.5555555555555p-5);    /* 1/4!  */
  p = MCT_FMA(p, r, 0x1.5555555555555p-3);    /* 1/3!  */
  p = MCT_FMA(p, r, 0.5);
  p = MCT_FMA(p, r, 1.0);
  p = MCT_FMA(p, r, 1.0);
  return mct_bits2d(mct_d2bits(p) + ((uint64_t)(int64_t)ki << 52));
}

/* x^(1/4) for x > 0: sqrt(sqrt(x)) followed by one Newton correction carried out
 * with an FMA-exact residual, which makes the result correctly rounded except for
 * near-halfway cases (src/utils.f90:133 `rho = 1.74*vp**0.25`). */
MCT_HD double mct_pow025(double x) {
  if (!(x > 0.0)) return (x == 0.0) ? 0.0 : (x - x) / (x - x);
  double y = MCT_SQRT(MCT_SQRT(x));
  double y2 = MCT_MUL(y, y);
  double e2 = MCT_FMA(y, y, -y2);           /* y*y = y2 + e2 exactly */
  double y4 = MCT_MUL(y2, y2);
  double e4 = MCT_FMA(y2, y2, -y4);         /* y2*y2 = y4 + e4 exactly */
  /* y^4 = (y2+e2)^2 ~= y4 + e4 + 2*y2*e2 */
  double res = MCT_ADD(MCT_ADD(MCT_ADD(x, -y4), -e4), -MCT_MUL(MCT_MUL(2.0, y2), e2));
  double corr = MCT_DIV(MCT_MUL(res, y), MCT_MUL(4.0, y4));
  return MCT_ADD(y, corr);
}

#endif /* MCT_MATH_H */
