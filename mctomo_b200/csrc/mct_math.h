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
 * Shape.  The polynomials are evaluated with Estrin's scheme (independent partial sums)
 * instead of Horner's: a dependent DFMA costs ~8 cycles of latency on sm_100a, and a
 * 13-step Horner chain left the FP64 pipe idle most of the time (profiles/).  On the device
 * the coefficients live in __constant__ memory so they arrive as constant-bank operands
 * rather than as two UMOV immediates per DFMA.
 *
 * Accuracy (tests/test_mct_math.py, against mpmath): sin/cos < 1.5 ulp for |x| <= 1e4,
 * exp < 1 ulp on [-700, 700].  Domain notes:
 *   mct_exp_core : no argument checks; requires |x| <= 700.  The kernels only ever need
 *                  x in [-60, 0] (exp(-2p), p < 16; exp(-exa), exa < 60).
 *   mct_exp      : adds the checks (0 below -700, clamps above 700, NaN through).
 *   mct_sincos   : 3-term Cody-Waite reduction; accurate for |x| < ~1e5, deterministic
 *                  (and identical on both sides) for any finite x with |x| < 2^50.
 */
#ifndef MCT_MATH_H
#define MCT_MATH_H

#include <stdint.h>
#include <string.h>

/* coefficient table: one definition, two storage classes */
#define MCT_KTAB_INIT                                                                         \
  {                                                                                           \
    /*  0 S1 */ -1.66666666666666324348e-01, /*  1 S2 */ 8.33333333332248946124e-03,          \
    /*  2 S3 */ -1.98412698298579493134e-04, /*  3 S4 */ 2.75573137070700676789e-06,          \
    /*  4 S5 */ -2.50507602534068634195e-08, /*  5 S6 */ 1.58969099521155010221e-10,          \
    /*  6 C1 */ 4.16666666666666019037e-02,  /*  7 C2 */ -1.38888888888741095749e-03,         \
    /*  8 C3 */ 2.48015872894767294178e-05,  /*  9 C4 */ -2.75573143513906633035e-07,         \
    /* 10 C5 */ 2.08757232129817482790e-09,  /* 11 C6 */ -1.13596475577881948265e-11,         \
    /* 12 2/pi */ 6.36619772367581382433e-01, /* 13 P1 */ 1.5707963267948966e+00,             \
    /* 14 P2 */ 6.123233995736766e-17,       /* 15 P3 */ -1.4973849048591698e-33,             \
    /* 16 1/ln2 */ 1.44269504088896338700e+00, /* 17 LN2_HI */ 6.93147180369123816490e-01,    \
    /* 18 LN2_LO */ 1.90821492927058770002e-10,                                               \
    /* 19.. 1/n!, n = 2..13 */ 0.5, 0x1.5555555555555p-3, 0x1.5555555555555p-5,               \
    0x1.1111111111111p-7, 0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-16, \
    0x1.71de3a556c734p-19, 0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26,                      \
    0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33,                                             \
    /* 31 64/ln2 */ 0x1.71547652b82fep+6,                                                     \
    /* 32 ln2/64 hi (28 trailing zero bits: n*HI is exact) */ 0x1.62e42f0000000p-7,           \
    /* 33 ln2/64 lo */ 0x1.df473de6af279p-32                                                  \
  }
#define MCT_K_S(i) MCT_K((i) - 1)        /* S1..S6 */
#define MCT_K_C(i) MCT_K(6 + (i) - 1)    /* C1..C6 */
#define MCT_K_2OPI MCT_K(12)
#define MCT_K_P1 MCT_K(13)
#define MCT_K_P2 MCT_K(14)
#define MCT_K_P3 MCT_K(15)
#define MCT_K_INVLN2 MCT_K(16)
#define MCT_K_LN2HI MCT_K(17)
#define MCT_K_LN2LO MCT_K(18)
#define MCT_K_F(n) MCT_K(19 + (n) - 2)   /* 1/n!, n = 2..13 */

#if defined(__CUDACC__)
__constant__ double mct_ktab_dev[34] = MCT_KTAB_INIT;
static const double mct_ktab_host[34] = MCT_KTAB_INIT;
#if defined(__CUDA_ARCH__)
#define MCT_K(i) mct_ktab_dev[i]
#else
#define MCT_K(i) mct_ktab_host[i]
#endif
#else
static const double mct_ktab_host[34] = MCT_KTAB_INIT;
#define MCT_K(i) mct_ktab_host[i]
#endif

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

/* ---- sin & cos on the reduced argument |r| <= pi/4 (fdlibm's minimax coefficients, Estrin order) */
MCT_HD double mct_ksin(double r) {
  const double z = MCT_MUL(r, r);
  const double z2 = MCT_MUL(z, z);
  const double a = MCT_FMA(z, MCT_K_S(2), MCT_K_S(1));
  const double b = MCT_FMA(z, MCT_K_S(4), MCT_K_S(3));
  const double c = MCT_FMA(z, MCT_K_S(6), MCT_K_S(5));
  const double z4 = MCT_MUL(z2, z2);
  double p = MCT_FMA(b, z2, a);
  p = MCT_FMA(c, z4, p);
  return MCT_FMA(MCT_MUL(z, r), p, r);
}
MCT_HD double mct_kcos(double r) {
  const double z = MCT_MUL(r, r);
  const double z2 = MCT_MUL(z, z);
  const double a = MCT_FMA(z, MCT_K_C(2), MCT_K_C(1));
  const double b = MCT_FMA(z, MCT_K_C(4), MCT_K_C(3));
  const double c = MCT_FMA(z, MCT_K_C(6), MCT_K_C(5));
  const double z4 = MCT_MUL(z2, z2);
  double p = MCT_FMA(b, z2, a);
  p = MCT_FMA(c, z4, p);
  /* 1 - z/2 + z^2 p, the -z/2 split off exactly (hz exact, 1-hz rounded once, error recovered) */
  const double hz = MCT_MUL(0.5, z);
  const double w = MCT_ADD(1.0, -hz);
  const double e = MCT_ADD(MCT_ADD(1.0, -w), -hz); /* exact rounding error of w */
  return MCT_ADD(w, MCT_FMA(z2, p, e));
}

/* sin(x) and cos(x) together.  */
MCT_HD void mct_sincos(double x, double* sn, double* cs) {
  const double t = MCT_FMA(x, MCT_K_2OPI, MCT_MAGIC);
  const uint32_t q = (uint32_t)mct_d2bits(t);
  const double k = MCT_ADD(t, -MCT_MAGIC);
  double r = MCT_FMA(-k, MCT_K_P1, x);
  r = MCT_FMA(-k, MCT_K_P2, r);
  r = MCT_FMA(-k, MCT_K_P3, r);
  const double s = mct_ksin(r), c = mct_kcos(r);
  const double a = (q & 1u) ? c : s;
  const double b = (q & 1u) ? s : c;
  *sn = (q & 2u) ? -a : a;
  *cs = ((q + 1u) & 2u) ? -b : b;
}

/* 2^(j/64), j = 0..63, correctly rounded (generated with mpmath) */
#define MCT_EXPTAB_INIT                                                                          \
  {                                                                                              \
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,      \
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,      \
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,      \
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,      \
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,      \
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,      \
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,      \
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,      \
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,      \
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,      \
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,      \
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,      \
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,      \
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,      \
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,      \
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0       \
  }
#if defined(__CUDACC__)
/* Lanes index the table independently, so it cannot live in the constant bank (a divergent LDC serialises).
 * It is staged in shared memory: one LDS per lookup, against an LDG plus 64-bit address arithmetic plus two
 * R2UR descriptor moves from global memory (profiles/: 12 of ~230 non-FP64 instructions per layer step).
 * Every kernel that evaluates mct_exp* must call mct_exptab_stage() first. */
__device__ const double mct_exptab_dev[64] = MCT_EXPTAB_INIT;
__shared__ double mct_exptab_s[64];
static const double mct_exptab_host[64] = MCT_EXPTAB_INIT;
__device__ __forceinline__ void mct_exptab_stage() {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) mct_exptab_s[i] = mct_exptab_dev[i];
  __syncthreads();
}
#if defined(__CUDA_ARCH__)
#define MCT_EXPTAB(j) mct_exptab_s[j]
#else
#define MCT_EXPTAB(j) mct_exptab_host[j]
#endif
#else
static const double mct_exptab_host[64] = MCT_EXPTAB_INIT;
#define MCT_EXPTAB(j) mct_exptab_host[j]
#endif

/* exp(x) for |x| <= 700, no argument checks: x = (64 k + j) ln2/64 + r, |r| <= ln2/128,
 * exp(x) = 2^k * 2^(j/64) * (1 + expm1(r)), expm1 by a degree-6 Taylor polynomial (truncation 3e-20).
 * 11 FP64 operations and one table load; half the length of a table-free polynomial. */
MCT_HD double mct_exp_core(double x) {
  const double INV = MCT_K(31), HI = MCT_K(32), LO = MCT_K(33); /* 64/ln2; ln2/64 split */
  const double t = MCT_FMA(x, INV, MCT_MAGIC);
  const int32_t n = (int32_t)(uint32_t)mct_d2bits(t);
  const double nf = MCT_ADD(t, -MCT_MAGIC);
  double r = MCT_FMA(-nf, HI, x);
  r = MCT_FMA(-nf, LO, r);
  const double T = MCT_EXPTAB(n & 63);
  const int32_t k = n >> 6; /* arithmetic shift: floor */
  const double r2 = MCT_MUL(r, r);
  const double a = MCT_FMA(r, MCT_K_F(3), MCT_K_F(2));
  const double b = MCT_FMA(r, MCT_K_F(5), MCT_K_F(4));
  const double q = MCT_FMA(r2, MCT_FMA(r2, MCT_K_F(6), b), a);
  const double s = MCT_FMA(q, r2, r); /* expm1(r) */
  const double p = MCT_FMA(T, s, T);
  return mct_bits2d(mct_d2bits(p) + ((uint64_t)(int64_t)k << 52));
}

/* The table-free form (degree-13 Taylor, Estrin order); kept for reference and A/B accuracy tests. */
MCT_HD double mct_exp_core_poly(double x) {
  const double t = MCT_FMA(x, MCT_K_INVLN2, MCT_MAGIC);
  const int32_t ki = (int32_t)(uint32_t)mct_d2bits(t);
  const double k = MCT_ADD(t, -MCT_MAGIC);
  double r = MCT_FMA(-k, MCT_K_LN2HI, x);
  r = MCT_FMA(-k, MCT_K_LN2LO, r);
  const double r2 = MCT_MUL(r, r);
  const double a1 = MCT_FMA(MCT_K_F(3), r, MCT_K_F(2));
  const double a2 = MCT_FMA(MCT_K_F(5), r, MCT_K_F(4));
  const double a3 = MCT_FMA(MCT_K_F(7), r, MCT_K_F(6));
  const double a4 = MCT_FMA(MCT_K_F(9), r, MCT_K_F(8));
  const double a5 = MCT_FMA(MCT_K_F(11), r, MCT_K_F(10));
  const double a6 = MCT_FMA(MCT_K_F(13), r, MCT_K_F(12));
  const double r4 = MCT_MUL(r2, r2);
  const double b1 = MCT_FMA(a2, r2, a1); /* r^2..r^5 terms, divided by r^2 */
  const double b2 = MCT_FMA(a4, r2, a3); /* r^6..r^9 */
  const double b3 = MCT_FMA(a6, r2, a5); /* r^10..r^13 */
  const double r8 = MCT_MUL(r4, r4);
  const double d = MCT_FMA(b2, r4, b1);
  const double q = MCT_FMA(b3, r8, d);   /* (exp(r) - 1 - r) / r^2 */
  const double s = MCT_FMA(q, r2, r);    /* exp(r) - 1, small: only the final 1 + s rounds at ulp(1) */
  const double p = MCT_ADD(1.0, s);
  return mct_bits2d(mct_d2bits(p) + ((uint64_t)(int64_t)ki << 52));
}

MCT_HD double mct_exp(double x) {
  if (!(x >= -700.0)) return (x != x) ? x : 0.0;
  if (x > 700.0) x = 700.0;
  return mct_exp_core(x);
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


/* log(x) for finite x > 0 (natural logarithm; the misfit's sum(log(sigma)), src/likelihood_surf.F90:404).
 * fdlibm's e_log.c scheme with explicit IEEE operations: x = 2^k (1+f), sqrt(2)/2 < 1+f < sqrt(2);
 * s = f/(2+f); log(1+f) = f - (f^2/2 - s (f^2/2 + R(s^2))); < 1 ulp.  Subnormal x is scaled by 2^54 first.
 * x <= 0, NaN and Inf are the caller's business (the kernels check sigma >= EPS before taking its logarithm). */
MCT_HD double mct_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t u = mct_d2bits(x);
  int32_t k = 0;
  if ((u >> 52) == 0) { /* subnormal */
    x = MCT_MUL(x, 18014398509481984.0); /* 2^54 */
    u = mct_d2bits(x);
    k = -54;
  }
  uint32_t hx = (uint32_t)(u >> 32);
  k += (int32_t)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  const uint32_t i = (hx + 0x95f64u) & 0x100000u; /* 1 when the mantissa exceeds sqrt(2): halve it */
  k += (int32_t)(i >> 20);
  u = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (u & 0xffffffffull);
  const double f = MCT_ADD(mct_bits2d(u), -1.0);
  const double dk = (double)k;
  const double s = MCT_DIV(f, MCT_ADD(2.0, f));
  const double z = MCT_MUL(s, s);
  const double w = MCT_MUL(z, z);
  const double t1 = MCT_MUL(w, MCT_FMA(w, MCT_FMA(w, Lg6, Lg4), Lg2));
  const double t2 = MCT_MUL(z, MCT_FMA(w, MCT_FMA(w, MCT_FMA(w, Lg7, Lg5), Lg3), Lg1));
  const double R = MCT_ADD(t2, t1);
  const double hfsq = MCT_MUL(0.5, MCT_MUL(f, f));
  /* k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f) */
  const double inner = MCT_ADD(MCT_MUL(s, MCT_ADD(hfsq, R)), MCT_MUL(dk, ln2_lo));
  return MCT_ADD(MCT_MUL(dk, ln2_hi), -MCT_ADD(MCT_ADD(hfsq, -inner), -f));
}

#endif /* MCT_MATH_H */
