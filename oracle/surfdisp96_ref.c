/*
 * oracle/surfdisp96_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's modal dispersion solver, written to
 * be the parity checker for the CUDA kernels in mctomo_b200/csrc/.  Nothing in the
 * product path may link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY STATUS: pinned on the reference's own source.  oracle/f77toc.py translates
 * /root/reference/surfmodes/surfdisp96.f mechanically to C (oracle/_ref/, git-ignored);
 * this restatement agrees with that translation bit for bit on 1 039 committed fixtures
 * (tests/golden/dispersion_ref.npz) and on fresh random stacks wherever oracle/_ref exists
 * (tests/test_oracle_vs_reference.py).  Not yet diffed against a gfortran binary: the recipe
 * (oracle/build_ref_surfdisp.sh) and its test are committed and reported skipped without a
 * Fortran compiler.  Physics known-answer tests: tests/test_oracle_dispersion.py,
 * tests/test_oracle_physics.py.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  The Fortran typing is kept variable by variable: what is
 * real*4 there is `float` here, what is double there is `double` here, mixed
 * expressions are widened exactly where Fortran widens them, and the file must be
 * compiled with -ffp-contract=off (no FMA contraction; the reference's release
 * flags do not enable FMA, src/makefile:57-66).
 *
 * math_mode: 0 = libm sin/cos/exp (what the Fortran binary calls), 1 = the portable
 * functions of mctomo_b200/csrc/mct_math.h (what the CUDA kernels use).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../mctomo_b200/csrc/mct_math.h"

#define NL 200 /* surfdisp96.f:57 */
#define NP 60  /* surfdisp96.f:59 */

typedef struct {
  int mmax, llw;
  float d[NL], a[NL], b[NL], rho[NL];
  int math_mode;
  /* instrumentation: identical search path <=> identical counts */
  int64_t n_dltar;  /* secular-function evaluations */
  int64_t n_layer;  /* layer steps inside them */
  double del1st;    /* getsol's SAVE variable (surfdisp96.f:759), kept per call */
} sd_ctx;

static inline double sgn1(double x) { return copysign(1.0, x); } /* dsign(1.0d0,x) */

static inline void o_sincos(const sd_ctx* S, double x, double* s, double* c) {
  if (S->math_mode) { mct_sincos(x, s, c); } else { *s = sin(x); *c = cos(x); }
}
static inline double o_exp(const sd_ctx* S, double x) {
  return S->math_mode ? mct_exp(x) : exp(x);
}

/* gtsolh: surfdisp96.f:711-732.  All single precision (implicit typing). */
static float gtsolh(float a, float b) {
  float c = 0.95f * b;
  for (int i = 0; i < 5; ++i) {
    float gamma = b / a;
    float kappa = c / b;
    float k2 = kappa * kappa;
    float gk2 = (gamma * kappa) * (gamma * kappa);
    float fac1 = sqrtf(1.0f - gk2);
    float fac2 = sqrtf(1.0f - k2);
    float fr = (2.0f - k2) * (2.0f - k2) - 4.0f * fac1 * fac2;
    float frp = -4.0f * (2.0f - k2) * kappa + 4.0f * fac2 * gamma * gamma * kappa / fac1 +
                4.0f * fac1 * kappa / fac2;
    frp = frp / b;
    c = c - fr / frp;
  }
  return c;
}

/* dltar1 (Love, Haskell): surfdisp96.f:1056-1115 */
static double dltar1(sd_ctx* S, double wvno, double omega) {
  const int mmax = S->mmax, llw = S->llw;
  double beta1 = (double)S->b[mmax - 1];
  double rho1 = (double)S->rho[mmax - 1];
  double xkb = omega / beta1;
  double wvnop = wvno + xkb;
  double wvnom = fabs(wvno - xkb);
  double rb = sqrt(wvnop * wvnom);
  double e1 = rho1 * rb;
  double e2 = 1.0 / (beta1 * beta1);
  for (int m = mmax - 1; m >= llw; --m) { /* Fortran m = mmax-1 .. llw, 1-based */
    S->n_layer++;
    beta1 = (double)S->b[m - 1];
    rho1 = (double)S->rho[m - 1];
    double xmu = rho1 * beta1 * beta1;
    xkb = omega / beta1;
    wvnop = wvno + xkb;
    wvnom = fabs(wvno - xkb);
    rb = sqrt(wvnop * wvnom);
    double q = (double)S->d[m - 1] * rb;
    double sinq, cosq, y, z;
    if (wvno < xkb) {
      o_sincos(S, q, &sinq, &cosq);
      y = sinq / rb;
      z = -rb * sinq;
    } else if (wvno == xkb) {
      cosq = 1.0;
      y = (double)S->d[m - 1];
      z = 0.0;
    } else {
      double fac = 0.0;
      if (q < 16.0) fac = o_exp(S, -2.0 * q);
      cosq = (1.0 + fac) * 0.5;
      sinq = (1.0 - fac) * 0.5;
      y = sinq / rb;
      z = rb * sinq;
    }
    double e10 = e1 * cosq + e2 * xmu * z;
    double e20 = e1 * y / xmu + e2 * cosq;
    double xnor = fabs(e10);
    double ynor = fabs(e20);
    if (ynor > xnor) xnor = ynor;
    if (xnor < 1.e-40) xnor = 1.0;
    e1 = e10 / xnor;
    e2 = e20 / xnor;
  }
  return e1;
}

/* var: surfdisp96.f:1220-1337.  The trailing cosq/y/z rescaling (:1330-1335) only
 * touches locals that are never read again, so it is not restated. */
typedef struct { double a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz, w, cosp; } var_out;

static void var_(const sd_ctx* S, double p, double q, double ra, double rb, double wvno,
                 double xka, double xkb, double dpth, var_out* o) {
  double pex = 0.0, sex = 0.0;
  double sinp, cosp, w, x, sinq, cosq, y, z, fac;
  if (wvno < xka) {
    o_sincos(S, p, &sinp, &cosp);
    w = sinp / ra;
    x = -ra * sinp;
  } else if (wvno == xka) {
    cosp = 1.0;
    w = dpth;
    x = 0.0;
  } else {
    pex = p;
    fac = 0.0;
    if (p < 16.0) fac = o_exp(S, -2.0 * p);
    cosp = (1.0 + fac) * 0.5;
    sinp = (1.0 - fac) * 0.5;
    w = sinp / ra;
    x = ra * sinp;
  }
  if (wvno < xkb) {
    o_sincos(S, q, &sinq, &cosq);
    y = sinq / rb;
    z = -rb * sinq;
  } else if (wvno == xkb) {
    cosq = 1.0;
    y = dpth;
    z = 0.0;
  } else {
    sex = q;
    fac = 0.0;
    if (q < 16.0) fac = o_exp(S, -2.0 * q);
    cosq = (1.0 + fac) * 0.5;
    sinq = (1.0 - fac) * 0.5;
    y = sinq / rb;
    z = rb * sinq;
  }
  double exa = pex + sex;
  double a0 = 0.0;
  if (exa < 60.0) a0 = o_exp(S, -exa);
  o->a0 = a0;
  o->cpcq = cosp * cosq;
  o->cpy = cosp * y;
  o->cpz = cosp * z;
  o->cqw = cosq * w;
  o->cqx = cosq * x;
  o->xy = x * y;
  o->xz = x * z;
  o->wy = w * y;
  o->wz = w * z;
  o->w = w;
  o->cosp = cosp;
}

/* dnka (Dunkin's matrix): surfdisp96.f:1370-1414.  ca[j][i] = ca(j+1,i+1). */
static void dnka(double ca[5][5], double wvno2, double gam, double gammk, double rho,
                 const var_out* v) {
  const double one = 1.0, two = 2.0;
  double a0 = v->a0, cpcq = v->cpcq, cpy = v->cpy, cpz = v->cpz, cqw = v->cqw, cqx = v->cqx,
         xy = v->xy, xz = v->xz, wy = v->wy, wz = v->wz;
  double gamm1 = gam - one;
  double twgm1 = gam + gamm1;
  double gmgmk = gam * gammk;
  double gmgm1 = gam * gamm1;
  double gm1sq = gamm1 * gamm1;
  double rho2 = rho * rho;
  double a0pq = a0 - cpcq;
  ca[0][0] = cpcq - two * gmgm1 * a0pq - gmgmk * xz - wvno2 * gm1sq * wy;
  ca[0][1] = (wvno2 * cpy - cqx) / rho;
  ca[0][2] = -(twgm1 * a0pq + gammk * xz + wvno2 * gamm1 * wy) / rho;
  ca[0][3] = (cpz - wvno2 * cqw) / rho;
  ca[0][4] = -(two * wvno2 * a0pq + xz + wvno2 * wvno2 * wy) / rho2;
  ca[1][0] = (gmgmk * cpz - gm1sq * cqw) * rho;
  ca[1][1] = cpcq;
  ca[1][2] = gammk * cpz - gamm1 * cqw;
  ca[1][3] = -wz;
  ca[1][4] = ca[0][3];
  ca[3][0] = (gm1sq * cpy - gmgmk * cqx) * rho;
  ca[3][1] = -xy;
  ca[3][2] = gamm1 * cpy - gammk * cqx;
  ca[3][3] = ca[1][1];
  ca[3][4] = ca[0][1];
  ca[4][0] = -(two * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2;
  ca[4][1] = ca[3][0];
  ca[4][2] = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho;
  ca[4][3] = ca[1][0];
  ca[4][4] = ca[0][0];
  double t = -two * wvno2;
  ca[2][0] = t * ca[4][2];
  ca[2][1] = t * ca[3][2];
  ca[2][2] = a0 + two * (cpcq - ca[0][0]);
  ca[2][3] = t * ca[1][2];
  ca[2][4] = t * ca[0][2];
}

/* normc: surfdisp96.f:1341-1366.  The dlog() of the scale (:1364) is returned in an
 * argument dltar4 overwrites and never reads (:1191), so it is not evaluated. */
static void normc(double ee[5]) {
  double t1 = 0.0;
  for (int i = 0; i < 5; ++i)
    if (fabs(ee[i]) > t1) t1 = fabs(ee[i]);
  if (t1 < 1.e-40) t1 = 1.0;
  for (int i = 0; i < 5; ++i) ee[i] = ee[i] / t1;
}

/* dltar4 (Rayleigh, Dunkin compound matrix): surfdisp96.f:1119-1217 */
static double dltar4(sd_ctx* S, double wvno, double omga) {
  const int mmax = S->mmax, llw = S->llw;
  double e[5], ee[5], ca[5][5];
  var_out v;
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  double wvno2 = wvno * wvno;
  double xka = omega / (double)S->a[mmax - 1];
  double xkb = omega / (double)S->b[mmax - 1];
  double wvnop = wvno + xka;
  double wvnom = fabs(wvno - xka);
  double ra = sqrt(wvnop * wvnom);
  wvnop = wvno + xkb;
  wvnom = fabs(wvno - xkb);
  double rb = sqrt(wvnop * wvnom);
  double t = (double)S->b[mmax - 1] / omega;
  double gammk = 2.0 * t * t;
  double gam = gammk * wvno2;
  double gamm1 = gam - 1.0;
  double rho1 = (double)S->rho[mmax - 1];
  e[0] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
  e[1] = -rho1 * ra;
  e[2] = rho1 * (gamm1 - gammk * ra * rb);
  e[3] = rho1 * rb;
  e[4] = wvno2 - ra * rb;
  for (int m = mmax - 1; m >= llw; --m) {
    S->n_layer++;
    xka = omega / (double)S->a[m - 1];
    xkb = omega / (double)S->b[m - 1];
    t = (double)S->b[m - 1] / omega;
    gammk = 2.0 * t * t;
    gam = gammk * wvno2;
    wvnop = wvno + xka;
    wvnom = fabs(wvno - xka);
    ra = sqrt(wvnop * wvnom);
    wvnop = wvno + xkb;
    wvnom = fabs(wvno - xkb);
    rb = sqrt(wvnop * wvnom);
    double dpth = (double)S->d[m - 1];
    rho1 = (double)S->rho[m - 1];
    double p = ra * dpth;
    double q = rb * dpth;
    var_(S, p, q, ra, rb, wvno, xka, xkb, dpth, &v);
    dnka(ca, wvno2, gam, gammk, rho1, &v);
    for (int i = 0; i < 5; ++i) {
      double cr = 0.0;
      for (int j = 0; j < 5; ++j) cr = cr + e[j] * ca[j][i];
      ee[i] = cr;
    }
    normc(ee);
    for (int i = 0; i < 5; ++i) e[i] = ee[i];
  }
  if (llw != 1) {
    /* water layer on top: surfdisp96.f:1196-1212 */
    xka = omega / (double)S->a[0];
    wvnop = wvno + xka;
    wvnom = fabs(wvno - xka);
    ra = sqrt(wvnop * wvnom);
    double dpth = (double)S->d[0];
    rho1 = (double)S->rho[0];
    double p = ra * dpth;
    double znul = 1.0e-05;
    var_(S, p, znul, ra, znul, wvno, xka, znul, dpth, &v);
    double w0 = -rho1 * v.w;
    return v.cosp * e[0] + w0 * e[1];
  }
  return e[0];
}

/* dltar: surfdisp96.f:1036-1052 */
static double dltar(sd_ctx* S, double wvno, double omega, int ifunc) {
  S->n_dltar++;
  return (ifunc == 1) ? dltar1(S, wvno, omega) : dltar4(S, wvno, omega);
}

/* half: surfdisp96.f:1022-1032 */
static void half_(sd_ctx* S, double c1, double c2, double* c3, double* del3, double omega, int ifunc) {
  *c3 = 0.5 * (c1 + c2);
  double wvno = omega / *c3;
  *del3 = dltar(S, wvno, omega, ifunc);
}

/* nevill: surfdisp96.f:903-1020.  x,y are 1-based in the Fortran (x(20),y(20)). */
static double nevill(sd_ctx* S, double t, double c1, double c2, double del1, double del2, int ifunc) {
  const double twopi = 2.0 * 3.141592653589793;
  double x[21], y[21];
  double c3, del3;
  int m = 1;
  double omega = twopi / t;
  half_(S, c1, c2, &c3, &del3, omega, ifunc);
  int nev = 1;
  int nctrl = 1;
  for (;;) {
    nctrl = nctrl + 1;
    if (nctrl >= 100) break;
    if (c3 < fmin(c1, c2) || c3 > fmax(c1, c2)) {
      nev = 0;
      half_(S, c1, c2, &c3, &del3, omega, ifunc);
    }
    double s13 = del1 - del3;
    double s32 = del3 - del2;
    if (sgn1(del3) * sgn1(del1) < 0.0) {
      c2 = c3;
      del2 = del3;
    } else {
      c1 = c3;
      del1 = del3;
    }
    if (fabs(c1 - c2) <= 1.e-6 * c1) break;
    if (sgn1(s13) != sgn1(s32)) nev = 0;
    double ss1 = fabs(del1);
    double s1 = (double)0.01f * ss1; /* `0.01*ss1`: default-real literal, :971 */
    double ss2 = fabs(del2);
    double s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || nev == 0) {
      half_(S, c1, c2, &c3, &del3, omega, ifunc);
      nev = 1;
      m = 1;
    } else {
      if (nev == 2) {
        x[m + 1] = c3;
        y[m + 1] = del3;
      } else {
        x[1] = c1;
        y[1] = del1;
        x[2] = c2;
        y[2] = del2;
        m = 1;
      }
      int fail = 0;
      for (int kk = 1; kk <= m; ++kk) {
        int j = m - kk + 1;
        double denom = y[m + 1] - y[j];
        if (fabs(denom) < 1.0e-10 * fabs(y[m + 1])) { fail = 1; break; }
        x[j] = (-y[j] * x[j + 1] + y[m + 1] * x[j]) / denom;
      }
      if (!fail) {
        c3 = x[1];
        double wvno = omega / c3;
        del3 = dltar(S, wvno, omega, ifunc);
        nev = 2;
        m = m + 1;
        if (m > 10) m = 10;
      } else {
        half_(S, c1, c2, &c3, &del3, omega, ifunc);
        nev = 1;
        m = 1;
      }
    }
  }
  return c3;
}

/* getsol: surfdisp96.f:734-827.  Returns iret (1 / -1); c1 updated in place. */
static int getsol(sd_ctx* S, double t1, double* c1_io, double clow, double dc, double cm, float betmx,
                  int ifunc, int ifirst) {
  const double twopi = 2.0 * 3.141592653589793;
  double c1 = *c1_io, c2, del1, del2;
  int idir;
  double omega = twopi / t1;
  double wvno = omega / c1;
  del1 = dltar(S, wvno, omega, ifunc);
  if (ifirst == 1) S->del1st = del1;
  double plmn = sgn1(S->del1st) * sgn1(del1);
  if (ifirst == 1) idir = +1;
  else if (plmn >= 0.0) idir = +1;
  else idir = -1;
  for (;;) {
    if (idir > 0) c2 = c1 + dc;
    else c2 = c1 - dc;
    if (c2 <= clow) {
      idir = +1;
      c1 = clow;
      continue; /* goto 1000 without re-evaluating del1 (:799-803) */
    }
    omega = twopi / t1;
    wvno = omega / c2;
    del2 = dltar(S, wvno, omega, ifunc);
    if (sgn1(del1) != sgn1(del2)) {
      double cn = nevill(S, t1, c1, c2, del1, del2, ifunc);
      c1 = cn;
      *c1_io = c1;
      if (c1 > (double)betmx) return -1;
      return 1;
    }
    c1 = c2;
    del1 = del2;
    if (c1 < cm) break;
    if (c1 >= ((double)betmx + dc)) break;
  }
  *c1_io = c1;
  return -1;
}

/*
 * surfdisp96 (mmode = 0; surfdisp96.f:52-382) and surfdisp_mmodes (mmode = 1;
 * surfdisp96.f:385-701) share everything except: output initialisation (100 vs 0),
 * the k>is,iq==1 start value (:272-282 vs :604-607), and the gvel<0 test (:324-327).
 * iflsph = 0 always (surfmodes.f90:82,94), so `sphere` (:831-899) is not restated.
 * cp, cg are (kmax, mode) column-major.
 */
static void surfdisp_core(sd_ctx* S, int mmode, int iwave, int mode, int igr, int kmax,
                          const double* t, double dphase, double* cp, double* cg, int* ierr) {
  const int mmax = S->mmax;
  double c[NP], cb[NP];
  *ierr = 0;
  for (int i = 0; i < kmax * mode; ++i) {
    cp[i] = mmode ? 0.0 : 100.0;
    cg[i] = mmode ? 0.0 : 100.0;
  }
  const int ifunc = iwave; /* idispl/idispr select exactly one of ifunc=1 (Love), 2 (Rayleigh) */
  float sone0 = 1.500f;
  float ddc0 = (float)dphase; /* :130 */
  float h0 = 0.005f;
  S->llw = 1;
  if (S->b[0] <= 0.0f) S->llw = 2;
  const double one = 1.0e-2;
  int jmn = 1, jsol = 1;
  float betmx = -1.e20f, betmn = 1.e20f;
  for (int i = 1; i <= mmax; ++i) {
    float bi = S->b[i - 1], ai = S->a[i - 1];
    if (bi > 0.01f && bi < betmn) {
      betmn = bi; jmn = i; jsol = 1;
    } else if (bi <= 0.01f && ai < betmn) {
      betmn = ai; jmn = i; jsol = 0;
    }
    if (bi > betmx) betmx = bi;
  }
  float ddc = ddc0, sone = sone0, h = h0;
  if (sone < 0.01f) sone = 2.0f;
  double onea = (double)sone;
  float cc1;
  if (jsol == 0) cc1 = betmn;
  else cc1 = gtsolh(S->a[jmn - 1], S->b[jmn - 1]);
  cc1 = .95f * cc1;
  cc1 = .90f * cc1;
  double cc = (double)cc1;
  double dc = (double)ddc;
  dc = fabs(dc);
  double c1 = cc;
  double cm = cc;
  for (int i = 0; i < kmax; ++i) { cb[i] = 0.0; c[i] = 0.0; }
  int ift = 999;
  for (int iq = 1; iq <= mode; ++iq) {
    const int is = 1, ie = kmax;
    int k;
    int failed = 0;
    for (k = is; k <= ie; ++k) {
      if (k >= ift) { failed = 1; break; }
      double t1 = t[k - 1];
      float t1a, t1b = 0.0f;
      if (igr > 0) {
        t1a = (float)(t1 / (double)(1.f + h));
        t1b = (float)(t1 / (double)(1.f - h));
        t1 = (double)t1a;
      } else {
        t1a = (float)t1;
      }
      double clow = 0.0;
      int ifirst = 0;
      if (k == is && iq == 1) {
        c1 = cc; clow = cc; ifirst = 1;
      } else if (k == is && iq > 1) {
        c1 = c[is - 1] + one * dc; clow = c1; ifirst = 1;
      } else if (k > is && iq > 1) {
        ifirst = 0;
        clow = c[k - 1] + one * dc;
        c1 = c[k - 2];
        if (c1 < clow) c1 = clow;
      } else { /* k > is, iq == 1 */
        ifirst = 0;
        if (mmode) {
          c1 = c[k - 2] - onea * dc;
        } else {
          c1 = cc;
          for (int previd = k - 1; previd >= 1; --previd) {
            if (c[previd - 1] > 0) { c1 = c[previd - 1] - onea * dc; break; }
          }
        }
        clow = cm;
      }
      int iret = getsol(S, t1, &c1, clow, dc, cm, betmx, ifunc, ifirst);
      if (iret == -1) { failed = 1; break; }
      c[k - 1] = c1;
      if (igr > 0) {
        t1 = (double)t1b;
        ifirst = 0;
        clow = cb[k - 1] + one * dc;
        c1 = c1 - onea * dc;
        iret = getsol(S, t1, &c1, clow, dc, cm, betmx, ifunc, ifirst);
        if (iret == -1) {
          c1 = c[k - 1];
          *ierr = 1;
        }
        cb[k - 1] = c1;
      } else {
        c1 = 0.0;
      }
      float cc0 = (float)c[k - 1];
      float cc1f = (float)c1;
      double* cpk = cp + (size_t)(iq - 1) * kmax + (k - 1);
      double* cgk = cg + (size_t)(iq - 1) * kmax + (k - 1);
      if (igr == 0) {
        *cpk = (double)cc0;
      } else {
        float gvel = (1.f / t1a - 1.f / t1b) / (1.f / (t1a * cc0) - 1.f / (t1b * cc1f));
        *cgk = (double)gvel;
        *cpk = (double)cc0;
        if (!mmode && (gvel < 0 || c[k - 1] == 0)) *ierr = 1;
      }
    }
    if (!failed) continue;
    /* labels 1700/1750 (:333-376, :652-695) */
    if (iq <= 1) *ierr = 1; /* iverb(ifunc) bookkeeping collapses to this */
    ift = k;
    for (int i = k; i <= ie; ++i) cg[(size_t)(iq - 1) * kmax + (i - 1)] = 0.0;
    *ierr = 1;
  }
}

/* ------------------------------------------------------------------------------------ */
/* setup_grt's nlvls1 predicate: surfmodes/surfmodes.f90:320-410.  modetype: 1 Rayleigh,
 * 0 Love.  Returns nlvls1, or -1 for "fluid layer below the first" (hard stop :342-345). */
int orc_nlvls1(const double* vp, const double* vs, int n, int modetype) {
  const double eps = (double)1e-6f; /* real(dp),parameter :: eps = 1E-6 */
  int ifs = 0;
  for (int i = 1; i <= n; ++i) {
    if (!(fabs(vs[i - 1]) > eps)) {
      if (i > 1) return -1;
      ifs++;
    }
  }
  double vs1, vss1 = 0.0;
  if (modetype == 1) {
    if (ifs > 0) { vs1 = vp[0]; vss1 = vs[ifs]; } else vs1 = vs[0];
  } else {
    vs1 = vs[ifs];
  }
  int nlvls1 = 0, nlvl1 = 0;
  for (int i = 2; i <= n - 1; ++i) {
    if (i > ifs && vs[i - 1] < vss1) nlvls1++;
    if (vp[i - 1] < vp[i] && vp[i - 1] < vp[i - 2]) {
      if (ifs == 0) {
        if (vs[i - 1] < vs1) nlvl1++;
      } else if (modetype == 1) {
        if (vp[i - 1] < vs1) nlvl1++;
      } else {
        if (vs[i - 1] > 0.) { if (vs[i - 1] < vs1) nlvl1++; }
      }
    }
  }
  if (ifs == 0 || modetype == 0) nlvls1 = nlvl1;
  return nlvls1;
}

/*
 * surfmodes / surfmmodes: surfmodes/surfmodes.f90:39-108, 110-183.
 * thick,vp,vs,rho: n layers (double); freqs: np (Hz); modetype 1 Rayleigh / 0 Love;
 * nmodes <= 0 selects surfmodes->surfdisp96 (mode=1, outputs preset to 100), nmodes >= 1
 * selects surfmmodes->surfdisp_mmodes.  phase, group: np*max(nmodes,1), index
 * ifreq + (imode-1)*np.  Returns 0, or 2 when the column would take the GRT branch
 * (nlvls1 != 0), which this oracle does not restate; outputs are then left untouched.
 * counters[0] += dltar calls, counters[1] += layer steps (may be NULL).
 */
/* Optional: the REFERENCE'S OWN surfdisp96 / surfdisp_mmodes (oracle/_ref/libsurfdisp96_f2c.so, the reference's
 * surfdisp96.f translated to C by oracle/f77toc.py; Fortran calling convention: every argument by reference).  When
 * set, math_mode == 2 routes the solve through them instead of the restatement: used by bench.py's reference arm
 * (cpu_baseline.kind "reference") and by the tests that compare the two.  Work counters are then not available. */
typedef void (*ref_surfdisp_fn)(void* thk, void* vp, void* vs, void* rho, void* nlayer, void* iflsph, void* iwave, void* mode,
                                void* igr, void* kmax, void* t, void* dphase, void* cp, void* cg, void* ierr);
static ref_surfdisp_fn g_ref96 = 0, g_refmm = 0;
void orc_set_reference_solver(void* f96, void* fmm) { g_ref96 = (ref_surfdisp_fn)f96; g_refmm = (ref_surfdisp_fn)fmm; }

int orc_surfmodes(const double* thick, const double* vp, const double* vs, const double* rho, int n,
                  const double* freqs, int np, int modetype, int phaseGroup, int nmodes, double dc,
                  int math_mode, double* phase, double* group, int* ierr, int64_t* counters) {
  if (n > NL || np > NP || n < 1) return 3;
  int lv = orc_nlvls1(vp, vs, n, modetype);
  *ierr = 0;
  if (lv != 0) return 2;
  if (math_mode == 2) {
    if (!g_ref96 || !g_refmm) return 9;
    float th4[NL], a4[NL], b4[NL], r4[NL];
    memset(th4, 0, sizeof th4); memset(a4, 0, sizeof a4); memset(b4, 0, sizeof b4); memset(r4, 0, sizeof r4);
    for (int i = 0; i < n; ++i) { th4[i] = (float)thick[i]; a4[i] = (float)vp[i]; b4[i] = (float)vs[i]; r4[i] = (float)rho[i]; }
    double tt[NP];
    memset(tt, 0, sizeof tt);
    for (int i = 0; i < np; ++i) tt[i] = 1 / freqs[i];
    int iflsph = 0, iwave = (modetype == 1) ? 2 : 1, mode = nmodes <= 0 ? 1 : nmodes, igr = phaseGroup, kmax = np, ie = 0;
    double dph = dc;
    (nmodes <= 0 ? g_ref96 : g_refmm)(th4, a4, b4, r4, &n, &iflsph, &iwave, &mode, &igr, &kmax, tt, &dph, phase, group, &ie);
    *ierr = ie;
    return 0;
  }
  sd_ctx S;
  memset(&S, 0, sizeof S);
  S.mmax = n;
  S.math_mode = math_mode;
  for (int i = 0; i < n; ++i) {
    S.d[i] = (float)thick[i]; /* real(thick,4) ... surfmodes.f90:81-83 */
    S.a[i] = (float)vp[i];
    S.b[i] = (float)vs[i];
    S.rho[i] = (float)rho[i];
  }
  double t[NP];
  for (int i = 0; i < np; ++i) t[i] = 1 / freqs[i]; /* dble(1/freqs) */
  int iwave = (modetype == 1) ? 2 : 1;
  if (nmodes <= 0) surfdisp_core(&S, 0, iwave, 1, phaseGroup, np, t, dc, phase, group, ierr);
  else surfdisp_core(&S, 1, iwave, nmodes, phaseGroup, np, t, dc, phase, group, ierr);
  if (counters) { counters[0] += S.n_dltar; counters[1] += S.n_layer; }
  return 0;
}

/* thin test hooks for the portable math functions */
double orc_mct_exp(double x) { return mct_exp(x); }
void orc_mct_sincos(double x, double* s, double* c) { mct_sincos(x, s, c); }
double orc_mct_pow025(double x) { return mct_pow025(x); }
float orc_gtsolh(float a, float b) { return gtsolh(a, b); }
