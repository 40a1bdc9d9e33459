/*
 * oracle/grt_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the generalized reflection/transmission branch of the reference's `surfmodes`
 * (columns with a low-velocity layer, surfmodes/surfmodes.f90:84-87,96-99): RayleighModes / LoveModes
 * (surfmodes.f90:185-306), setup_grt (:320-450), SearchRayleigh with FundaMode / StMode (SearchRayleigh.f90:1-84,
 * 283-470, 781-933), SearchLove with its FundaMode (SearchLove.f90), C_Interval / N_cf (C_interval.f90),
 * C_Interval_L / N_cf_L (C_interval_L.f90), startl, SecFunSurf, SecFunSt, Stoneley, EinvE, EinvE_f, propup, propdn_f
 * (Rayleigh.f90), SecFuns_L, EinvE_L, propdn_L, propup_L (Love.f90), bisecim, det3, sort (util.f90), csq (GRT.f90).
 * Only what `surfmodes` reaches is restated (allmodes = 0: the fundamental / Stoneley mode per frequency);
 * `surfmmodes` prints "not supported yet" for such columns (surfmodes.f90:153,165).
 *
 * PARITY STATUS: pinned end to end -- orc_grt_modes returns, bit for bit, the phase velocities of the reference's own setup_grt,
 * C_Interval[_L], FundaMode (+ CR0_Finder) / StMode (+ St_Finder), startl, SecFunSurf / SecFunSt / SecFuns_L and bisecim,
 * translated mechanically by oracle/f90toc_love.py, for Love and for Rayleigh with and without a water layer; each routine
 * is also compared alone; group velocities (second search + CalGroup, translated) likewise (tests/test_oracle_grt.py).
 * The reference ships no test, golden value or compiled object for these files and
 * no Fortran compiler exists in this image (oracle/f77toc.py translates FORTRAN 77, not this Fortran 90).  The
 * restatement is pinned by physics only (tests/test_oracle_grt.py: the roots it returns are zeros of an independent
 * 50-digit propagator-matrix secular function and are the lowest mode) and by line-by-line reading; the complex
 * primitives alone are held bit for bit to what gcc emits for double complex under -fcx-fortran-rules and to libm's cexp.
 *
 * Conventions kept from the Fortran: default-real literals are float-rounded (0.12, 0.90, 3.1415926, 0.005, 1E-6, 1.1);
 * complex multiplication is (ac-bd, ad+bc), complex division is the range-reduced form gfortran emits
 * (-fcx-fortran-rules); sqrt only ever sees dcmplx(real) arguments, i.e. it is a real square root on one axis;
 * exp(x+iy) = e^x (cos y, sin y); matmul accumulates over k ascending.  Where the Fortran reads an undefined value
 * (the Love `check` on an unset `c`, array overruns of vvv(20000)) the restatement says so at the spot.
 * math_mode: 0 = libm, 1 = the portable functions of mctomo_b200/csrc/mct_math.h (what the CUDA kernels use).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../mctomo_b200/csrc/mct_math.h"

#define GNL 200     /* layers */
#define GNV 20000   /* vvv / ccc length (C_interval.f90:5,14) */

typedef struct { double re, im; } cx;
static inline cx CX(double a, double b) { cx z = {a, b}; return z; }
static inline cx cadd(cx a, cx b) { return CX(a.re + b.re, a.im + b.im); }
static inline cx csub(cx a, cx b) { return CX(a.re - b.re, a.im - b.im); }
static inline cx cneg(cx a) { return CX(-a.re, -a.im); }
static inline cx cmul(cx a, cx b) { return CX(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cx crs(double r, cx a) { return CX(r * a.re, r * a.im); }   /* real * complex */
static inline cx cdr(cx a, double r) { return CX(a.re / r, a.im / r); }   /* complex / real */
static inline cx cdiv(cx a, cx b) { /* gcc expand_complex_div_wide (Fortran rules) */
  double ratio, div, tr, ti;
  if (fabs(b.re) < fabs(b.im)) {
    ratio = b.re / b.im; div = (b.re * ratio) + b.im;
    tr = (a.re * ratio) + a.im; ti = (a.im * ratio) - a.re;
  } else {
    ratio = b.im / b.re; div = (b.im * ratio) + b.re;
    tr = (a.im * ratio) + a.re; ti = a.im - (a.re * ratio);
  }
  return CX(tr / div, ti / div);
}
static const cx IC = {1.0, 0.0};

typedef struct {
  int n, modetype, math_mode;
  double d[GNL + 1], vp[GNL + 1], vs[GNL + 1], rho[GNL + 1], mu[GNL + 1]; /* 1-based */
  double v[2 * GNL + 2];
  int nv;
  double mu0, vsy, vs1, vsm, vss1;
  int ifs, no_lvl, no_lvl_fl, nlvl1, nlvls1, lvlast, L1;
  int lvls[GNL / 2 + 3];
  double dc, dc2, dcm, w, tol, smin;
  int index_a, ll;
  /* module variables of Rayleigh.f90 / Love.f90 that survive between calls */
  cx cp[2], cs[2], la[2];
  int64_t n_secfun, n_layer; /* instrumentation: secular-function evaluations and interface steps inside them */
  double* ccc; double* vvv;
  int overflow;
} grt_t;

static inline double g_exp(const grt_t* G, double x) { return G->math_mode ? mct_exp(x) : exp(x); }
static inline void g_sincos(const grt_t* G, double x, double* s, double* c) {
  if (G->math_mode) mct_sincos(x, s, c); else { *s = sin(x); *c = cos(x); }
}
static inline cx g_cexp(const grt_t* G, cx z) {
  double e = g_exp(G, z.re), s, c;
  if (z.im == 0.0) { s = z.im; c = 1.0; } else g_sincos(G, z.im, &s, &c);
  return CX(e * c, e * s);
}
/* csq: GRT.f90:111-116.  sqrt(dcmplx(1-(c/vel)**2)): real argument, branch cut as csqrt (x<0 -> +i sqrt(-x)). */
static inline cx csq(double c, double vel) {
  double t = c / vel;
  double x = 1 - t * t;
  return x >= 0 ? CX(sqrt(x), 0.0) : CX(0.0, sqrt(-x));
}

/* test entry: the complex primitives above, so that tests/test_oracle_grt.py can hold them to what the compiler itself emits
 * for double complex under -fcx-fortran-rules (gfortran's rules, same middle end) and to libm's cexp / csqrt.
 * op 0: a*b, 1: a/b, 2: exp(a) (libm mode), 3: csq(a.re, b.re). */
void orc_grt_cprim(int op, double are, double aim, double bre, double bim, double* out2) {
  grt_t G;
  G.math_mode = 0;
  cx a = CX(are, aim), b = CX(bre, bim), r;
  if (op == 0) r = cmul(a, b);
  else if (op == 1) r = cdiv(a, b);
  else if (op == 2) r = g_cexp(&G, a);
  else r = csq(are, bre);
  out2[0] = r.re; out2[1] = r.im;
}

static int cmp_d(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
/* util.f90 sort(arr,n,1): a Shell sort; any ascending sort yields the same array */
static void sort_up(double* a, int n) { qsort(a, (size_t)n, sizeof(double), cmp_d); }

/* setup_grt: surfmodes.f90:320-450 (after init_grt, GRT.f90:70-109).  Returns -1 for a fluid layer below the first. */
static int setup_grt(grt_t* G, const double* thick, const double* vp, const double* vs, const double* rho, int n, int modetype,
                     double dc, double dcm, double dc2) {
  const double eps = (double)1e-6f; /* m_surfmodes eps = 1E-6 */
  memset(G->lvls, 0, sizeof G->lvls);
  G->n = n; G->modetype = modetype;
  G->dc = dc; G->dc2 = dc2; G->dcm = dcm;
  G->ifs = 0; G->no_lvl = G->no_lvl_fl = G->nlvl1 = G->nlvls1 = 0; G->L1 = 0; G->ll = 0;
  G->vss1 = 0; G->vs1 = 0; G->vsm = 0;
  for (int i = 1; i <= n; ++i) { G->d[i] = thick[i - 1]; G->vp[i] = vp[i - 1]; G->vs[i] = vs[i - 1]; G->rho[i] = rho[i - 1]; }
  int idx = 0;
  memset(G->v, 0, sizeof G->v);
  for (int i = 1; i <= n; ++i) {
    if (fabs(G->vs[i]) > eps) { idx++; G->v[idx] = G->vs[i]; }
    else { if (i > 1) return -1; G->ifs++; }
    idx++; G->v[idx] = G->vp[i];
  }
  sort_up(G->v + 1, idx);
  G->nv = idx;
  int cnt = 0; double mu0 = 0;
  for (int i = 1; i <= n; ++i) {
    double mu = G->rho[i] * (G->vs[i] * G->vs[i]);
    if (fabs(mu) > eps) { cnt++; mu0 = mu0 + mu; }
    G->mu[i] = mu;
  }
  mu0 = mu0 / cnt;
  for (int i = 1; i <= n; ++i) G->mu[i] = G->mu[i] / mu0;
  G->mu0 = mu0;
  G->vsy = G->vs[1];
  for (int i = 2; i <= n; ++i) if (G->vs[i] > G->vsy) G->vsy = G->vs[i];
  if (modetype == 1) {
    if (G->ifs > 0) { G->vs1 = G->vp[1]; G->vss1 = G->vs[G->ifs + 1]; } else G->vs1 = G->vs[1];
    G->vsm = G->v[1];
  } else {
    G->vs1 = G->vs[1 + G->ifs];
    G->vsm = G->vs[G->ifs + 1];
    for (int i = G->ifs + 2; i <= n; ++i) if (G->vs[i] < G->vsm) G->vsm = G->vs[i];
  }
  for (int i = 2; i <= n - 1; ++i) {
    if (i > G->ifs && G->vs[i] < G->vss1) G->nlvls1++;
    if (G->vp[i] < G->vp[i + 1] && G->vp[i] < G->vp[i - 1]) {
      G->no_lvl++; G->lvls[G->no_lvl] = i;
      if (G->ifs == 0) { if (G->vs[i] < G->vs1) G->nlvl1++; }
      else if (modetype == 1) { if (G->vp[i] < G->vs1) G->nlvl1++; }
      else { if (G->vs[i] > 0.) { if (G->vs[i] < G->vs1) G->nlvl1++; } else G->no_lvl_fl++; }
    }
  }
  if (G->ifs == 0 || modetype == 0) G->nlvls1 = G->nlvl1;
  G->lvls[G->no_lvl + 1] = modetype == 1 ? 1 : 1 + G->ifs;
  if (G->no_lvl == 0) G->lvlast = 1;
  else {
    G->lvlast = G->lvls[G->no_lvl];
    for (int i = 1; i <= G->no_lvl - 1; ++i)
      for (int j = 1; j <= G->no_lvl - i; ++j)
        if (G->vp[G->lvls[j]] > G->vp[G->lvls[j + 1]]) { int k = G->lvls[j]; G->lvls[j] = G->lvls[j + 1]; G->lvls[j + 1] = k; }
  }
  if (G->ifs + 2 > G->lvlast) G->lvlast = G->ifs + 2;
  for (int i = 1; i <= G->no_lvl; ++i)
    if (G->lvls[i] > G->ifs && G->vs[G->lvls[i]] < G->vs1) { G->L1 = i; break; }
  return 0;
}

/* startl: Rayleigh.f90:62-105 */
static void startl(grt_t* G, double c) {
  int sl = G->n;
  double vk = G->w / c, su = 0;
  for (int j = G->lvlast; j <= G->n - 1; ++j) {
    if (c < G->vs[j]) {
      su = su + vk * csq(c, G->vs[j]).re * G->d[j];
      if (su > 46.0) { sl = j; break; } /* expo = 46d0 */
    } else su = 0;
  }
  G->ll = sl;
}

/* EinvE: Rayleigh.f90:358-394.  a (k=1: the layer below interface j) and, for iq == 0, b (k=0: the layer above). */
typedef struct { cx m[4][4]; } m44;
static void einve_a(grt_t* G, int j, double c, m44* A) { /* case(1) */
  double ap = G->vp[j + 1], as = G->vs[j + 1], am = G->mu[j + 1];
  G->cp[1] = csq(c, ap); G->cs[1] = csq(c, as);
  double t = c / as;
  cx xi = CX(1 - t * t / 2., 0.0);
  cx cpk = G->cp[1], csk = G->cs[1];
  cx col1[4] = {IC, cpk, crs(-am, cpk), crs(-am, xi)};
  cx col2[4] = {csk, IC, crs(-am, xi), crs(-am, csk)};
  for (int r = 0; r < 4; ++r) { A->m[r][0] = col1[r]; A->m[r][1] = col2[r]; A->m[r][2] = col1[r]; A->m[r][3] = col2[r]; }
  for (int r = 1; r <= 2; ++r) for (int q = 2; q <= 3; ++q) A->m[r][q] = cneg(A->m[r][q]); /* pp=>a44(2:3,3:4); pp=-pp */
}
static void einve_b(grt_t* G, int j, double c, m44* B) { /* case(0) */
  double ap = G->vp[j], as = G->vs[j], am = G->mu[j];
  G->cp[0] = csq(c, ap); G->cs[0] = csq(c, as);
  double t = c / as;
  cx xi = CX(1 - t * t / 2., 0.0);
  cx cpk = G->cp[0], csk = G->cs[0];
  cx row1[4] = {IC, cneg(cdiv(xi, cpk)), cneg(cdiv(CX(1.0, 0.0), crs(am, cpk))), cdr(IC, am)};
  cx row2[4] = {cneg(cdiv(xi, csk)), IC, cdr(IC, am), cneg(cdiv(CX(1.0, 0.0), crs(am, csk)))};
  for (int q = 0; q < 4; ++q) { B->m[0][q] = row1[q]; B->m[1][q] = row2[q]; B->m[2][q] = row1[q]; B->m[3][q] = row2[q]; }
  for (int r = 2; r <= 3; ++r) for (int q = 1; q <= 2; ++q) B->m[r][q] = cneg(B->m[r][q]); /* pp=>b44(3:4,2:3); pp=-pp */
  cx den = crs(2.0, csub(IC, xi)); /* 2*(1-xi) */
  for (int r = 0; r < 4; ++r) for (int q = 0; q < 4; ++q) B->m[r][q] = cdiv(B->m[r][q], den);
}
static void einve(grt_t* G, int j, double c, int iq, m44* E) {
  m44 A, B;
  einve_a(G, j, c, &A);
  if (iq == 0) {
    einve_b(G, j, c, &B);
    for (int r = 0; r < 4; ++r) for (int q = 0; q < 4; ++q) { /* matmul(b44,a44) */
      cx s = CX(0, 0);
      for (int k = 0; k < 4; ++k) s = cadd(s, cmul(B.m[r][k], A.m[k][q]));
      E->m[r][q] = s;
    }
  } else *E = A;
}
typedef struct { cx m[2][2]; } m22;
static m22 mm22(const m22* a, const m22* b) {
  m22 r;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r.m[i][j] = cadd(cadd(CX(0, 0), cmul(a->m[i][0], b->m[0][j])), cmul(a->m[i][1], b->m[1][j]));
  return r;
}
static m22 sub22(const m44* E, int r0, int c0) { m22 r; for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r.m[i][j] = E->m[r0 + i][c0 + j]; return r; }
static m22 add22(const m22* a, const m22* b) { m22 r; for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r.m[i][j] = cadd(a->m[i][j], b->m[i][j]); return r; }
static m22 inv2(const m22* a) { /* Rayleigh.f90:26-32 */
  cx det = csub(cmul(a->m[0][0], a->m[1][1]), cmul(a->m[1][0], a->m[0][1]));
  m22 r;
  r.m[0][0] = cdiv(a->m[1][1], det); r.m[1][0] = cdiv(cneg(a->m[1][0]), det);
  r.m[0][1] = cdiv(cneg(a->m[0][1]), det); r.m[1][1] = cdiv(a->m[0][0], det);
  return r;
}
/* la = exp(-d*vk*[p,s]) */
static void set_la(grt_t* G, double d, double vk, cx p, cx s) {
  double f = -d * vk;
  G->la[0] = g_cexp(G, crs(f, p)); G->la[1] = g_cexp(G, crs(f, s));
}

/* propup: Rayleigh.f90:449-491; returns Rdu(:,:,j1) (only that one is consumed by the callers restated here) */
static m22 propup(grt_t* G, double c, int j2, int j1) {
  double vk = G->w / c;
  m44 E;
  einve(G, j2, c, 0, &E);
  m22 t = sub22(&E, 0, 0);
  m22 a22 = inv2(&t);
  set_la(G, G->d[j2], vk, G->cp[0], G->cs[0]);
  for (int k = 0; k < 2; ++k) for (int i = 0; i < 2; ++i) a22.m[i][k] = cmul(a22.m[i][k], G->la[k]);
  m22 e31 = sub22(&E, 2, 0);
  m22 Rdu = mm22(&e31, &a22);
  G->n_layer++;
  for (int j = j2 - 1; j >= j1; --j) {
    einve(G, j, c, 0, &E);
    m22 b22 = Rdu;
    set_la(G, G->d[j + 1], vk, G->cp[1], G->cs[1]);
    for (int k = 0; k < 2; ++k) for (int q = 0; q < 2; ++q) b22.m[k][q] = cmul(b22.m[k][q], G->la[k]);
    m22 e11 = sub22(&E, 0, 0), e13 = sub22(&E, 0, 2), e31b = sub22(&E, 2, 0), e33 = sub22(&E, 2, 2);
    m22 p = mm22(&e13, &b22);
    m22 s = add22(&e11, &p);
    a22 = inv2(&s);
    set_la(G, G->d[j], vk, G->cp[0], G->cs[0]);
    for (int k = 0; k < 2; ++k) for (int i = 0; i < 2; ++i) a22.m[i][k] = cmul(a22.m[i][k], G->la[k]);
    m22 p2 = mm22(&e33, &b22);
    b22 = add22(&e31b, &p2);
    Rdu = mm22(&b22, &a22);
    G->n_layer++;
  }
  return Rdu;
}

/* SecFunSurf: Rayleigh.f90:125-154 (isurf = 0) */
static double secfun_surf(grt_t* G, double c, double* imf) {
  G->n_secfun++;
  m22 Rdu = propup(G, c, G->ll - 1, 1);
  m44 A;
  einve(G, 0, c, 1, &A);
  m22 b22 = sub22(&A, 2, 2);
  for (int k = 0; k < 2; ++k) for (int i = 0; i < 2; ++i) b22.m[i][k] = cmul(b22.m[i][k], G->la[k]);
  m22 e31 = sub22(&A, 2, 0);
  m22 p = mm22(&b22, &Rdu);
  m22 a22 = add22(&e31, &p);
  cx dsp = csub(cmul(a22.m[0][0], a22.m[1][1]), cmul(a22.m[0][1], a22.m[1][0]));
  *imf = dsp.im;
  return dsp.re;
}

/* EinvE_f (iq = 1 only is reached): Rayleigh.f90:238-266 */
static void einve_f1(grt_t* G, int j, double c, m22* a) {
  double ap = G->vp[j + 1];
  G->cp[1] = csq(c, ap);
  cx xi = CX(G->rho[j + 1] * (c * c) / (2. * G->mu0), 0.0);
  a->m[0][0] = IC; a->m[1][0] = cdiv(xi, G->cp[1]);
  a->m[0][1] = IC; a->m[1][1] = cneg(a->m[1][0]);
}
/* Stoneley + SecFunSt for ONE water layer (ifs = 1; deeper fluid stacks stop in setup_grt): Rayleigh.f90:157-221 */
static double secfun_st(grt_t* G, double c, double* imf) {
  G->n_secfun++;
  const int ifs = G->ifs;
  double vk = G->w / c;
  /* propdn_f(c,1,ifs): Rud(1,1,0) = exp(-d(1)*vk*csq(c,vp(1))); the loop j=1..ifs-1 is empty */
  cx rud0 = g_cexp(G, crs(-G->d[1] * vk, csq(c, G->vp[1])));
  m22 Rdu = propup(G, c, G->ll - 1, ifs + 1);
  m44 A;
  einve(G, ifs, c, 1, &A);
  set_la(G, G->d[ifs + 1], vk, G->cp[1], G->cs[1]);
  cx b33[3][3], a33[3][3], r33[3][3];
  memset(b33, 0, sizeof b33);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) b33[i][j] = cmul(Rdu.m[i][j], G->la[i]); /* row k scaled by la(k) */
  m22 b22;
  einve_f1(G, ifs - 1, c, &b22); /* sets cp(1) to the water layer's */
  b33[2][2] = cmul(rud0, g_cexp(G, crs(-G->d[ifs] * vk, G->cp[1])));
  for (int i = 0; i < 3; ++i) { a33[i][0] = A.m[1 + i][2]; a33[i][1] = A.m[1 + i][3]; }
  a33[0][2] = cneg(b22.m[0][0]); a33[1][2] = cneg(CX(0, 0)); a33[2][2] = cneg(b22.m[1][0]);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    cx s = CX(0, 0);
    for (int k = 0; k < 3; ++k) s = cadd(s, cmul(a33[i][k], b33[k][j]));
    r33[i][j] = s;
  }
  for (int i = 0; i < 3; ++i) { a33[i][0] = A.m[1 + i][0]; a33[i][1] = A.m[1 + i][1]; }
  a33[0][2] = cneg(b22.m[0][1]); a33[1][2] = cneg(CX(0, 0)); a33[2][2] = cneg(b22.m[1][1]);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a33[i][j] = cadd(a33[i][j], r33[i][j]);
  /* det3: util.f90 */
  cx q1 = csub(cmul(a33[1][1], a33[2][2]), cmul(a33[1][2], a33[2][1]));
  cx q2 = csub(cmul(a33[1][0], a33[2][2]), cmul(a33[1][2], a33[2][0]));
  cx q3 = csub(cmul(a33[1][0], a33[2][1]), cmul(a33[1][1], a33[2][0]));
  cx dsp = cadd(csub(cmul(a33[0][0], q1), cmul(a33[0][1], q2)), cmul(a33[0][2], q3));
  *imf = dsp.im;
  return dsp.re;
}

/* Love.f90: EinvE_L, propdn_L (the part reached with lay = 1+ifs), propup_L, SecFuns_L */
static void einve_L(grt_t* G, int j, double c, int iq, m22* out) {
  m22 a, b;
  double as = G->vs[j + 1], am = G->mu[j + 1];
  G->cs[1] = csq(c, as);
  a.m[0][1] = IC; a.m[1][1] = crs(am, G->cs[1]);
  a.m[0][0] = IC; a.m[1][0] = cneg(a.m[1][1]);
  if (iq == 0) {
    as = G->vs[j]; am = G->mu[j];
    G->cs[0] = csq(c, as);
    b.m[0][0] = crs(am, G->cs[0]); b.m[1][0] = b.m[0][0];
    b.m[0][1] = cneg(IC); b.m[1][1] = IC;
    cx den = crs(2., b.m[0][0]);
    for (int i = 0; i < 2; ++i) for (int k = 0; k < 2; ++k) b.m[i][k] = cdiv(b.m[i][k], den);
    *out = mm22(&b, &a);
  } else *out = a;
}
static double secfun_L(grt_t* G, double c, double* imf) {
  G->n_secfun++;
  const int lay = 1 + G->ifs;
  double vk = G->w / c;
  m22 a22;
  /* propdn_L(c,1+ifs,lay): loop empty */
  einve_L(G, G->ifs, c, 1, &a22);
  cx la2 = g_cexp(G, crs(-G->d[1 + G->ifs] * vk, G->cs[1]));
  cx rud = cdiv(cmul(cneg(a22.m[1][1]), la2), a22.m[1][0]);
  /* propup_L(c,ll-1,lay) */
  int j2 = G->ll - 1;
  einve_L(G, j2, c, 0, &a22);
  cx td = cdiv(g_cexp(G, crs(-G->d[j2] * vk, G->cs[0])), a22.m[0][0]);
  cx rdu = cmul(a22.m[1][0], td);
  G->n_layer++;
  for (int j = j2 - 1; j >= lay; --j) {
    einve_L(G, j, c, 0, &a22);
    cx r = cmul(g_cexp(G, crs(-G->d[j + 1] * vk, G->cs[1])), rdu);
    td = cdiv(g_cexp(G, crs(-G->d[j] * vk, G->cs[0])), cadd(a22.m[0][0], cmul(a22.m[0][1], r)));
    rdu = cmul(cadd(a22.m[1][0], cmul(a22.m[1][1], r)), td);
    G->n_layer++;
  }
  cx dsp = csub(IC, cmul(rud, rdu));
  *imf = dsp.re;
  return dsp.im;
}

typedef double (*secf)(grt_t*, double, double*);

/* bisecim: util.f90:68-122 */
static double bisecim(grt_t* G, secf f, double x1, double x2, double f1, double f2, int* iq) {
  double x[4], y[4], xt1, xt2, fa, dx, dxt, u1, u2, imf = 0, imf2 = 0;
  const double smin2 = G->smin * G->smin * 2.;
  int nc = 0;
  dx = fabs(x1 - x2);
  fa = f1;
  x[1] = x1; x[2] = x2; y[1] = fa; y[2] = f2;
  xt1 = (x[1] + x[2]) / 2.;
  for (;;) {
    x[3] = (x[1] + x[2]) / 2.;
    y[3] = f(G, x[3], &imf);
    u1 = (x[2] - x[1]) / (y[2] - y[1]);
    u2 = (x[2] - x[3]) / (y[2] - y[3]);
    xt2 = x[1] - y[1] * (u1 - y[2] * ((u2 - u1) / (y[3] - y[1])));
    dxt = fabs(xt2 - xt1);
    dx = dx / 2.;
    if (dx < dxt) dxt = dx;
    if (dxt < G->tol) {
      u1 = f(G, xt2, &imf2);
      u2 = y[3];
      if (u1 * u1 + imf2 * imf2 < smin2 || u2 * u2 + imf * imf < smin2) { *iq = 0; return fabs(u1) < fabs(u2) ? xt2 : x[3]; }
      *iq = -1;
      return 0.;
    }
    xt1 = xt2;
    if (fa * y[3] < 0) { x[2] = x[3]; y[2] = y[3]; } else { x[1] = x[3]; y[1] = y[3]; }
    nc++;
    if (nc >= 1000) { *iq = -1; return 0; }
  }
}

/* N_cf: C_interval.f90:178-199; N_cf_L: C_interval_L.f90.  Only the real part of the complex sum is used. */
static double n_cf(const grt_t* G, double c, int love) {
  const double pi_s = love ? 3.1415926535897932 : (double)3.1415926f; /* N_cf has its own single-precision pi; N_cf_L uses m_GRT's */
  double sum = 0.;
  for (int i = 1; i <= G->n - 1; ++i) {
    double yp = 0., ys = 0.;
    if (!love) { double t = c / G->vp[i]; double x = t * t - 1; yp = x >= 0 ? sqrt(x) : 0.; }
    if (G->vs[i] > 0.0) { double t = c / G->vs[i]; double x = t * t - 1; ys = x >= 0 ? sqrt(x) : 0.; }
    sum = sum + 2.0 * (love ? ys : (yp + ys)) * G->d[i] / c;
  }
  return G->w / (2.0 * pi_s) * sum;
}

/* C_Interval (C_interval.f90:1-176) / C_Interval_L (C_interval_L.f90): the trial phase velocities ccc(1..ncc), im1 */
static void c_interval(grt_t* G, int love, int* ncc, int* im1) {
  double* vvv = G->vvv; double* ccc = G->ccc; /* 1-based */
  const double pi_c = love ? 3.1415926535897932 : (double)3.1415926f;
  const double eps = 1e-10; /* m_GRT eps */
  const double freq = G->w / (2.0 * pi_c);
  const int NN = (int)lround(n_cf(G, G->vsy, love) - n_cf(G, G->vsm, love));
  int index0;
  const double lowv = love ? G->vsm : G->v[1];
#define PUSH(val) do { if (index0 >= GNV - 1) { G->overflow = 1; } else { index0++; vvv[index0] = (val); } } while (0)
  if (freq < (double)0.12f || NN < 2) {
    int n0 = (NN + 1) * (love ? 512 : 1024);
    if (n0 > GNV - 200) { G->overflow = 1; n0 = GNV - 200; } /* the Fortran would overrun vvv(20000) here */
    const double off = love ? 0.01 : 0.1;
    for (int i = 1; i <= n0; ++i) vvv[i] = lowv - off + (double)i * (G->vsy - lowv + off) / (double)n0;
    index0 = n0;
    for (int i = 1; i <= 100; ++i) PUSH(G->vsy * (1. - .008 * i));
  } else {
    index0 = 0;
    double c1 = lowv, c2 = G->vsy, c0 = c2, dc = 0, dN;
    PUSH(c1);
    int NNc = NN;
    long guard = 0;
    while (c0 > c1 && !G->overflow) {
      if (NNc > 0) dc = (c2 - c1) / (NNc);
      c0 = c2 - dc;
      for (;;) {
        dN = n_cf(G, c2, love) - n_cf(G, c0, love);
        if (dN < .5) { PUSH(c0); NNc = NNc - 1; c2 = c0; break; }
        c0 = (c2 + c0) / 2.0;
        if (++guard > 2000000) { G->overflow = 1; break; }
      }
    }
    int ij = 1;
    while (ij <= G->nv && G->v[ij] <= G->vsy) ij++;
    for (int i = 1; i <= ij - 2; ++i) {
      double c01 = G->v[i];
      int ii = 1;
      for (int j = 1; j <= G->n; ++j) if (fabs(G->vs[j] - c01) < eps || fabs(G->vp[j] - c01) < eps) ii = j;
      double hi = G->d[ii];
      double Ni = 2.0 * freq * hi / c01 + eps;
      int nj = (int)floor(Ni);
      for (int j = 1; j <= nj; ++j) {
        double q = (double)j / Ni;
        double c00 = c01 / sqrt(1.0 - q * q);
        if (c00 <= G->vsy) PUSH(c00);
      }
    }
    sort_up(vvv + 1, index0);
    for (int j = 1; j <= 2; ++j) {
      int intemp = index0;
      for (int i = 1; i <= intemp - 1; ++i) PUSH((vvv[i] + vvv[i + 1]) / 2.0);
    }
    for (int i = 1; i <= 100; ++i) PUSH(G->vsy - (double)i / 100. * 0.1);
  }
  for (int i = 1; i <= 10; ++i) PUSH(G->vsy - (i * 10) * G->tol);
  PUSH(G->vsy - 3.0 * G->tol);
  PUSH(G->vs1 + 4 * G->tol);
  PUSH(G->vs1 - 4 * G->tol);
#undef PUSH
  sort_up(vvv + 1, index0);
  for (int i = 1; i <= index0; ++i) {
    if (vvv[i] > G->vsm) {
      int ii = index0;
      index0 = index0 - (i - 1) + 1;
      memmove(&vvv[2], &vvv[i], sizeof(double) * (size_t)(ii - i + 1));
      vvv[1] = G->vsm + G->tol;
      break;
    }
  }
  int ii = index0;
  while (ii >= 1 && vvv[ii] >= G->vsy) ii--;
  const double tol0 = 10 * G->tol;
  ccc[1] = vvv[1];
  index0 = 1;
  for (int i = 2; i <= ii; ++i) {
    if (vvv[i] - vvv[i - 1] < tol0) continue;
    index0++;
    ccc[index0] = vvv[i];
  }
  *ncc = index0;
  if (!love) {
    *im1 = 0;
    for (int i = 1; i <= index0; ++i) if (ccc[i] >= G->vs1) { *im1 = i; break; }
    if (*im1 == 0) *im1 = index0;
  } else { /* C_Interval_L leaves im1 as the caller set it (0) when nothing exceeds vs1; SearchLove's FundaMode never reads it */
    *im1 = 0;
    for (int i = 1; i <= index0; ++i) if (ccc[i] > G->vs1) { *im1 = i; break; }
  }
}

/* CR0_Finder: SearchRayleigh.f90:895-933 */
static double cr0_finder(double v1, double v2) {
  const double tol = 1e-7;
  double c = 0.8 * v1, R, DR;
  for (int it = 0; it < 100000; ++it) {
    double ps = 1.0 / v1, pp = 1.0 / v2, p = 1.0 / c;
    double p2 = p * p, ps2 = ps * ps, pp2 = pp * pp;
    double sps = sqrt(p2 - ps2), spp = sqrt(p2 - pp2);
    double t = ps2 - 2.0 * p2;
    R = t * t - 4.0 * p2 * sps * spp;
    DR = p2 * (8.0 * p * (ps2 - 2.0 * p2) + 8.0 * p * sps * spp + 4.0 * (p2 * p) * (spp / sps + sps / spp));
    c = c - R / DR;
    if (v1 - c < tol || v2 - c < tol || c != c) break;
    if (fabs(R / DR) < tol) break;
  }
  return c; /* CRo = c: the reset to 0.8*v1 that follows only changes the local c */
}

/* St_Finder: SearchRayleigh.f90:846-893.  Returns 0 with *ok = 0 where the Fortran STOPs ('Wrong input for Stoneley mode!'). */
static double getSt(const grt_t* G, int n, double x) {
  double t = x / G->vp[n], a = 1 - t * t;
  t = x / G->vp[n + 1]; double b = 1 - t * t;
  t = x / G->vs[n + 1]; double c1 = t * t * (G->rho[n] / G->rho[n + 1]);
  t = G->vs[n + 1] / x; double c2 = t * t;
  double c = 1 - 1. / c2;
  double u = 1 + c;
  return c1 * sqrt(b / a) + c2 * (u * u - 4 * sqrt(b * c));
}
static double st_finder(const grt_t* G, int n, double cst_in, int* ok) {
  const double tolSt = 1e-7;
  double c2 = G->vp[n] < G->vs[n + 1] ? G->vp[n] : G->vs[n + 1];
  double c1 = c2 * .8;
  c2 = c2 - (c2 - c1) / 1e4;
  double a2 = getSt(G, n, c2), a1 = getSt(G, n, c1), dc = c2 - c1, c0 = (c2 + c1) / 2., a0;
  *ok = 1;
  if (a1 * a2 < 0.) {
    while (dc >= tolSt) {
      a0 = getSt(G, n, c0);
      if (a0 * a1 < 0) { a2 = a0; c2 = c0; } else { a1 = a0; c1 = c0; }
      dc = c2 - c1;
      c0 = (c2 + c1) / 2.;
    }
  } else { *ok = 0; return 0.; }
  a0 = getSt(G, n, c0);
  if (fabs(a0) < .5) return c0;
  return cst_in; /* intent(out) left undefined by the Fortran: the caller's value survives in practice */
}

/* scan ccc(from..index0) for a sign change of f, startl at every upper end: the common tail of FundaMode / StMode */
static int scan_ccc(grt_t* G, secf f, int from, int index0, double* root) {
  double imf, k1, k2, f1, f2;
  int iq = -1;
  k2 = G->ccc[from];
  startl(G, k2);
  f2 = f(G, k2, &imf);
  for (int ip = from + 1; ip <= index0; ++ip) {
    k1 = k2; f1 = f2;
    k2 = G->ccc[ip];
    startl(G, k2);
    f2 = f(G, k2, &imf);
    iq = -1;
    if (f1 * f2 < 0.) {
      double kt = bisecim(G, f, k1, k2, f1, f2, &iq);
      if (iq == 0) { *root = kt; return 0; }
    }
  }
  return iq == 0 ? 0 : 1;
}

/* FundaMode (Rayleigh, no water): SearchRayleigh.f90:416-606 */
static int fundamode_R(grt_t* G, int index0, int im1, double* cray) {
  double cmn, cmx, kk[101], imf;
  int nk, iq = -1, ierr = 1;
  if (*cray > 0.) { cmn = (double)0.90f * *cray; nk = 10; }
  else { cmn = cr0_finder(G->vs1, G->vp[1]); cmn = cmn - .1; nk = 100; }
  cmx = G->vs1 - 4 * G->tol;
  if (cmx > cmn) {
    for (int i = 1; i <= nk; ++i) kk[i] = cmn + (cmx - cmn) / (double)nk * i;
    startl(G, kk[nk]);
    int j = 1;
    iq = 1;
    double k1, k2 = kk[j], f1, f2 = secfun_surf(G, k2, &imf);
    for (;;) {
      k1 = k2; f1 = f2;
      j++;
      k2 = kk[j];
      f2 = secfun_surf(G, k2, &imf);
      if (f1 * f2 < 0.) {
        double kt = bisecim(G, secfun_surf, k1, k2, f1, f2, &iq);
        if (iq == 0) { ierr = 0; *cray = kt; break; }
      }
      if (j == nk) break;
    }
  }
  if (iq != 0 && G->nlvl1 > 0) {
    double r;
    if (scan_ccc(G, secfun_surf, im1, index0, &r) == 0) { *cray = r; ierr = 0; }
  } else if (iq != 0 && G->nlvl1 == 0) {
    int im2 = im1;
    while (G->ccc[im2] < *cray - 2 * G->dc) { im2++; if (im2 == index0) break; }
    double r;
    if (scan_ccc(G, secfun_surf, im2, index0, &r) == 0) { *cray = r; ierr = 0; }
  }
  return ierr;
}

/* StMode (Rayleigh under a water layer): SearchRayleigh.f90:283-414 */
static int stmode_R(grt_t* G, int index0, double* cSt) {
  const double dc = 1e-3;
  double imf, cmn = *cSt * .75, k1, k2, f1, f2;
  double a = G->vs[G->ifs + 1], b = G->vp[G->ifs];
  double cmx = (a < b ? a : b) - 4 * G->tol;
  int iq = -1, ierr = 1;
  startl(G, cmx);
  k2 = cmx;
  f2 = secfun_st(G, k2, &imf);
  for (;;) {
    k1 = k2; f1 = f2;
    k2 = k2 - dc;
    if (k2 < cmn) break;
    f2 = secfun_st(G, k2, &imf);
    if (f1 * f2 < 0.) {
      double kt = bisecim(G, secfun_st, k1, k2, f1, f2, &iq);
      if (iq == 0) { *cSt = kt; ierr = 0; break; }
    }
  }
  if (iq != 0) {
    int im2 = 1;
    while (G->ccc[im2] < cmx) { im2++; if (im2 == index0) break; }
    double r;
    /* the Fortran loop leaves iq at its last value: found <=> iq == 0 */
    if (scan_ccc(G, secfun_st, im2, index0, &r) == 0) { *cSt = r; ierr = 0; iq = 0; }
  }
  if (iq != 0) *cSt = 0;
  return ierr;
}

/* FundaMode of SearchLove (SearchLove.f90): `check` there tests the host's never-assigned `c` against the layer
 * velocities; with c undefined (0 in practice) it cannot match a velocity, so every root passes. */
static int fundamode_L(grt_t* G, int index0, double* cray) {
  double c0 = *cray, r, imf;
  int ierr = 1;
  if (scan_ccc(G, secfun_L, 1, index0, &r) == 0) { *cray = r; ierr = 0; }
  if (ierr == 1) {
    double k2 = (double)1.1f * c0, k1, f1, f2;
    startl(G, k2);
    f2 = secfun_L(G, k2, &imf);
    for (;;) {
      k1 = k2 - G->dc;
      if (!(k1 >= G->vsm)) break; /* if(k1<vsm) exit; a NaN k1 (c0 = 0) would loop for ever in the Fortran only if vsm were NaN */
      startl(G, k1);
      f1 = secfun_L(G, k1, &imf);
      int iq = -1;
      if (f1 * f2 < 0.) {
        double kt = bisecim(G, secfun_L, k1, k2, f1, f2, &iq);
        if (iq == 0) { *cray = kt; ierr = 0; break; }
      }
      k2 = k1; f2 = f1;
    }
  }
  if (ierr == 1) *cray = 0;
  return ierr;
}

/* SearchRayleigh / SearchLove with allmodes = 0: one root per call */
static int search_one(grt_t* G, double c0, double* root) {
  int ncc = 0, im1 = 0, ierr;
  memset(G->ccc, 0, sizeof(double) * GNV);
  const int love = G->modetype == 0;
  c_interval(G, love, &ncc, &im1);
  if (G->overflow) return 1;
  double cray = c0;
  if (love) ierr = fundamode_L(G, ncc, &cray);
  else if (G->ifs == 0) ierr = fundamode_R(G, ncc, im1, &cray);
  else {
    if (cray <= 0) { int ok; cray = st_finder(G, G->ifs, cray, &ok); if (!ok) return 1; }
    ierr = stmode_R(G, ncc, &cray);
  }
  *root = cray;
  return ierr;
}

/*
 * The GRT branch of surfmodes (surfmodes.f90:84-87,96-99 -> RayleighModes / LoveModes :185-306).
 * par = {tolmin, tolmax, smin_min, smin_max, dcm, dc2} (T_MODES_PARA, surfmodes.f90:23-30); dc = paras%dc.
 * phase/group entries the Fortran never assigns are left untouched.  Returns ierr (0/1); -1 fluid below the top;
 * counters[0] += secular-function evaluations, counters[1] += interface steps.
 */
int orc_grt_modes(const double* thick, const double* vp, const double* vs, const double* rho, int n, const double* freqs, int np,
                  int modetype, int phaseGroup, double dc, const double* par, int math_mode, double* phase, double* group,
                  int64_t* counters) {
  if (n > GNL || n < 2) return 3;
  grt_t* G = (grt_t*)calloc(1, sizeof(grt_t));
  G->ccc = (double*)calloc(GNV + 2, sizeof(double));
  G->vvv = (double*)calloc(GNV + 2, sizeof(double));
  G->math_mode = math_mode;
  int ierr = 0;
  if (setup_grt(G, thick, vp, vs, rho, n, modetype, dc, par[4], par[5]) < 0) ierr = -1;
  else if (modetype == 1 && G->ifs > 1) ierr = -1;
  else {
    const double pi_m = (double)3.1415926f; /* m_surfmodes pi */
    const double dh = (double)0.005f;
    const double tolmin = par[0], tolmax = par[1], smn = par[2], smx = par[3];
    double c0 = 0;
    for (int i = 1; i <= np; ++i) {
      G->w = freqs[i - 1] * 2 * pi_m;
      G->tol = tolmin + (np + 1 - i) * (tolmax - tolmin) / np;
      G->smin = smn + (i - 1) * (smx - smn) / np;
      G->index_a = i;
      double root = 0;
      int ierr1 = search_one(G, c0, &root);
      if (ierr1 == 1) { ierr = 1; break; }
      phase[i - 1] = root;
      c0 = phase[i - 1];
      if (phaseGroup == 1) {
        double freq0 = freqs[i - 1] + dh;
        G->w = freq0 * 2 * pi_m;
        double root0 = 0;
        ierr = search_one(G, c0, &root0);
        if (ierr == 1) break;
        double g = (freqs[i - 1] + dh) / root0 - freqs[i - 1] / phase[i - 1]; /* CalGroup :308-320 */
        group[i - 1] = g > 0 ? dh / g : 0;
      }
    }
  }
  if (counters) { counters[0] += G->n_secfun; counters[1] += G->n_layer; }
  free(G->ccc); free(G->vvv); free(G);
  return ierr;
}

/* test hook: what setup_grt + startl leave in the T_GRT of a column at phase velocity c -- the inputs of SecFuns_L / SecFunSurf,
 * for the comparison with the mechanical translations of Love.f90 and Rayleigh.f90 (oracle/f90toc_love.py).
 * d / vp / vs / mu: n values; ints = {ifs, ll, lvlast}. */
int orc_grt_state(const double* thick, const double* vp, const double* vs, const double* rho, int n, double freq, int modetype, double c,
                  double* d_out, double* vp_out, double* vs_out, double* mu_out, int* ints, double* w_out) {
  grt_t* G = (grt_t*)calloc(1, sizeof(grt_t));
  G->math_mode = 0;
  if (setup_grt(G, thick, vp, vs, rho, n, modetype, 1e-3, 1e-3, 1e-3) < 0) { free(G); return -1; }
  G->w = freq * 2 * (double)3.1415926f;
  startl(G, c);
  for (int j = 1; j <= n; ++j) { d_out[j - 1] = G->d[j]; vp_out[j - 1] = G->vp[j]; vs_out[j - 1] = G->vs[j]; mu_out[j - 1] = G->mu[j]; }
  ints[0] = G->ifs; ints[1] = G->ll; ints[2] = G->lvlast;
  *w_out = G->w;
  free(G);
  return 0;
}

/* test hook: one root refinement as the searches issue it -- startl at the upper end k2 of a bracket, the secular function at
 * both ends, bisecim in between (util.f90:90-167) -- for the comparison with the mechanical translation of bisecim driving the
 * translated secular functions.  out = {root, f1, f2}; returns iq, or -2 when setup fails. */
int orc_grt_bisecim(const double* thick, const double* vp, const double* vs, const double* rho, int n, double freq, int modetype,
                    double k1, double k2, double smin, double tol, double* out) {
  grt_t* G = (grt_t*)calloc(1, sizeof(grt_t));
  G->math_mode = 0;
  if (setup_grt(G, thick, vp, vs, rho, n, modetype, 1e-3, 1e-3, 1e-3) < 0) { free(G); return -2; }
  G->w = freq * 2 * (double)3.1415926f;
  G->smin = smin; G->tol = tol;
  secf f = modetype == 0 ? secfun_L : (G->ifs == 0 ? secfun_surf : secfun_st);
  double imf = 0;
  startl(G, k2);
  const double f1 = f(G, k1, &imf), f2 = f(G, k2, &imf);
  int iq = -1;
  out[0] = bisecim(G, f, k1, k2, f1, f2, &iq);
  out[1] = f1; out[2] = f2;
  free(G);
  return iq;
}

/* test hooks: the trial phase velocities of a frequency (C_Interval / C_Interval_L with N_cf / N_cf_L and sort) and the rest of
 * the T_GRT they read.  ccc_out: 20000 doubles; v_out: 2n doubles; extra = {vsy, vsm, vs1}; counts = {ncc, im1, nv, overflow}. */
int orc_grt_cinterval(const double* thick, const double* vp, const double* vs, const double* rho, int n, double freq, int modetype,
                      double tol, double* ccc_out, double* v_out, double* extra, int* counts) {
  grt_t* G = (grt_t*)calloc(1, sizeof(grt_t));
  G->math_mode = 0;
  if (setup_grt(G, thick, vp, vs, rho, n, modetype, 1e-3, 1e-3, 1e-3) < 0) { free(G); return -1; }
  G->w = freq * 2 * (double)3.1415926f;
  G->tol = tol;
  G->vvv = (double*)calloc(GNV + 8, sizeof(double));
  G->ccc = (double*)calloc(GNV + 8, sizeof(double));
  int ncc = 0, im1 = 0;
  c_interval(G, modetype == 0, &ncc, &im1);
  for (int i = 1; i <= ncc && i <= GNV; ++i) ccc_out[i - 1] = G->ccc[i];
  for (int i = 1; i <= G->nv; ++i) v_out[i - 1] = G->v[i];
  extra[0] = G->vsy; extra[1] = G->vsm; extra[2] = G->vs1;
  counts[0] = ncc; counts[1] = im1; counts[2] = G->nv; counts[3] = G->overflow;
  free(G->vvv); free(G->ccc); free(G);
  return 0;
}

/* test hook: everything setup_grt (surfmodes.f90:320-450) leaves in the T_GRT.  mu_out: n, v_out: 2n, lvls_out: n/2+1 values;
 * ints = {ifs, no_lvl, no_lvl_fl, nlvl1, nlvls1, lvlast, L1, nv}; dbl = {mu0, vsy, vs1, vsm, vss1}. */
int orc_grt_setup(const double* thick, const double* vp, const double* vs, const double* rho, int n, int modetype, double* mu_out,
                  double* v_out, int* lvls_out, int* ints, double* dbl) {
  grt_t* G = (grt_t*)calloc(1, sizeof(grt_t));
  int rc = setup_grt(G, thick, vp, vs, rho, n, modetype, 1e-3, 1e-3, 1e-3);
  for (int j = 1; j <= n; ++j) mu_out[j - 1] = G->mu[j];
  for (int j = 1; j <= 2 * n; ++j) v_out[j - 1] = j <= G->nv ? G->v[j] : 0.0;
  for (int j = 1; j <= n / 2 + 1; ++j) lvls_out[j - 1] = G->lvls[j];
  ints[0] = G->ifs; ints[1] = G->no_lvl; ints[2] = G->no_lvl_fl; ints[3] = G->nlvl1; ints[4] = G->nlvls1; ints[5] = G->lvlast;
  ints[6] = G->L1; ints[7] = G->nv;
  dbl[0] = G->mu0; dbl[1] = G->vsy; dbl[2] = G->vs1; dbl[3] = G->vsm; dbl[4] = G->vss1;
  free(G);
  return rc;
}

/* test hook: the secular function itself (modetype 1: SecFunSurf / SecFunSt, 0: SecFuns_L) at phase velocity c */
int orc_grt_secfun(const double* thick, const double* vp, const double* vs, const double* rho, int n, double freq, int modetype,
                   double c, int math_mode, double* re, double* im) {
  grt_t* G = (grt_t*)calloc(1, sizeof(grt_t));
  G->math_mode = math_mode;
  if (setup_grt(G, thick, vp, vs, rho, n, modetype, 1e-3, 1e-3, 1e-3) < 0) { free(G); return -1; }
  G->w = freq * 2 * (double)3.1415926f;
  startl(G, c);
  double imf = 0, r;
  if (modetype == 0) r = secfun_L(G, c, &imf);
  else if (G->ifs == 0) r = secfun_surf(G, c, &imf);
  else r = secfun_st(G, c, &imf);
  *re = r; *im = imf;
  free(G);
  return 0;
}
