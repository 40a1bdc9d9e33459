// k5_grt.cuh -- the generalized reflection/transmission branch of `surfmodes` on the device: columns with a low-velocity
// layer, which the reference does not hand to surfdisp96 (surfmodes/surfmodes.f90:84-87,96-99) but to RayleighModes /
// LoveModes (:185-306) -> SearchRayleigh / SearchLove (allmodes = 0: one root per frequency, FundaMode / StMode,
// SearchRayleigh.f90:283-606; SearchLove.f90 FundaMode), C_Interval[_L] + N_cf[_L] (C_interval.f90, C_interval_L.f90),
// startl / SecFunSurf / SecFunSt / Stoneley / EinvE / EinvE_f / propup / propdn_f (Rayleigh.f90), SecFuns_L / EinvE_L /
// propdn_L / propup_L (Love.f90), bisecim / det3 / sort (util.f90).  `surfmmodes` has no such branch (:153,165).
//
// ONE WARP PER COLUMN.  The reference's search is a long scan of a secular function over a list of trial phase
// velocities (hundreds to thousands of points per frequency) followed by a short bisection.  The scan points are
// independent: the 32 lanes evaluate 32 consecutive points at once, the sign changes are then consumed in the
// reference's order; bisecim's evaluations depend on each other and run on every lane alike.  The trial list itself
// (C_Interval: an arithmetic fill or N_cf-driven subdivision, two sorts, a de-duplication) is built cooperatively:
// fills and the dedup are lane-parallel, the sorts are warp bitonic sorts in the column's scratch, the short serial
// parts run on lane 0.  Layer data are float64 (GRT%d = thick ...: no real(.,4) narrowing on this branch), re-derived
// from the model column by lane 0 with convert_to_layer's rules.
//
// Arithmetic: the reference's operations in the reference's order -- complex products (ac-bd, ad+bc), range-reduced
// complex division, sqrt of dcmplx(real) as a real square root on one axis, exp(x+iy) = e^x (cos y, sin y) with the
// portable exp / sincos of mct_math.h -- so the results are bit-identical to oracle/grt_ref.c in its portable math
// mode.  One structural shortcut, exact: EinvE(j,c,0) = matmul(b44,a44) has the block form [[X,Y],[Y,X]] (rows 3:4 of
// b44 and columns 3:4 of a44 repeat rows 1:2 / columns 1:2 with the middle two entries negated, and negation commutes
// with rounding), so 16 of the 64 complex products are formed and summed with signs.
#pragma once

#define GRT_NV 20000                                  // vvv / ccc length (C_interval.f90:5,14)
#define GRT_NVPAD 32768                               // bitonic sort pads to a power of two
#define GRT_LAY 5                                     // doubles per layer record: d, vp, vs, rho, mu
#define GRT_SCRATCH (GRT_LAY * (MCT_MAX_LAYERS + 2) + 2 * (MCT_MAX_LAYERS + 2) + GRT_NVPAD + GRT_NV + 8) // doubles per column

struct GrtParams {
  // model columns (layered by lane 0) or pre-layered input (mct_surfmodes_batch)
  const double* vp; const double* vs; const double* rho;
  int32_t ny, nz, ix0, iy0, wx, wy;
  long long model_stride;
  double dz, waterDepth, scaling, layer_eps, water_thresh;
  const double* pl_thick; const double* pl_vp; const double* pl_vs; const double* pl_rho; const long long* pl_off; // pre-layered (else null)
  int32_t modetype, phaseGroup, np;
  double dc, tolmin, tolmax, smin_min, smin_max, dcm, dc2;
  double freqs[MCT_MAX_PERIODS];
  const int32_t* list; int32_t nlist; // columns to solve
  int32_t* next;                      // the next list entry to hand out (zeroed by the host): blocks are persistent, columns differ widely in cost
  const int32_t* skip; int32_t cols_per_model;
  double* scratch;
  double* pvel; double* gvel; int32_t* ierr;
  unsigned long long* counters; // [0],[1]: secular-function evaluations, interface steps as the reference would spend them
  int32_t* flags;               // per model: flags[2*b+1] = max condition code (raised when a column cannot be solved here)
};

struct cxd { double re, im; };
__device__ __forceinline__ cxd CXD(double a, double b) { cxd z; z.re = a; z.im = b; return z; }
__device__ __forceinline__ cxd c_add(cxd a, cxd b) { return CXD(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cxd c_sub(cxd a, cxd b) { return CXD(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cxd c_neg(cxd a) { return CXD(-a.re, -a.im); }
__device__ __forceinline__ cxd c_mul(cxd a, cxd b) { return CXD(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cxd c_rs(double r, cxd a) { return CXD(r * a.re, r * a.im); }
__device__ __forceinline__ cxd c_dr(cxd a, double r) { return CXD(a.re / r, a.im / r); }
__device__ __forceinline__ cxd c_div(cxd a, cxd b) {
  double ratio, div, tr, ti;
  if (fabs(b.re) < fabs(b.im)) {
    ratio = b.re / b.im; div = (b.re * ratio) + b.im;
    tr = (a.re * ratio) + a.im; ti = (a.im * ratio) - a.re;
  } else {
    ratio = b.im / b.re; div = (b.im * ratio) + b.re;
    tr = (a.im * ratio) + a.re; ti = a.im - (a.re * ratio);
  }
  return CXD(tr / div, ti / div);
}
__device__ __forceinline__ cxd c_exp(cxd z) {
  const double e = mct_exp(z.re);
  double s, c;
  if (z.im == 0.0) { s = z.im; c = 1.0; } else mct_sincos(z.im, &s, &c);
  return CXD(e * c, e * s);
}
__device__ __forceinline__ cxd g_csq(double c, double vel) { // GRT.f90:111-116
  const double t = c / vel;
  const double x = 1 - t * t;
  return x >= 0 ? CXD(sqrt(x), 0.0) : CXD(0.0, sqrt(-x));
}
struct m2 { cxd a, b, c, d; }; // [[a, b], [c, d]]
__device__ __forceinline__ cxd dot2(cxd p, cxd q, cxd r, cxd s) { return c_add(c_add(CXD(0, 0), c_mul(p, q)), c_mul(r, s)); } // matmul element, k ascending
__device__ __forceinline__ m2 mm2(const m2& x, const m2& y) {
  m2 r;
  r.a = dot2(x.a, y.a, x.b, y.c); r.b = dot2(x.a, y.b, x.b, y.d);
  r.c = dot2(x.c, y.a, x.d, y.c); r.d = dot2(x.c, y.b, x.d, y.d);
  return r;
}
__device__ __forceinline__ m2 add2(const m2& x, const m2& y) { m2 r; r.a = c_add(x.a, y.a); r.b = c_add(x.b, y.b); r.c = c_add(x.c, y.c); r.d = c_add(x.d, y.d); return r; }
__device__ __forceinline__ m2 inv2m(const m2& x) { // Rayleigh.f90:26-32
  const cxd det = c_sub(c_mul(x.a, x.d), c_mul(x.c, x.b));
  m2 r;
  r.a = c_div(x.d, det); r.c = c_div(c_neg(x.c), det); r.b = c_div(c_neg(x.b), det); r.d = c_div(x.a, det);
  return r;
}

// the column's state, in shared memory (one warp per block)
struct GrtCol {
  int n, modetype, ifs, nlvl1, lvlast, nv;
  double mu0, vsy, vs1, vsm, v1;
  double dc, w, tol, smin;
  const double* lay; // GRT_LAY doubles per layer, 1-based: lay[GRT_LAY*j + {0:d,1:vp,2:vs,3:rho,4:mu}]
  const double* v;   // sorted velocities, 1-based
  double* vvv;       // 1-based work array (GRT_NVPAD)
  double* ccc;       // 1-based trial velocities
  int ncc, im1, overflow;
};
#define GL_D(G, j) ((G).lay[GRT_LAY * (j) + 0])
#define GL_VP(G, j) ((G).lay[GRT_LAY * (j) + 1])
#define GL_VS(G, j) ((G).lay[GRT_LAY * (j) + 2])
#define GL_RHO(G, j) ((G).lay[GRT_LAY * (j) + 3])
#define GL_MU(G, j) ((G).lay[GRT_LAY * (j) + 4])

// startl: Rayleigh.f90:62-105 (returns ll)
__device__ __forceinline__ int g_startl(const GrtCol& G, double c) {
  int sl = G.n;
  const double vk = G.w / c;
  double su = 0;
  for (int j = G.lvlast; j <= G.n - 1; ++j) {
    const double vsj = GL_VS(G, j);
    if (c < vsj) {
      su = su + vk * g_csq(c, vsj).re * GL_D(G, j);
      if (su > 46.0) { sl = j; break; }
    } else su = 0;
  }
  return sl;
}

// EinvE(j,c,0) = [[X,Y],[Y,X]]; also returns cp/cs of the layers above (0) and below (1) the interface
__device__ __forceinline__ void g_einve0(const GrtCol& G, int j, double c, m2& X, m2& Y, cxd& cp0, cxd& cs0, cxd& cp1, cxd& cs1) {
  // a44 columns 1:2 (layer j+1, Rayleigh.f90:374-379)
  const double as1 = GL_VS(G, j + 1), am1 = GL_MU(G, j + 1);
  cp1 = g_csq(c, GL_VP(G, j + 1)); cs1 = g_csq(c, as1);
  double t = c / as1;
  const cxd xi1 = CXD(1 - t * t / 2., 0.0);
  const cxd IC = CXD(1.0, 0.0);
  const cxd A[4][2] = {{IC, cs1}, {cp1, IC}, {c_rs(-am1, cp1), c_rs(-am1, xi1)}, {c_rs(-am1, xi1), c_rs(-am1, cs1)}};
  // b44 rows 1:2 (layer j, :380-386), divided by 2*(1-xi)
  const double as0 = GL_VS(G, j), am0 = GL_MU(G, j);
  cp0 = g_csq(c, GL_VP(G, j)); cs0 = g_csq(c, as0);
  t = c / as0;
  const cxd xi0 = CXD(1 - t * t / 2., 0.0);
  const cxd den = c_rs(2.0, c_sub(IC, xi0));
  cxd B[2][4] = {{IC, c_neg(c_div(xi0, cp0)), c_neg(c_div(IC, c_rs(am0, cp0))), c_dr(IC, am0)},
                 {c_neg(c_div(xi0, cs0)), IC, c_dr(IC, am0), c_neg(c_div(IC, c_rs(am0, cs0)))}};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) B[r][k] = c_div(B[r][k], den);
  cxd xs[2][2], ys[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const cxd t0 = c_mul(B[r][0], A[0][q]), t1 = c_mul(B[r][1], A[1][q]), t2 = c_mul(B[r][2], A[2][q]), t3 = c_mul(B[r][3], A[3][q]);
      xs[r][q] = c_add(c_add(c_add(c_add(CXD(0, 0), t0), t1), t2), t3);
      ys[r][q] = c_add(c_sub(c_sub(c_add(CXD(0, 0), t0), t1), t2), t3); // rows/columns 2,3 negated: x + (-t) == x - t exactly
    }
  X.a = xs[0][0]; X.b = xs[0][1]; X.c = xs[1][0]; X.d = xs[1][1];
  Y.a = ys[0][0]; Y.b = ys[0][1]; Y.c = ys[1][0]; Y.d = ys[1][1];
}

// propup(c, j2, j1): Rdu(:,:,j1); la_out = the module variable `la` as the routine leaves it (Rayleigh.f90:449-491)
__device__ __forceinline__ m2 g_propup(const GrtCol& G, double c, int j2, int j1, cxd la_out[2], unsigned& nlay) {
  const double vk = G.w / c;
  m2 X, Y;
  cxd cp0, cs0, cp1, cs1;
  g_einve0(G, j2, c, X, Y, cp0, cs0, cp1, cs1);
  m2 a22 = inv2m(X);
  double f = -GL_D(G, j2) * vk;
  cxd la0 = c_exp(c_rs(f, cp0)), la1 = c_exp(c_rs(f, cs0));
  a22.a = c_mul(a22.a, la0); a22.c = c_mul(a22.c, la0); a22.b = c_mul(a22.b, la1); a22.d = c_mul(a22.d, la1);
  m2 Rdu = mm2(Y, a22);
  nlay += 1;
  for (int j = j2 - 1; j >= j1; --j) {
    g_einve0(G, j, c, X, Y, cp0, cs0, cp1, cs1);
    m2 b22 = Rdu;
    f = -GL_D(G, j + 1) * vk;
    la0 = c_exp(c_rs(f, cp1)); la1 = c_exp(c_rs(f, cs1));
    b22.a = c_mul(b22.a, la0); b22.b = c_mul(b22.b, la0); b22.c = c_mul(b22.c, la1); b22.d = c_mul(b22.d, la1);
    a22 = inv2m(add2(X, mm2(Y, b22)));
    f = -GL_D(G, j) * vk;
    la0 = c_exp(c_rs(f, cp0)); la1 = c_exp(c_rs(f, cs0));
    a22.a = c_mul(a22.a, la0); a22.c = c_mul(a22.c, la0); a22.b = c_mul(a22.b, la1); a22.d = c_mul(a22.d, la1);
    b22 = add2(Y, mm2(X, b22));
    Rdu = mm2(b22, a22);
    nlay += 1;
  }
  la_out[0] = la0; la_out[1] = la1;
  return Rdu;
}

// a44 = EinvE(j,c,1) rows 2:4 (0-based 1..3) of columns 1:2 (lo) and 3:4 (hi), layer j+1
__device__ __forceinline__ void g_einve1_rows(const GrtCol& G, int j, double c, cxd lo[3][2], cxd hi[3][2], cxd& cp1, cxd& cs1) {
  const double as1 = GL_VS(G, j + 1), am1 = GL_MU(G, j + 1);
  cp1 = g_csq(c, GL_VP(G, j + 1)); cs1 = g_csq(c, as1);
  const double t = c / as1;
  const cxd xi1 = CXD(1 - t * t / 2., 0.0);
  const cxd IC = CXD(1.0, 0.0);
  lo[0][0] = cp1; lo[0][1] = IC;
  lo[1][0] = c_rs(-am1, cp1); lo[1][1] = c_rs(-am1, xi1);
  lo[2][0] = c_rs(-am1, xi1); lo[2][1] = c_rs(-am1, cs1);
  hi[0][0] = c_neg(lo[0][0]); hi[0][1] = c_neg(lo[0][1]); // rows 2:3 of columns 3:4 negated
  hi[1][0] = c_neg(lo[1][0]); hi[1][1] = c_neg(lo[1][1]);
  hi[2][0] = lo[2][0]; hi[2][1] = lo[2][1];
}

// SecFunSurf(0,c): Rayleigh.f90:125-154
__device__ __noinline__ double g_secfun_surf(const GrtCol& G, double c, int ll, double* imf, unsigned& nlay) {
  cxd la[2];
  const m2 Rdu = g_propup(G, c, ll - 1, 1, la, nlay);
  cxd lo[3][2], hi[3][2], cp1, cs1;
  g_einve1_rows(G, 0, c, lo, hi, cp1, cs1);
  m2 b22;
  b22.a = c_mul(hi[1][0], la[0]); b22.c = c_mul(hi[2][0], la[0]); b22.b = c_mul(hi[1][1], la[1]); b22.d = c_mul(hi[2][1], la[1]);
  m2 e31; e31.a = lo[1][0]; e31.b = lo[1][1]; e31.c = lo[2][0]; e31.d = lo[2][1];
  const m2 a22 = add2(e31, mm2(b22, Rdu));
  const cxd dsp = c_sub(c_mul(a22.a, a22.d), c_mul(a22.b, a22.c));
  *imf = dsp.im;
  return dsp.re;
}

// SecFunSt(ifs,c) with ifs = 1: Stoneley (Rayleigh.f90:157-221), EinvE_f(0,c,1) (:238-266), det3 (util.f90)
__device__ __noinline__ double g_secfun_st(const GrtCol& G, double c, int ll, double* imf, unsigned& nlay) {
  const int ifs = G.ifs;
  const double vk = G.w / c;
  const cxd rud0 = c_exp(c_rs(-GL_D(G, 1) * vk, g_csq(c, GL_VP(G, 1))));
  cxd la[2];
  const m2 Rdu = g_propup(G, c, ll - 1, ifs + 1, la, nlay);
  cxd lo[3][2], hi[3][2], cp1, cs1;
  g_einve1_rows(G, ifs, c, lo, hi, cp1, cs1);
  const double f = -GL_D(G, ifs + 1) * vk;
  la[0] = c_exp(c_rs(f, cp1)); la[1] = c_exp(c_rs(f, cs1));
  cxd b33[3][3];
  const cxd Z = CXD(0, 0);
  b33[0][0] = c_mul(Rdu.a, la[0]); b33[0][1] = c_mul(Rdu.b, la[0]); b33[0][2] = Z;
  b33[1][0] = c_mul(Rdu.c, la[1]); b33[1][1] = c_mul(Rdu.d, la[1]); b33[1][2] = Z;
  b33[2][0] = Z; b33[2][1] = Z;
  // EinvE_f(ifs-1,c,1): the water layer
  const cxd cpw = g_csq(c, GL_VP(G, ifs));
  const cxd xi = CXD(GL_RHO(G, ifs) * (c * c) / (2. * G.mu0), 0.0);
  const cxd IC = CXD(1.0, 0.0);
  const cxd w11 = IC, w21 = c_div(xi, cpw), w12 = IC, w22 = c_neg(w21);
  b33[2][2] = c_mul(rud0, c_exp(c_rs(-GL_D(G, ifs) * vk, cpw)));
  cxd a33[3][3], r33[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { a33[i][0] = hi[i][0]; a33[i][1] = hi[i][1]; }
  a33[0][2] = c_neg(w11); a33[1][2] = c_neg(Z); a33[2][2] = c_neg(w21);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      cxd s = Z;
#pragma unroll
      for (int k = 0; k < 3; ++k) s = c_add(s, c_mul(a33[i][k], b33[k][j]));
      r33[i][j] = s;
    }
#pragma unroll
  for (int i = 0; i < 3; ++i) { a33[i][0] = lo[i][0]; a33[i][1] = lo[i][1]; }
  a33[0][2] = c_neg(w12); a33[1][2] = c_neg(Z); a33[2][2] = c_neg(w22);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) a33[i][j] = c_add(a33[i][j], r33[i][j]);
  const cxd q1 = c_sub(c_mul(a33[1][1], a33[2][2]), c_mul(a33[1][2], a33[2][1]));
  const cxd q2 = c_sub(c_mul(a33[1][0], a33[2][2]), c_mul(a33[1][2], a33[2][0]));
  const cxd q3 = c_sub(c_mul(a33[1][0], a33[2][1]), c_mul(a33[1][1], a33[2][0]));
  const cxd dsp = c_add(c_sub(c_mul(a33[0][0], q1), c_mul(a33[0][1], q2)), c_mul(a33[0][2], q3));
  *imf = dsp.im;
  return dsp.re;
}

// Love.f90: EinvE_L(j,c,0) = matmul(b22,a22)
__device__ __forceinline__ m2 g_einveL0(const GrtCol& G, int j, double c, cxd& cs0, cxd& cs1) {
  const cxd IC = CXD(1.0, 0.0);
  cs1 = g_csq(c, GL_VS(G, j + 1));
  m2 a;
  a.b = IC; a.d = c_rs(GL_MU(G, j + 1), cs1);
  a.a = IC; a.c = c_neg(a.d);
  cs0 = g_csq(c, GL_VS(G, j));
  m2 b;
  b.a = c_rs(GL_MU(G, j), cs0); b.c = b.a;
  b.b = c_neg(IC); b.d = IC;
  const cxd den = c_rs(2., b.a);
  b.a = c_div(b.a, den); b.b = c_div(b.b, den); b.c = c_div(b.c, den); b.d = c_div(b.d, den);
  return mm2(b, a);
}
// SecFuns_L(1+ifs, c): Love.f90:108-125
__device__ __noinline__ double g_secfun_L(const GrtCol& G, double c, int ll, double* imf, unsigned& nlay) {
  const int lay = 1 + G.ifs;
  const double vk = G.w / c;
  const cxd IC = CXD(1.0, 0.0);
  // propdn_L(c,1+ifs,lay): EinvE_L(ifs,c,1), RudL(ifs)
  const cxd cs1f = g_csq(c, GL_VS(G, G.ifs + 1));
  const cxd a22_22 = c_rs(GL_MU(G, G.ifs + 1), cs1f);
  const cxd a22_21 = c_neg(a22_22);
  const cxd la2 = c_exp(c_rs(-GL_D(G, 1 + G.ifs) * vk, cs1f));
  const cxd rud = c_div(c_mul(c_neg(a22_22), la2), a22_21);
  // propup_L(c,ll-1,lay)
  const int j2 = ll - 1;
  cxd cs0, cs1;
  m2 e = g_einveL0(G, j2, c, cs0, cs1);
  cxd td = c_div(c_exp(c_rs(-GL_D(G, j2) * vk, cs0)), e.a);
  cxd rdu = c_mul(e.c, td);
  nlay += 1;
  for (int j = j2 - 1; j >= lay; --j) {
    e = g_einveL0(G, j, c, cs0, cs1);
    const cxd r = c_mul(c_exp(c_rs(-GL_D(G, j + 1) * vk, cs1)), rdu);
    td = c_div(c_exp(c_rs(-GL_D(G, j) * vk, cs0)), c_add(e.a, c_mul(e.b, r)));
    rdu = c_mul(c_add(e.c, c_mul(e.d, r)), td);
    nlay += 1;
  }
  const cxd dsp = c_sub(IC, c_mul(rud, rdu));
  *imf = dsp.re;
  return dsp.im;
}

// which secular function: 0 SecFunSurf, 1 SecFunSt, 2 SecFuns_L
__device__ __forceinline__ double g_secf(const GrtCol& G, int kind, double c, int ll, double* imf, unsigned& nlay) {
  if (kind == 0) return g_secfun_surf(G, c, ll, imf, nlay);
  if (kind == 1) return g_secfun_st(G, c, ll, imf, nlay);
  return g_secfun_L(G, c, ll, imf, nlay);
}

// bisecim: util.f90:68-122.  The bracket follows a pure bisection path (the interpolated estimate xt2 only enters the
// stopping test and the result), so the warp speculates five levels at a time: lane t-1 evaluates node t of the binary
// tree of mid-points below the current bracket (heap numbering: left child = the half the Fortran keeps when
// fa*y(3) < 0), then the path is walked with the reference's own arithmetic and only the evaluations it consumes are
// counted.  Mid-points are formed from the same operands in the same order as the sequential code forms them.
__device__ double g_bisecim(const GrtCol& G, int kind, int ll, double x1, double x2, double f1, double f2, int* iq, unsigned& nsec, unsigned& nlay,
                            int lane) {
  double xa = x1, xb = x2, ya = f1, yb = f2, xt1, xt2, dx, dxt, u1, u2;
  const double fa = f1;
  const double smin2 = G.smin * G.smin * 2.;
  int nc = 0;
  dx = fabs(x1 - x2);
  xt1 = (xa + xb) / 2.;
  const int t = lane + 1;                       // node 1..31 (lane 31 repeats node 16's work unused)
  const int L = 31 - __clz(t > 31 ? 16 : t);    // level of the node, 0..4
  const int tt = t > 31 ? 16 : t;
  for (;;) {
    // this lane's node: descend from the current bracket
    double a = xa, b = xb;
    for (int d = 0; d < L; ++d) {
      const double m = (a + b) / 2.;
      if ((tt >> (L - 1 - d)) & 1) a = m; else b = m;
    }
    const double xn = (a + b) / 2.;
    double imfn = 0;
    unsigned layn = 0;
    const double yn = g_secf(G, kind, xn, ll, &imfn, layn);
    int cur = 1;
    for (int lev = 0; lev < 5; ++lev) {
      const double x3 = __shfl_sync(0xffffffffu, xn, cur - 1), y3 = __shfl_sync(0xffffffffu, yn, cur - 1);
      const double imf = __shfl_sync(0xffffffffu, imfn, cur - 1);
      nsec += 1; nlay += __shfl_sync(0xffffffffu, layn, cur - 1);
      u1 = (xb - xa) / (yb - ya);
      u2 = (xb - x3) / (yb - y3);
      xt2 = xa - ya * (u1 - yb * ((u2 - u1) / (y3 - ya)));
      dxt = fabs(xt2 - xt1);
      dx = dx / 2.;
      if (dx < dxt) dxt = dx;
      if (dxt < G.tol) {
        double imf2 = 0;
        u1 = g_secf(G, kind, xt2, ll, &imf2, nlay); nsec++;
        u2 = y3;
        if (u1 * u1 + imf2 * imf2 < smin2 || u2 * u2 + imf * imf < smin2) { *iq = 0; return fabs(u1) < fabs(u2) ? xt2 : x3; }
        *iq = -1;
        return 0.;
      }
      xt1 = xt2;
      if (fa * y3 < 0) { xb = x3; yb = y3; cur = 2 * cur; } else { xa = x3; ya = y3; cur = 2 * cur + 1; }
      nc++;
      if (nc >= 1000) { *iq = -1; return 0; }
    }
  }
}

// N_cf / N_cf_L (C_interval.f90:178-199, C_interval_L.f90): real part of the sum only
__device__ double g_ncf(const GrtCol& G, double c, int love) {
  const double pi_s = love ? 3.1415926535897932 : (double)3.1415926f;
  double sum = 0.;
  for (int i = 1; i <= G.n - 1; ++i) {
    double yp = 0., ys = 0.;
    if (!love) { const double t = c / GL_VP(G, i); const double x = t * t - 1; yp = x >= 0 ? sqrt(x) : 0.; }
    const double vsi = GL_VS(G, i);
    if (vsi > 0.0) { const double t = c / vsi; const double x = t * t - 1; ys = x >= 0 ? sqrt(x) : 0.; }
    sum = sum + 2.0 * (love ? ys : (yp + ys)) * GL_D(G, i) / c;
  }
  return G.w / (2.0 * pi_s) * sum;
}

// ascending sort of a[1..n] by the warp: a bitonic network, in shared memory when n <= GRT_SSORT, else in the column's
// scratch (pads to a power of two with +inf; n <= GRT_NVPAD).  Any ascending sort yields the array util.f90's sort does.
#define GRT_SSORT 2048
__device__ void g_bitonic(double* a, int m, int lane) {
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const double x = a[i], y = a[p];
          const bool up = (i & k) == 0;
          if (up ? (x > y) : (x < y)) { a[i] = y; a[p] = x; }
        }
      }
      __syncwarp();
    }
}
__device__ void g_sort(double* a1, int n, int lane, double* ssort) {
  double* a = a1 + 1;
  int m = 1;
  while (m < n) m <<= 1;
  if (m <= GRT_SSORT) {
    for (int i = lane; i < m; i += 32) ssort[i] = i < n ? a[i] : 1.0e300;
    __syncwarp();
    g_bitonic(ssort, m, lane);
    for (int i = lane; i < n; i += 32) a[i] = ssort[i];
    __syncwarp();
    return;
  }
  for (int i = n + lane; i < m; i += 32) a[i] = 1.0e300;
  __syncwarp();
  g_bitonic(a, m, lane);
}
// a[1..n0] ascending already, a[n0+1..n0+t] arbitrary (t <= GRT_SSORT): sort the tail in shared memory, then every element
// computes its place in the merged order by a binary search in the other run; tmp[1..n0+t] is scratch
__device__ void g_sort_tail(double* a1, int n0, int t, int lane, double* ssort, double* tmp1) {
  int m = 1;
  while (m < t) m <<= 1;
  for (int i = lane; i < m; i += 32) ssort[i] = i < t ? a1[n0 + 1 + i] : 1.0e300;
  __syncwarp();
  g_bitonic(ssort, m, lane);
  for (int i = 1 + lane; i <= n0; i += 32) { // prefix element: + the tail elements strictly below it
    const double x = a1[i];
    int lo = 0, hi = t;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ssort[mid] < x) lo = mid + 1; else hi = mid; }
    tmp1[i + lo] = x;
  }
  for (int k = lane; k < t; k += 32) { // tail element: + the prefix elements not above it
    const double x = ssort[k];
    int lo = 0, hi = n0;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a1[1 + mid] <= x) lo = mid + 1; else hi = mid; }
    tmp1[k + 1 + lo] = x;
  }
  __syncwarp();
  for (int i = 1 + lane; i <= n0 + t; i += 32) a1[i] = tmp1[i];
  __syncwarp();
}

// C_Interval / C_Interval_L: fills G.ccc, G.ncc, G.im1 (lane 0 owns the scalars)
__device__ void g_cinterval(GrtCol& G, int love, int lane, double* ssort) {
  double* vvv = G.vvv;
  double* ccc = G.ccc;
  const double pi_c = love ? 3.1415926535897932 : (double)3.1415926f;
  const double eps = 1e-10;
  const double freq = G.w / (2.0 * pi_c);
  const double lowv = love ? G.vsm : G.v1;
  int index0 = 0, ovf = 0;
  int NN = 0;
  int sorted_prefix = 0; // vvv(1..sorted_prefix) is ascending by construction (the arithmetic fill of the first branch)
  if (lane == 0) { // nint: halves away from zero
    const double dn = g_ncf(G, G.vsy, love) - g_ncf(G, G.vsm, love);
    NN = dn >= 0 ? (int)floor(dn + 0.5) : -(int)floor(-dn + 0.5);
  }
  NN = __shfl_sync(0xffffffffu, NN, 0);
  if (freq < (double)0.12f || NN < 2) {
    int n0 = (NN + 1) * (love ? 512 : 1024);
    if (n0 > GRT_NV - 200) { ovf = 1; n0 = GRT_NV - 200; }
    const double off = love ? 0.01 : 0.1;
    for (int i = 1 + lane; i <= n0; i += 32) vvv[i] = lowv - off + (double)i * (G.vsy - lowv + off) / (double)n0;
    index0 = n0 < 0 ? 0 : n0;
    sorted_prefix = index0;
    for (int i = 1 + lane; i <= 100; i += 32) vvv[index0 + i] = G.vsy * (1. - .008 * i);
    index0 += 100;
    __syncwarp();
  } else {
    if (lane == 0) {
#define GPUSH(val) do { if (index0 >= GRT_NV - 1) ovf = 1; else { index0++; vvv[index0] = (val); } } while (0)
      double c1 = lowv, c2 = G.vsy, c0 = c2, dc = 0, dN;
      GPUSH(c1);
      int NNc = NN;
      long guard = 0;
      while (c0 > c1 && !ovf) {
        if (NNc > 0) dc = (c2 - c1) / (NNc);
        c0 = c2 - dc;
        for (;;) {
          dN = g_ncf(G, c2, love) - g_ncf(G, c0, love);
          if (dN < .5) { GPUSH(c0); NNc = NNc - 1; c2 = c0; break; }
          c0 = (c2 + c0) / 2.0;
          if (++guard > 2000000) { ovf = 1; break; }
        }
      }
      int ij = 1;
      while (ij <= G.nv && G.v[ij] <= G.vsy) ij++;
      for (int i = 1; i <= ij - 2; ++i) {
        const double c01 = G.v[i];
        int ii = 1;
        for (int j = 1; j <= G.n; ++j) if (fabs(GL_VS(G, j) - c01) < eps || fabs(GL_VP(G, j) - c01) < eps) ii = j;
        const double hi = GL_D(G, ii);
        const double Ni = 2.0 * freq * hi / c01 + eps;
        const int nj = (int)floor(Ni);
        for (int j = 1; j <= nj; ++j) {
          const double q = (double)j / Ni;
          const double c00 = c01 / sqrt(1.0 - q * q);
          if (c00 <= G.vsy) GPUSH(c00);
        }
      }
    }
    index0 = __shfl_sync(0xffffffffu, index0, 0);
    ovf = __shfl_sync(0xffffffffu, ovf, 0);
    __syncwarp();
    g_sort(vvv, index0, lane, ssort);
    for (int j = 1; j <= 2; ++j) { // the mid-point of every sequential pair, twice (the second round runs over the unsorted result of the first)
      const int intemp = index0;
      int add = intemp - 1;
      if (add < 0) add = 0;
      if (index0 + add > GRT_NV - 1) { ovf = 1; add = GRT_NV - 1 - index0; if (add < 0) add = 0; }
      for (int i = 1 + lane; i <= add; i += 32) vvv[intemp + i] = (vvv[i] + vvv[i + 1]) / 2.0;
      index0 += add;
      __syncwarp();
    }
    if (index0 + 100 > GRT_NV - 1) ovf = 1;
    else {
      for (int i = 1 + lane; i <= 100; i += 32) vvv[index0 + i] = G.vsy - (double)i / 100. * 0.1;
      index0 += 100;
    }
    __syncwarp();
  }
  if (index0 + 13 > GRT_NV - 1) ovf = 1;
  else if (lane == 0) {
    for (int i = 1; i <= 10; ++i) GPUSH(G.vsy - (i * 10) * G.tol);
    GPUSH(G.vsy - 3.0 * G.tol);
    GPUSH(G.vs1 + 4 * G.tol);
    GPUSH(G.vs1 - 4 * G.tol);
#undef GPUSH
  }
  index0 = __shfl_sync(0xffffffffu, index0, 0);
  ovf = __shfl_sync(0xffffffffu, ovf, 0);
  __syncwarp();
  if (sorted_prefix > GRT_SSORT / 2 && index0 - sorted_prefix <= GRT_SSORT) g_sort_tail(vvv, sorted_prefix, index0 - sorted_prefix, lane, ssort, ccc);
  else g_sort(vvv, index0, lane, ssort);
  // "ensure no points are less than vsm": the sequence becomes s(1) = vsm + tol, s(m) = vvv(i0 + m - 2), m = 2 .. ; sorted input, so
  // i0 = 1 + #{vvv <= vsm}; when nothing exceeds vsm the array stays as it is
  int le = 0;
  for (int i = 1 + lane; i <= index0; i += 32) le += (vvv[i] <= G.vsm) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) le += __shfl_xor_sync(0xffffffffu, le, o);
  const bool shifted = le < index0;
  const int i0 = le + 1;
  const int slen = shifted ? index0 - (i0 - 1) + 1 : index0;
  const double s1 = G.vsm + G.tol;
#define GSEQ(m) (shifted ? ((m) == 1 ? s1 : vvv[i0 + (m) - 2]) : vvv[(m)])
  // "ensure no points are greater than vsy": ii = last m with s(m) < vsy (scanning down from the end)
  int ge = 0;
  for (int m = 1 + lane; m <= slen; m += 32) ge += (GSEQ(m) >= G.vsy) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ge += __shfl_xor_sync(0xffffffffu, ge, o);
  // the Fortran walks down while vvv(ii) >= vsy: with a sorted tail that removes exactly the trailing run of such entries;
  // s(1) = vsm + tol may break the order only at the front, where it cannot be >= vsy
  const int ii = slen - ge;
  // drop points closer than 10*tol to their predecessor (in s, not in the kept list)
  const double tol0 = 10 * G.tol;
  int kept = 0; // warp-uniform running count
  for (int m0 = 1; m0 <= ii; m0 += 32) {
    const int m = m0 + lane;
    bool keep = false;
    double val = 0;
    if (m <= ii) {
      val = GSEQ(m);
      keep = (m == 1) || !(val - GSEQ(m - 1) < tol0);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) ccc[kept + __popc(bal & ((1u << lane) - 1u)) + 1] = val;
    kept += __popc(bal);
  }
#undef GSEQ
  __syncwarp();
  const int ncc = ii >= 1 ? kept : 1;
  if (ii < 1 && lane == 0) ccc[1] = shifted ? s1 : vvv[1];
  __syncwarp();
  // im1
  int first = 0x7fffffff;
  for (int i = 1 + lane; i <= ncc; i += 32) {
    const bool hit = love ? (ccc[i] > G.vs1) : (ccc[i] >= G.vs1);
    if (hit) { first = i; break; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  if (lane == 0) {
    G.ncc = ncc;
    G.im1 = first == 0x7fffffff ? (love ? 0 : ncc) : first;
    G.overflow = ovf;
  }
  __syncwarp();
}

// Scan pts[1..npts] for a sign change of the secular function and refine it, in the reference's order.
//   per_point_ll: startl at every point (else the fixed `ll0`); swap: the Love fall-back passes (current, previous) to bisecim.
// Returns true and the root when one is accepted.  Evaluations are counted as the reference would spend them.
__device__ bool g_scan(const GrtCol& G, int kind, const double* pts, int npts, bool per_point_ll, int ll0, bool swap, int lane, double* root,
                       unsigned& nsec, unsigned& nlay) {
  if (npts < 1) return false;
  double fprev = 0, kprev = 0;
  for (int b = 1; b <= npts; b += 32) {
    const int ip = b + lane;
    double k = 0, f = 0, imf;
    int ll = ll0;
    unsigned mylay = 0;
    if (ip <= npts) {
      k = pts[ip];
      if (per_point_ll) ll = g_startl(G, k);
      f = g_secf(G, kind, k, ll, &imf, mylay);
    }
    double fl = __shfl_up_sync(0xffffffffu, f, 1), kl = __shfl_up_sync(0xffffffffu, k, 1);
    if (lane == 0) { fl = fprev; kl = kprev; }
    const bool has_prev = ip >= 2 && ip <= npts;
    unsigned sc = __ballot_sync(0xffffffffu, has_prev && (fl * f < 0.));
    const unsigned act = __ballot_sync(0xffffffffu, ip <= npts);
    int consumed = 0; // points of this batch the reference has evaluated so far
    while (sc) {
      const int l = __ffs(sc) - 1;
      sc &= sc - 1;
      // account for the scan evaluations up to and including lane l
      for (; consumed <= l; ++consumed) { nsec += 1; nlay += __shfl_sync(0xffffffffu, mylay, consumed); }
      const double k1 = __shfl_sync(0xffffffffu, kl, l), k2 = __shfl_sync(0xffffffffu, k, l);
      const double f1 = __shfl_sync(0xffffffffu, fl, l), f2 = __shfl_sync(0xffffffffu, f, l);
      const int llb = __shfl_sync(0xffffffffu, ll, l);
      int iq = -1;
      const double kt = swap ? g_bisecim(G, kind, llb, k2, k1, f2, f1, &iq, nsec, nlay, lane) : g_bisecim(G, kind, llb, k1, k2, f1, f2, &iq, nsec, nlay, lane);
      if (iq == 0) { *root = kt; return true; }
    }
    const int nact = __popc(act);
    for (; consumed < nact; ++consumed) { nsec += 1; nlay += __shfl_sync(0xffffffffu, mylay, consumed); }
    fprev = __shfl_sync(0xffffffffu, f, 31);
    kprev = __shfl_sync(0xffffffffu, k, 31);
  }
  return false;
}

// CR0_Finder: SearchRayleigh.f90:895-933
__device__ double g_cr0(double v1, double v2) {
  const double tol = 1e-7;
  double c = 0.8 * v1, R, DR;
  for (int it = 0; it < 100000; ++it) {
    const double ps = 1.0 / v1, pp = 1.0 / v2, p = 1.0 / c;
    const double p2 = p * p, ps2 = ps * ps, pp2 = pp * pp;
    const double sps = sqrt(p2 - ps2), spp = sqrt(p2 - pp2);
    const double t = ps2 - 2.0 * p2;
    R = t * t - 4.0 * p2 * sps * spp;
    DR = p2 * (8.0 * p * (ps2 - 2.0 * p2) + 8.0 * p * sps * spp + 4.0 * (p2 * p) * (spp / sps + sps / spp));
    c = c - R / DR;
    if (v1 - c < tol || v2 - c < tol || c != c) break;
    if (fabs(R / DR) < tol) break;
  }
  return c;
}
// St_Finder: SearchRayleigh.f90:846-893
__device__ double g_getSt(const GrtCol& G, int n, double x) {
  double t = x / GL_VP(G, n);
  const double a = 1 - t * t;
  t = x / GL_VP(G, n + 1);
  const double b = 1 - t * t;
  t = x / GL_VS(G, n + 1);
  const double c1 = t * t * (GL_RHO(G, n) / GL_RHO(G, n + 1));
  t = GL_VS(G, n + 1) / x;
  const double c2 = t * t;
  const double c = 1 - 1. / c2;
  const double u = 1 + c;
  return c1 * sqrt(b / a) + c2 * (u * u - 4 * sqrt(b * c));
}
__device__ double g_stfinder(const GrtCol& G, int n, double cst_in, int* ok) {
  const double tolSt = 1e-7;
  double c2 = GL_VP(G, n) < GL_VS(G, n + 1) ? GL_VP(G, n) : GL_VS(G, n + 1);
  double c1 = c2 * .8;
  c2 = c2 - (c2 - c1) / 1e4;
  double a2 = g_getSt(G, n, c2), a1 = g_getSt(G, n, c1), dc = c2 - c1, c0 = (c2 + c1) / 2., a0;
  *ok = 1;
  if (a1 * a2 < 0.) {
    while (dc >= tolSt) {
      a0 = g_getSt(G, n, c0);
      if (a0 * a1 < 0) { a2 = a0; c2 = c0; } else { a1 = a0; c1 = c0; }
      dc = c2 - c1;
      c0 = (c2 + c1) / 2.;
    }
  } else { *ok = 0; return 0.; }
  a0 = g_getSt(G, n, c0);
  if (fabs(a0) < .5) return c0;
  return cst_in;
}

// SearchRayleigh / SearchLove, allmodes = 0: one root.  Returns ierr (0/1).
__device__ int g_search_one(GrtCol& G, double c0, double* root, int lane, unsigned& nsec, unsigned& nlay, double* ssort) {
  const int love = G.modetype == 0;
  // (the Fortran zeroes ccc(20000) here; only ccc(1..ncc), which C_Interval fills, is ever read)
  g_cinterval(G, love, lane, ssort);
  if (G.overflow) return 1;
  const int index0 = G.ncc, im1 = G.im1;
  double* pts = G.vvv; // free again: the generated point lists of the fixed-step scans live here
  double cray = c0;
  int ierr = 1;
  if (love) { // FundaMode of SearchLove
    double r;
    if (g_scan(G, 2, G.ccc, index0, true, 0, false, lane, &r, nsec, nlay)) { cray = r; ierr = 0; }
    if (ierr == 1) {
      // k2 = 1.1*c0, then k1 = k2 - dc while k1 >= vsm
      int np = 0;
      if (lane == 0) {
        double k2 = (double)1.1f * c0;
        pts[++np] = k2;
        for (;;) {
          const double k1 = k2 - G.dc;
          if (!(k1 >= G.vsm) || np >= GRT_NVPAD - 2) break;
          pts[++np] = k1;
          k2 = k1;
        }
      }
      np = __shfl_sync(0xffffffffu, np, 0);
      __syncwarp();
      if (g_scan(G, 2, pts, np, true, 0, true, lane, &r, nsec, nlay)) { cray = r; ierr = 0; }
    }
    if (ierr == 1) cray = 0;
  } else if (G.ifs == 0) { // FundaMode of SearchRayleigh
    double cmn;
    int nk;
    if (cray > 0.) { cmn = (double)0.90f * cray; nk = 10; }
    else { cmn = g_cr0(G.vs1, GL_VP(G, 1)); cmn = cmn - .1; nk = 100; }
    const double cmx = G.vs1 - 4 * G.tol;
    bool found = false;
    if (cmx > cmn) {
      for (int i = 1 + lane; i <= nk; i += 32) pts[i] = cmn + (cmx - cmn) / (double)nk * i;
      __syncwarp();
      const int ll = g_startl(G, pts[nk]);
      double r;
      if (g_scan(G, 0, pts, nk, false, ll, false, lane, &r, nsec, nlay)) { cray = r; ierr = 0; found = true; }
    }
    if (!found) {
      int from = im1;
      if (G.nlvl1 == 0) { // search larger than vs1, from the first point within 2*dc of cray
        while (G.ccc[from] < cray - 2 * G.dc) { from++; if (from == index0) break; }
      }
      double r;
      if (g_scan(G, 0, G.ccc + (from - 1), index0 - from + 1, true, 0, false, lane, &r, nsec, nlay)) { cray = r; ierr = 0; }
    }
  } else { // Stoneley mode under one water layer (StMode)
    if (cray <= 0) { int ok; cray = g_stfinder(G, G.ifs, cray, &ok); if (!ok) return 1; }
    const double cmn = cray * .75;
    const double a = GL_VS(G, G.ifs + 1), b = GL_VP(G, G.ifs);
    const double cmx = (a < b ? a : b) - 4 * G.tol;
    const int ll = g_startl(G, cmx);
    int np = 0;
    if (lane == 0) {
      double k2 = cmx;
      pts[++np] = k2;
      for (;;) {
        k2 = k2 - 1e-3;
        if (k2 < cmn || np >= GRT_NVPAD - 2) break;
        pts[++np] = k2;
      }
    }
    np = __shfl_sync(0xffffffffu, np, 0);
    __syncwarp();
    double r;
    bool found = false;
    if (g_scan(G, 1, pts, np, false, ll, false, lane, &r, nsec, nlay)) { cray = r; ierr = 0; found = true; }
    if (!found) {
      int im2 = 1;
      while (G.ccc[im2] < cmx) { im2++; if (im2 == index0) break; }
      if (g_scan(G, 1, G.ccc + (im2 - 1), index0 - im2 + 1, true, 0, false, lane, &r, nsec, nlay)) { cray = r; ierr = 0; found = true; }
    }
    if (!found) cray = 0;
  }
  *root = cray;
  return ierr;
}

// convert_to_layer in float64 (likelihood_surf.F90:523-629 / forward_modelling.f90:72-175), or the pre-layered input;
// then setup_grt (surfmodes.f90:320-450).  Lane 0 only.  Returns n (0 = cannot be solved here).
__device__ int g_setup(const GrtParams& P, int col, GrtCol& G, double* lay, double* v) {
  int n = 0;
  if (P.pl_off) {
    const long long o = P.pl_off[col];
    n = (int)(P.pl_off[col + 1] - o);
    if (n > MCT_MAX_LAYERS || n < 2) return 0;
    for (int i = 1; i <= n; ++i) {
      lay[GRT_LAY * i + 0] = P.pl_thick[o + i - 1]; lay[GRT_LAY * i + 1] = P.pl_vp[o + i - 1];
      lay[GRT_LAY * i + 2] = P.pl_vs[o + i - 1]; lay[GRT_LAY * i + 3] = P.pl_rho[o + i - 1];
    }
  } else {
    const int cpm = P.wx * P.wy;
    const int b = col / cpm, cm = col - b * cpm;
    const int i = P.ix0 + cm / P.wy, j = P.iy0 + cm % P.wy;
    const size_t base = (size_t)b * (size_t)P.model_stride + ((size_t)(i - 1) * P.ny + (size_t)(j - 1)) * P.nz;
    const double* vp = P.vp + base;
    const double* vs = P.vs + base;
    const double* rho = P.rho + base;
    auto emit = [&](double th, double a, double bb, double r) {
      n = n + 1;
      if (n > MCT_MAX_LAYERS) return;
      double thk = th / P.scaling;
      if (n == 1 && P.waterDepth > 0) thk = P.waterDepth;
      lay[GRT_LAY * n + 0] = thk; lay[GRT_LAY * n + 1] = a; lay[GRT_LAY * n + 2] = bb; lay[GRT_LAY * n + 3] = r;
    };
    if (P.waterDepth > P.water_thresh) emit(P.waterDepth, 1.5, 0.0, 1.0);
    double last_vp = vp[0], last_vs = vs[0], last_rho = rho[0];
    int last_k = 1;
    for (int k = 2; k <= P.nz; ++k) {
      const double vv = vs[k - 1];
      if (fabs(vv - last_vs) > P.layer_eps) {
        emit((double)(k - last_k) * P.dz, last_vp, last_vs, last_rho);
        last_vp = vp[k - 1]; last_vs = vv; last_rho = rho[k - 1]; last_k = k;
      }
    }
    emit(0.0, vp[P.nz - 1], vs[P.nz - 1], rho[P.nz - 1]);
    if (n > MCT_MAX_LAYERS || n < 2) return 0;
  }
  const double eps = (double)1e-6f;
  G.n = n; G.modetype = P.modetype;
  G.lay = lay; G.v = v;
  int ifs = 0, idx = 0;
  for (int i = 1; i <= n; ++i) {
    const double vsi = lay[GRT_LAY * i + 2];
    if (fabs(vsi) > eps) { idx++; v[idx] = vsi; }
    else { if (i > 1) return 0; ifs++; }
    idx++; v[idx] = lay[GRT_LAY * i + 1];
  }
  for (int i = 2; i <= idx; ++i) { // insertion sort (<= 400 values)
    const double x = v[i];
    int j = i - 1;
    while (j >= 1 && v[j] > x) { v[j + 1] = v[j]; j--; }
    v[j + 1] = x;
  }
  G.nv = idx; G.ifs = ifs;
  if (P.modetype == 1 && ifs > 1) return 0;
  int cnt = 0;
  double mu0 = 0;
  for (int i = 1; i <= n; ++i) {
    const double vsi = lay[GRT_LAY * i + 2];
    const double mu = lay[GRT_LAY * i + 3] * (vsi * vsi);
    if (fabs(mu) > eps) { cnt++; mu0 = mu0 + mu; }
    lay[GRT_LAY * i + 4] = mu;
  }
  mu0 = mu0 / cnt;
  for (int i = 1; i <= n; ++i) lay[GRT_LAY * i + 4] = lay[GRT_LAY * i + 4] / mu0;
  G.mu0 = mu0;
  double vsy = lay[GRT_LAY * 1 + 2];
  for (int i = 2; i <= n; ++i) vsy = lay[GRT_LAY * i + 2] > vsy ? lay[GRT_LAY * i + 2] : vsy;
  G.vsy = vsy;
  G.v1 = v[1];
  double vs1;
  if (P.modetype == 1) {
    vs1 = ifs > 0 ? lay[GRT_LAY * 1 + 1] : lay[GRT_LAY * 1 + 2];
    G.vsm = v[1];
  } else {
    vs1 = lay[GRT_LAY * (1 + ifs) + 2];
    double m = lay[GRT_LAY * (ifs + 1) + 2];
    for (int i = ifs + 2; i <= n; ++i) m = lay[GRT_LAY * i + 2] < m ? lay[GRT_LAY * i + 2] : m;
    G.vsm = m;
  }
  G.vs1 = vs1;
  int no_lvl = 0, nlvl1 = 0, last_lvl = 0;
  for (int i = 2; i <= n - 1; ++i) {
    const double vpi = lay[GRT_LAY * i + 1], vsi = lay[GRT_LAY * i + 2];
    if (vpi < lay[GRT_LAY * (i + 1) + 1] && vpi < lay[GRT_LAY * (i - 1) + 1]) {
      no_lvl++; last_lvl = i;
      if (ifs == 0) { if (vsi < vs1) nlvl1++; }
      else if (P.modetype == 1) { if (vpi < vs1) nlvl1++; }
      else { if (vsi > 0.) { if (vsi < vs1) nlvl1++; } }
    }
  }
  G.nlvl1 = nlvl1;
  int lvlast = no_lvl == 0 ? 1 : last_lvl; // lvls(no_lvl) BEFORE the sort by vp: the deepest one
  if (ifs + 2 > lvlast) lvlast = ifs + 2;
  G.lvlast = lvlast;
  G.dc = P.dc;
  G.overflow = 0;
  return n;
}

// One warp per column; a block (= warp) is persistent and fetches columns from the list until it is empty -- the grid is sized
// to the warps the GPU can hold (218 registers: 8 per SM), so 4 096 columns are one launch with no under-filled chunk tails.
__global__ void __launch_bounds__(32) grt_kernel(const __grid_constant__ GrtParams P) {
  mct_exptab_stage();
  __shared__ GrtCol G;
  __shared__ int s_n, s_next;
  __shared__ double s_sort[GRT_SSORT];
  const int lane = threadIdx.x;
  double* base = P.scratch + (size_t)blockIdx.x * GRT_SCRATCH;
  double* lay = base;                                   // GRT_LAY * (MAX_LAYERS + 2)
  double* v = lay + GRT_LAY * (MCT_MAX_LAYERS + 2);     // 2 * (MAX_LAYERS + 2)
  double* vvv = v + 2 * (MCT_MAX_LAYERS + 2);           // GRT_NVPAD (1-based inside)
  double* ccc = vvv + GRT_NVPAD;                        // GRT_NV + 8
  const double pi_m = (double)3.1415926f; // m_surfmodes' pi
  const double dh = (double)0.005f;
  const int np = P.np;
  for (;;) {
    __syncwarp();
    if (lane == 0) s_next = atomicAdd(P.next, 1);
    __syncwarp();
    const int k = s_next;
    if (k >= P.nlist) break;
    const int col = P.list[k];
    if (P.skip && P.skip[2 * (col / P.cols_per_model)] != 0) continue; // a model check_model rejected: nothing is solved
    if (lane == 0) {
      G.vvv = vvv - 0; G.ccc = ccc;
      s_n = g_setup(P, col, G, lay, v);
    }
    __syncwarp();
    double* pv = P.pvel + (size_t)col * P.np;
    double* gv = P.gvel + (size_t)col * P.np;
    if (s_n == 0) { // stays as the dispersion kernel left it (ierr = 2, preset values)
      if (lane == 0 && P.flags) atomicMax(&P.flags[2 * (col / P.cols_per_model) + 1], 2);
      continue;
    }
    unsigned nsec = 0, nlay = 0;
    int ierr = 0;
    double c0 = 0;
    for (int i = 1; i <= np; ++i) {
      if (lane == 0) {
        G.w = P.freqs[i - 1] * 2 * pi_m;
        G.tol = P.tolmin + (np + 1 - i) * (P.tolmax - P.tolmin) / np;
        G.smin = P.smin_min + (i - 1) * (P.smin_max - P.smin_min) / np;
      }
      __syncwarp();
      double root = 0;
      const int ierr1 = g_search_one(G, c0, &root, lane, nsec, nlay, s_sort);
      if (ierr1 == 1) { ierr = 1; break; }
      if (lane == 0) pv[i - 1] = root;
      c0 = root;
      if (P.phaseGroup == 1) {
        const double freq0 = P.freqs[i - 1] + dh;
        __syncwarp();
        if (lane == 0) G.w = freq0 * 2 * pi_m;
        __syncwarp();
        double root0 = 0;
        ierr = g_search_one(G, c0, &root0, lane, nsec, nlay, s_sort);
        if (ierr == 1) break;
        const double gg = (P.freqs[i - 1] + dh) / root0 - P.freqs[i - 1] / root;
        if (lane == 0) gv[i - 1] = gg > 0 ? dh / gg : 0;
      }
      __syncwarp();
    }
    if (lane == 0) {
      P.ierr[col] = ierr;
      if (P.counters) { atomicAdd(&P.counters[0], (unsigned long long)nsec); atomicAdd(&P.counters[1], (unsigned long long)nlay); }
    }
  }
}

// the columns whose status says "low-velocity layer": list + count
__global__ void grt_collect_kernel(const int32_t* status, int ncol, int32_t* list, int32_t* count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool hit = c < ncol && status[c] == 2;
  const unsigned bal = __ballot_sync(0xffffffffu, hit);
  if (bal == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(bal) - 1) base = atomicAdd(count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  if (hit) list[base + __popc(bal & ((1u << lane) - 1u))] = c;
}
