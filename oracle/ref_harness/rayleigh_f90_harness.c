/* oracle/ref_harness/rayleigh_f90_harness.c -- TEST INFRASTRUCTURE.  Driver around the mechanical translation of the part of the
 * reference's surfmodes/Rayleigh.f90 that a column without a water layer reaches (oracle/f90toc_love.py ->
 * oracle/_ref/rayleigh_f2c.c, included below): startl, SecFunSurf, propup, EinvE, inv2.  It fills a T_GRT with what setup_grt
 * leaves there, lets the translated startl choose the deepest layer (GRT%ll) and calls SecFunSurf(0, c, GRT, Imf) as
 * SearchRayleigh does. */
#include RAYLEIGH_F2C_SOURCE
#include <string.h>

/* the dummy procedure of bisecim: SearchRayleigh passes SecFunSurf for a column without water (SearchRayleigh.f90 FundaMode) */
static double f_(void* ilay, void* c, void* grt, void* imf) { return secfunsurf_(ilay, c, grt, imf); }

int ref_rayleigh_secfunsurf(int n, const double* d, const double* vp, const double* vs, const double* mu, int lvlast, double w, double c,
                            double* value, double* imf, int* ll_out) {
  T_GRT g;
  memset(&g, 0, sizeof g);
  g.nlayers = n;
  g.d = (double*)d; g.d_d1 = n; g.d_l1 = 1;
  g.vp = (double*)vp; g.vp_d1 = n; g.vp_l1 = 1;
  g.vs = (double*)vs; g.vs_d1 = n; g.vs_l1 = 1;
  g.mu = (double*)mu; g.mu_d1 = n; g.mu_l1 = 1;
  g.ifs = 0; g.lvlast = lvlast; g.w = w;
  int isurf = 0;
  init_rayleigh_(&n);
  startl_(&c, &g);
  *ll_out = g.ll;
  *value = secfunsurf_(&isurf, &c, &g, imf);
  delete_rayleigh_();
  return 0;
}

/* one root refinement as FundaMode issues it: startl at k2, SecFunSurf at both ends, bisecim.  out = {root, f1, f2}; returns iq */
int ref_rayleigh_bisecim(int n, const double* d, const double* vp, const double* vs, const double* mu, int lvlast, double w, double k1,
                         double k2, double smin, double tol, double* out) {
  T_GRT g;
  memset(&g, 0, sizeof g);
  g.nlayers = n;
  g.d = (double*)d; g.d_d1 = n; g.d_l1 = 1;
  g.vp = (double*)vp; g.vp_d1 = n; g.vp_l1 = 1;
  g.vs = (double*)vs; g.vs_d1 = n; g.vs_l1 = 1;
  g.mu = (double*)mu; g.mu_d1 = n; g.mu_l1 = 1;
  g.ifs = 0; g.lvlast = lvlast; g.w = w; g.smin = smin; g.tol = tol;
  int isurf = 0, iq = -1;
  double imf = 0;
  init_rayleigh_(&n);
  startl_(&k2, &g);
  double f1 = secfunsurf_(&isurf, &k1, &g, &imf), f2 = secfunsurf_(&isurf, &k2, &g, &imf);
  out[0] = bisecim_(0, &isurf, &k1, &k2, &f1, &f2, &g, &iq);
  out[1] = f1; out[2] = f2;
  delete_rayleigh_();
  return iq;
}

/* the trial phase velocities of one frequency: C_Interval on the T_GRT setup_grt leaves (v: the sorted layer velocities,
 * nv of them, 2n allocated).  ccc: 20000 doubles.  counts = {ncc, im1}. */
int ref_rayleigh_cinterval(int n, const double* d, const double* vp, const double* vs, const double* v, int nv, double vsy, double vsm, double vs1,
            double w, double tol, double* ccc, int* counts) {
  T_GRT g;
  memset(&g, 0, sizeof g);
  g.nlayers = n;
  g.d = (double*)d; g.d_d1 = n; g.d_l1 = 1;
  g.vp = (double*)vp; g.vp_d1 = n; g.vp_l1 = 1;
  g.vs = (double*)vs; g.vs_d1 = n; g.vs_l1 = 1;
  double* vv = (double*)calloc((size_t)2 * n + 8, sizeof(double));   /* allocate( GRT%v(2*nlayers) ); GRT%v = 0 */
  for (int i = 0; i < nv; ++i) vv[i] = v[i];
  g.v = vv; g.v_d1 = 2 * n; g.v_l1 = 1;
  g.vsy = vsy; g.vsm = vsm; g.vs1 = vs1; g.w = w; g.tol = tol;
  int ncc = 0, im1 = 0;
  c_interval_(&g, ccc, &ncc, &im1);
  counts[0] = ncc; counts[1] = im1;
  free(vv);
  return 0;
}
