/* oracle/ref_harness/love_f90_harness.c -- TEST INFRASTRUCTURE.  Driver around the mechanical translation of the reference's
 * surfmodes/Love.f90 (oracle/f90toc_love.py -> oracle/_ref/love_f2c.c, included below).  It fills a T_GRT with what setup_grt
 * and startl leave there for one column and calls SecFuns_L(1 + ifs, c, GRT, Imf) as SearchLove does. */
#include LOVE_F2C_SOURCE
#include <string.h>

/* the dummy procedure of bisecim: SearchLove passes SecFuns_L */
static double f_(void* ilay, void* c, void* grt, void* imf) { return secfuns_l_(ilay, c, grt, imf); }

int ref_love_secfun(int n, const double* d, const double* vs, const double* mu, int ifs, int ll, double w, double c, double* value,
                    double* imf) {
  T_GRT g;
  memset(&g, 0, sizeof g);
  g.nlayers = n;
  g.d = (double*)d; g.d_d1 = n; g.d_l1 = 1;
  g.vs = (double*)vs; g.vs_d1 = n; g.vs_l1 = 1;
  g.mu = (double*)mu; g.mu_d1 = n; g.mu_l1 = 1;
  g.ifs = ifs; g.ll = ll; g.w = w;
  int lay = 1 + ifs;
  init_love_(&n);
  *value = secfuns_l_(&lay, &c, &g, imf);
  delete_love_();
  return 0;
}

/* one root refinement as SearchLove's FundaMode issues it: the secular function at both ends of a bracket (ll as startl left it
 * for the upper end: supplied), bisecim.  out = {root, f1, f2}; returns iq */
int ref_love_bisecim(int n, const double* d, const double* vs, const double* mu, int ifs, int ll, double w, double k1, double k2,
                     double smin, double tol, double* out) {
  T_GRT g;
  memset(&g, 0, sizeof g);
  g.nlayers = n;
  g.d = (double*)d; g.d_d1 = n; g.d_l1 = 1;
  g.vs = (double*)vs; g.vs_d1 = n; g.vs_l1 = 1;
  g.mu = (double*)mu; g.mu_d1 = n; g.mu_l1 = 1;
  g.ifs = ifs; g.ll = ll; g.w = w; g.smin = smin; g.tol = tol;
  int lay = 1 + ifs, iq = -1;
  double imf = 0;
  init_love_(&n);
  double f1 = secfuns_l_(&lay, &k1, &g, &imf), f2 = secfuns_l_(&lay, &k2, &g, &imf);
  out[0] = bisecim_(0, &lay, &k1, &k2, &f1, &f2, &g, &iq);
  out[1] = f1; out[2] = f2;
  delete_love_();
  return iq;
}

/* the trial phase velocities of one frequency: C_Interval_L on the T_GRT setup_grt leaves (v: the sorted layer velocities,
 * nv of them, 2n allocated).  ccc: 20000 doubles.  counts = {ncc, im1}. */
int ref_love_cinterval(int n, const double* d, const double* vp, const double* vs, const double* v, int nv, double vsy, double vsm, double vs1,
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
  c_interval_l_(&g, ccc, &ncc, &im1);
  counts[0] = ncc; counts[1] = im1;
  free(vv);
  return 0;
}
