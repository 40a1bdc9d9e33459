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
