/* oracle/ref_harness/love_f90_harness.c -- TEST INFRASTRUCTURE.  Driver around the mechanical translation of the reference's
 * surfmodes/Love.f90 (oracle/f90toc_love.py -> oracle/_ref/love_f2c.c, included below).  It fills a T_GRT with what setup_grt
 * and startl leave there for one column and calls SecFuns_L(1 + ifs, c, GRT, Imf) as SearchLove does. */
#include LOVE_F2C_SOURCE
#include <string.h>

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
