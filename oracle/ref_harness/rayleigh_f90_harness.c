/* oracle/ref_harness/rayleigh_f90_harness.c -- TEST INFRASTRUCTURE.  Driver around the mechanical translation of the part of the
 * reference's surfmodes/Rayleigh.f90 that a column without a water layer reaches (oracle/f90toc_love.py ->
 * oracle/_ref/rayleigh_f2c.c, included below): startl, SecFunSurf, propup, EinvE, inv2.  It fills a T_GRT with what setup_grt
 * leaves there, lets the translated startl choose the deepest layer (GRT%ll) and calls SecFunSurf(0, c, GRT, Imf) as
 * SearchRayleigh does. */
#include RAYLEIGH_F2C_SOURCE
#include <string.h>

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
