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

/* A whole column, phase velocities of the fundamental Love mode -- surfmodes with modetype = 0 for a column with a low-velocity
 * layer.  Written out here: init_grt (GRT.f90:44-91), surfmodes' assignments and dispatch (surfmodes.f90:57-99), the frequency
 * loop of LoveModes (:266-281) and the allmodes = 0 path of SearchLove (SearchLove.f90:24-40,179).  Translated, i.e. the
 * reference's own statements: setup_grt, C_Interval_L (N_cf_L, sort), init_love, FundaMode (check, startl, SecFuns_L and below,
 * bisecim), delete_love.  par = {tolmin, tolmax, smin_min, smin_max, dcm, dc2}; dc as surfmodes' caller sets it.
 * Returns ierr; -1 when setup_grt STOPs; -2 when the column has no low-velocity layer (surfdisp96's case). */
int ref_love_modes(int n, const double* thick, const double* vp, const double* vs, const double* rho, int nf, const double* freqs, double dc,
                   const double* par, double* phase, double* group /* or null: phase velocities only */) {
  T_GRT g;
  T_MODES_PARA para;
  memset(&g, 0, sizeof g);
  memset(&para, 0, sizeof para);
  para.modetype = 0; para.tolmin = par[0]; para.tolmax = par[1]; para.smin_min = par[2]; para.smin_max = par[3];
  para.dc = dc; para.dcm = par[4]; para.dc2 = par[5];
  g.nlayers = n;
  g.smin = (double)1E-4f; g.tol = (double)1E-5f; g.dc = (double)1E-4f; g.dc2 = (double)1E-4f; g.dcm = (double)1E-4f;
  double* buf = (double*)calloc((size_t)8 * n + 16, sizeof(double));
  int* lv = (int*)calloc((size_t)n / 2 + 8, sizeof(int));
  double* ccc = (double*)calloc(20008, sizeof(double));
  g.d = buf; g.vp = buf + n; g.vs = buf + 2 * n; g.rho = buf + 3 * n; g.mu = buf + 4 * n; g.v = buf + 5 * n;
  g.d_d1 = g.vp_d1 = g.vs_d1 = g.rho_d1 = g.mu_d1 = n; g.v_d1 = 2 * n;
  g.d_l1 = g.vp_l1 = g.vs_l1 = g.rho_l1 = g.mu_l1 = g.v_l1 = 1;
  g.lvls = lv; g.lvls_d1 = n / 2 + 1; g.lvls_l1 = 1;
  g.vsy = 1.7976931348623157e308;
  for (int i = 0; i < n; ++i) { g.d[i] = thick[i]; g.vp[i] = vp[i]; g.vs[i] = vs[i]; g.rho[i] = rho[i]; }
  f90_stopped = 0;
  setup_grt_(&g, &para);
  int ierr = 0;
  if (f90_stopped) ierr = -1;
  else if (g.nlvls1 == 0) ierr = -2;
  else {
    grt = &g;                                             /* the host variable the internal procedure `check` reads */
    double c0 = 0;
    for (int i = 1; i <= nf; ++i) {
      g.w = freqs[i - 1] * 2 * pi_8;                      /* m_surfmodes' pi = 3.1415926 */
      g.tol = para.tolmin + (nf + 1 - i) * (para.tolmax - para.tolmin) / nf;
      g.smin = para.smin_min + (i - 1) * (para.smin_max - para.smin_min) / nf;
      g.index_a = i;
      int index0 = 0, im1 = 0, ierr1 = 0;
      for (int k = 0; k < 20000; ++k) ccc[k] = 0;         /* ccc = 0 */
      c_interval_l_(&g, ccc, &index0, &im1);
      init_love_(&g.nlayers);
      double cray = c0;
      fundamode_(&g, ccc, &index0, &cray, &ierr1);
      delete_love_();
      if (ierr1 == 1) { ierr = 1; break; }
      phase[i - 1] = cray;
      c0 = cray;
      if (group) {                                        /* paras%phaseGroup == 1 (LoveModes :282-290): the search again at freq + dh */
        double dh = (double)0.005f, freq0 = freqs[i - 1] + dh, f = freqs[i - 1];
        g.w = freq0 * 2 * pi_8;
        index0 = 0; im1 = 0;
        for (int k = 0; k < 20000; ++k) ccc[k] = 0;
        c_interval_l_(&g, ccc, &index0, &im1);
        init_love_(&g.nlayers);
        double cp0 = c0;
        fundamode_(&g, ccc, &index0, &cp0, &ierr);
        delete_love_();
        if (ierr == 1) break;
        calgroup_(&phase[i - 1], &cp0, &f, &dh, &group[i - 1]);
      }
    }
  }
  free(buf); free(lv); free(ccc);
  return ierr;
}
