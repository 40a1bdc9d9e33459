/* oracle/ref_harness/rayleigh_f90_harness.c -- TEST INFRASTRUCTURE.  Driver around the mechanical translation of the part of the
 * reference's surfmodes/Rayleigh.f90 that a column without a water layer reaches (oracle/f90toc_love.py ->
 * oracle/_ref/rayleigh_f2c.c, included below): startl, SecFunSurf, propup, EinvE, inv2.  It fills a T_GRT with what setup_grt
 * leaves there, lets the translated startl choose the deepest layer (GRT%ll) and calls SecFunSurf(0, c, GRT, Imf) as
 * SearchRayleigh does. */
#include RAYLEIGH_F2C_SOURCE
#include <string.h>

/* the dummy procedure of bisecim: SearchRayleigh passes SecFunSurf for a column without water (FundaMode), SecFunSt for one with a
 * water layer on top (StMode) */
static __thread int g_f_st;
static double f_(void* ilay, void* c, void* grt, void* imf) { return g_f_st ? secfunst_(ilay, c, grt, imf) : secfunsurf_(ilay, c, grt, imf); }

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

/* setup_grt on a fresh T_GRT: the allocations and zero / huge() initial values are init_grt's (GRT.f90:44-91), the assignments of
 * the column to GRT%d, %vp, %vs, %rho are surfmodes' (surfmodes.f90:60-75); everything else is the translated setup_grt.
 * Outputs as orc_grt_setup.  Returns 1 when setup_grt STOPs (a fluid layer below the first). */
int ref_setup_grt(int n, const double* thick, const double* vp, const double* vs, const double* rho, int modetype, double* mu_out,
                  double* v_out, int* lvls_out, int* ints, double* dbl) {
  T_GRT g;
  T_MODES_PARA para;
  memset(&g, 0, sizeof g);
  memset(&para, 0, sizeof para);
  para.modetype = modetype; para.dc = 1e-3; para.dcm = 1e-3; para.dc2 = 1e-3;
  g.nlayers = n;
  g.smin = (double)1E-4f; g.tol = (double)1E-5f; g.dc = (double)1E-4f; g.dc2 = (double)1E-4f; g.dcm = (double)1E-4f;
  double* buf = (double*)calloc((size_t)8 * n + 16, sizeof(double));
  int* lv = (int*)calloc((size_t)n / 2 + 8, sizeof(int));
  g.d = buf; g.vp = buf + n; g.vs = buf + 2 * n; g.rho = buf + 3 * n; g.mu = buf + 4 * n; g.v = buf + 5 * n;
  g.d_d1 = g.vp_d1 = g.vs_d1 = g.rho_d1 = g.mu_d1 = n; g.v_d1 = 2 * n;
  g.d_l1 = g.vp_l1 = g.vs_l1 = g.rho_l1 = g.mu_l1 = g.v_l1 = 1;
  g.lvls = lv; g.lvls_d1 = n / 2 + 1; g.lvls_l1 = 1;
  g.vsy = 1.7976931348623157e308; /* huge(grt%vsy) */
  for (int i = 0; i < n; ++i) { g.d[i] = thick[i]; g.vp[i] = vp[i]; g.vs[i] = vs[i]; g.rho[i] = rho[i]; }
  f90_stopped = 0;
  setup_grt_(&g, &para);
  int nv = 0;
  for (int i = 0; i < n; ++i) nv += (vs[i] > 1e-6 || vs[i] < -1e-6) ? 2 : 1;
  for (int i = 0; i < n; ++i) mu_out[i] = g.mu[i];
  for (int i = 0; i < 2 * n; ++i) v_out[i] = g.v[i];
  for (int i = 0; i < n / 2 + 1; ++i) lvls_out[i] = lv[i];
  ints[0] = g.ifs; ints[1] = g.no_lvl; ints[2] = g.no_lvl_fl; ints[3] = g.nlvl1; ints[4] = g.nlvls1; ints[5] = g.lvlast; ints[6] = g.l1;
  ints[7] = nv;
  dbl[0] = g.mu0; dbl[1] = g.vsy; dbl[2] = g.vs1; dbl[3] = g.vsm; dbl[4] = g.vss1;
  free(buf); free(lv);
  return f90_stopped;
}

/* A whole column WITHOUT a water layer, phase velocities of the fundamental Rayleigh mode -- surfmodes with modetype = 1 for a
 * column with a low-velocity layer.  Written out here: init_grt (GRT.f90:44-91), surfmodes' assignments and dispatch
 * (surfmodes.f90:57-99), the frequency loop of RayleighModes (:209-221) and the allmodes = 0, ifs = 0 path of SearchRayleigh
 * (SearchRayleigh.f90:26-33,60-64,263).  Translated, i.e. the reference's own statements: setup_grt, C_Interval (N_cf, sort),
 * init_rayleigh, FundaMode (CR0_Finder with Rayhomo, startl, SecFunSurf and below, bisecim), delete_rayleigh.
 * With a water layer on top (ifs = 1): St_Finder (with its internal getSt) when there is no previous root, then StMode
 * (:65-68), over SecFunSt, Stoneley, propdn_f, EinvE_f, det3 -- translated as well.
 * par = {tolmin, tolmax, smin_min, smin_max, dcm, dc2}.  Returns ierr; -1: setup_grt or St_Finder STOPs; -2: no low-velocity
 * layer (surfdisp96's column). */
int ref_rayleigh_modes(int n, const double* thick, const double* vp, const double* vs, const double* rho, int nf, const double* freqs,
                       double dc, const double* par, double* phase, double* group /* or null: phase velocities only */) {
  T_GRT g;
  T_MODES_PARA para;
  memset(&g, 0, sizeof g);
  memset(&para, 0, sizeof para);
  para.modetype = 1; para.tolmin = par[0]; para.tolmax = par[1]; para.smin_min = par[2]; para.smin_max = par[3];
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
    double c0 = 0;
    g_f_st = g.ifs != 0;
    for (int i = 1; i <= nf; ++i) {
      g.w = freqs[i - 1] * 2 * pi_8;                      /* m_surfmodes' pi = 3.1415926 */
      g.tol = para.tolmin + (nf + 1 - i) * (para.tolmax - para.tolmin) / nf;
      g.smin = para.smin_min + (i - 1) * (para.smin_max - para.smin_min) / nf;
      g.index_a = i;
      int index0 = 0, im1 = 0, ierr1 = 0;
      for (int k = 0; k < 20000; ++k) ccc[k] = 0;         /* ccc = 0 */
      c_interval_(&g, ccc, &index0, &im1);
      init_rayleigh_(&g.nlayers);
      double cray = c0;
      if (g.ifs == 0) fundamode_(&g, ccc, &index0, &im1, &cray, &ierr1);
      else {
        if (cray <= 0) st_finder_(&g.ifs, &g, &cray);
        if (f90_stopped) { delete_rayleigh_(); ierr = -1; break; }
        stmode_(&g, ccc, &index0, &im1, &cray, &ierr1);
      }
      delete_rayleigh_();
      if (ierr1 == 1) { ierr = 1; break; }
      phase[i - 1] = cray;
      c0 = cray;
      if (group) {                                        /* paras%phaseGroup == 1 (RayleighModes :226-234): the search again at freq + dh */
        double dh = (double)0.005f, freq0 = freqs[i - 1] + dh, f = freqs[i - 1];
        g.w = freq0 * 2 * pi_8;
        index0 = 0; im1 = 0;
        for (int k = 0; k < 20000; ++k) ccc[k] = 0;
        c_interval_(&g, ccc, &index0, &im1);
        init_rayleigh_(&g.nlayers);
        double cp0 = c0;
        if (g.ifs == 0) fundamode_(&g, ccc, &index0, &im1, &cp0, &ierr);
        else {
          if (cp0 <= 0) st_finder_(&g.ifs, &g, &cp0);
          stmode_(&g, ccc, &index0, &im1, &cp0, &ierr);
        }
        delete_rayleigh_();
        if (ierr == 1) break;
        calgroup_(&phase[i - 1], &cp0, &f, &dh, &group[i - 1]);
      }
    }
  }
  free(buf); free(lv); free(ccc);
  return ierr;
}

/* the secular function of a column WITH a water layer on top (ifs = 1): startl + SecFunSt(ifs, c, GRT, Imf) -- Stoneley, propdn_f,
 * EinvE_f, propup, EinvE, det3 all translated */
int ref_rayleigh_secfunst(int n, const double* d, const double* vp, const double* vs, const double* rho, const double* mu, double mu0,
                          int ifs, int lvlast, double w, double c, double* value, double* imf, int* ll_out) {
  T_GRT g;
  memset(&g, 0, sizeof g);
  g.nlayers = n;
  g.d = (double*)d; g.d_d1 = n; g.d_l1 = 1;
  g.vp = (double*)vp; g.vp_d1 = n; g.vp_l1 = 1;
  g.vs = (double*)vs; g.vs_d1 = n; g.vs_l1 = 1;
  g.rho = (double*)rho; g.rho_d1 = n; g.rho_l1 = 1;
  g.mu = (double*)mu; g.mu_d1 = n; g.mu_l1 = 1;
  g.mu0 = mu0; g.ifs = ifs; g.lvlast = lvlast; g.w = w;
  init_rayleigh_(&n);
  startl_(&c, &g);
  *ll_out = g.ll;
  *value = secfunst_(&ifs, &c, &g, imf);
  delete_rayleigh_();
  return 0;
}
