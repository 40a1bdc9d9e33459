/*
 * oracle/fm2d_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's 2-D fast-marching eikonal solver as `surf_likelihood` drives it for
 * phase-velocity data (src/likelihood_surf.F90:295-336 with uar = 1: travel times only, no ray paths):
 *   modrays            fm2d/fm2dray_cartesian.f90:67-478   (source loop, source-grid refinement, mapping back)
 *   gridder            :490-590                             (cubic B-spline dicing of the velocity vertices)
 *   bsplrefine         :598-668                             (B-spline velocities on the refined source grid)
 *   srtimes            :676-770                             (bilinear receiver times; near-source special case)
 *   travel, fouds1, fouds2, addtree, downtree, updtree, bilinear   fm2d/fm2d_ttime.f90
 * Not restated: the `dynamic` restart (switched off in the reference itself, fm2dray_cartesian.f90:251-252).
 *
 * PARITY STATUS: no Fortran compiler in this image, no golden values in the reference.
 *   PINNED on the reference's own source: all of fm2d_ttime.f90 (travel, fouds1, fouds2, addtree, downtree, updtree,
 *   bilinear), gridder / bsplrefine / srtimes / rpaths and the body of modrays' source loop, translated statement by
 *   statement to C by oracle/f90toc.py (oracle/_ref/libfm2d_ttime_f2c.so) -- bit-identical per routine (fields, node
 *   status, heap; urg 0/1/2; both operator orders; ties; rough media) and for whole calls of orc_fm2d_times and
 *   orc_fm2d_rays (every receiver time, field and ray point): tests/test_oracle_fm2d_vs_reference.py, committed fixtures +
 *   fresh random cases.  Deliberately different where the Fortran is undefined (flagged instead, see rpaths below and
 *   the test): 0/0 gradient, btg overrun, ipzr unassigned with asgr = 0.
 * All arithmetic is double (REAL(KIND=i10) = c_double); default-real literals in the source are exactly representable;
 * x**2 is x*x, x**3 is (x*x)*x as gfortran expands integer powers.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int nvx, nvz, nnx, nnz, gdx, gdz, fom, sgdl;
  int vnl, vnr, vnt, vnb;
  double gox, goz, dnx, dnz, dvx, dvz;
  int ld;                 /* leading dimension of veln / ttn / nsts (rows = z) */
  double *velv, *veln, *velnb, *ttn, *ttnr;
  int *nsts, *nstsr;
  int *bpx, *bpz;         /* btg(:)%px, %pz */
  int ntr, maxbt;
  size_t cells;           /* allocated nodes of veln / ttn / nsts */
  int64_t n_accept, n_update; /* instrumentation: nodes made alive, fouds calls */
  int nnxr, nnzr; double goxr, gozr, dnxr, dnzr; /* "Backup refined grid information" (fm2dray_cartesian.f90:374-381) */
  int error;
} fm_t;

#define VELV(i, j) F->velv[(size_t)(j) * (F->nvz + 2) + (i)]           /* velv(0:nvz+1, 0:nvx+1) */
#define VELN(k, j) F->veln[(size_t)((j) - 1) * F->ld + ((k) - 1)]
#define VELNB(k, j) F->velnb[(size_t)((j) - 1) * F->ld + ((k) - 1)]
#define TTN(k, j) F->ttn[(size_t)((j) - 1) * F->ld + ((k) - 1)]
#define TTNR(k, j) F->ttnr[(size_t)((j) - 1) * F->ld + ((k) - 1)]
#define NSTS(k, j) F->nsts[(size_t)((j) - 1) * F->ld + ((k) - 1)]
#define NSTSR(k, j) F->nstsr[(size_t)((j) - 1) * F->ld + ((k) - 1)]

static inline double sq(double x) { return x * x; }

/* test hook: where set (per calling thread), fm2d_core writes for every source the number of nodes its march left unreached
 * (status != alive); such a source's field is the previous source's in the Fortran (see the comment in fm2d_core) */
static __thread int* g_unreached = 0;
void orc_fm2d_set_unreached(int* buf) { g_unreached = buf; }

/* heap: fm2d_ttime.f90 addtree / downtree / updtree */
static void sift_up(fm_t* F, int iz, int ix, int tpc) {
  int tpp = tpc / 2;
  while (tpp > 0) {
    if (TTN(iz, ix) < TTN(F->bpz[tpp], F->bpx[tpp])) {
      NSTS(iz, ix) = tpp;
      NSTS(F->bpz[tpp], F->bpx[tpp]) = tpc;
      int ex = F->bpx[tpc], ez = F->bpz[tpc];
      F->bpx[tpc] = F->bpx[tpp]; F->bpz[tpc] = F->bpz[tpp];
      F->bpx[tpp] = ex; F->bpz[tpp] = ez;
      tpc = tpp;
      tpp = tpc / 2;
    } else tpp = 0;
  }
}
static void addtree(fm_t* F, int iz, int ix) {
  if (F->ntr + 1 > F->maxbt) { F->error = 2; return; } /* the Fortran would overrun btg(maxbt) */
  F->ntr++;
  NSTS(iz, ix) = F->ntr;
  F->bpx[F->ntr] = ix; F->bpz[F->ntr] = iz;
  sift_up(F, iz, ix, F->ntr);
}
static void updtree(fm_t* F, int iz, int ix) { sift_up(F, iz, ix, NSTS(iz, ix)); }
static void swap_nodes(fm_t* F, int tpp, int tpc) {
  NSTS(F->bpz[tpp], F->bpx[tpp]) = tpc;
  NSTS(F->bpz[tpc], F->bpx[tpc]) = tpp;
  int ex = F->bpx[tpc], ez = F->bpz[tpc];
  F->bpx[tpc] = F->bpx[tpp]; F->bpz[tpc] = F->bpz[tpp];
  F->bpx[tpp] = ex; F->bpz[tpp] = ez;
}
static void downtree(fm_t* F) {
  if (F->ntr == 1) { F->ntr--; return; }
  NSTS(F->bpz[F->ntr], F->bpx[F->ntr]) = 1;
  F->bpx[1] = F->bpx[F->ntr]; F->bpz[1] = F->bpz[F->ntr];
  F->ntr--;
  int tpp = 1, tpc = 2;
  while (tpc < F->ntr) {
    double rd1 = TTN(F->bpz[tpc], F->bpx[tpc]), rd2 = TTN(F->bpz[tpc + 1], F->bpx[tpc + 1]);
    if (rd1 > rd2) tpc = tpc + 1;
    rd1 = TTN(F->bpz[tpc], F->bpx[tpc]);
    rd2 = TTN(F->bpz[tpp], F->bpx[tpp]);
    if (rd1 < rd2) { swap_nodes(F, tpp, tpc); tpp = tpc; tpc = 2 * tpp; }
    else tpc = F->ntr + 1;
  }
  if (tpc == F->ntr) {
    double rd1 = TTN(F->bpz[tpc], F->bpx[tpc]), rd2 = TTN(F->bpz[tpp], F->bpx[tpp]);
    if (rd1 < rd2) swap_nodes(F, tpp, tpc);
  }
}

/* the quadratic of every stencil: rd1 = b**2-4.0*a*c; clamp; tdsh = (-b+sqrt(rd1))/(2.0*a) */
static inline double qsolve(double a, double b, double c) {
  double rd1 = b * b - 4.0 * a * c;
  if (rd1 < 0.0) rd1 = 0.0;
  return (-b + sqrt(rd1)) / (2.0 * a);
}

/* fouds1: first-order upwind update of ttn(iz,ix) (fm2d_ttime.f90:138-197) */
static void fouds1(fm_t* F, int iz, int ix) {
  int tsw1 = 0;
  double travm = 0, slown = 1.0 / VELN(iz, ix);
  const double dnx = F->dnx, dnz = F->dnz;
  F->n_update++;
  for (int j = ix - 1; j <= ix + 1; j += 2)
    for (int k = iz - 1; k <= iz + 1; k += 2) {
      if (j < 1 || j > F->nnx || k < 1 || k > F->nnz) continue;
      int swsol = 0;
      double a = 0, b = 0, c = 0, tref = 0;
      if (NSTS(iz, j) == 0) {
        swsol = 1;
        if (NSTS(k, ix) == 0) {
          double u = dnx, v = dnz, em = TTN(k, ix) - TTN(iz, j);
          a = u * u + v * v;
          b = -2.0 * (u * u) * em;
          c = (u * u) * (em * em - (v * v) * (slown * slown));
          tref = TTN(iz, j);
        } else { a = 1.0; b = 0.0; c = -(slown * slown) * (dnx * dnx); tref = TTN(iz, j); }
      } else if (NSTS(k, ix) == 0) {
        swsol = 1;
        a = 1.0; b = 0.0; c = -sq(slown * dnz); tref = TTN(k, ix);
      }
      if (swsol) {
        double trav = tref + qsolve(a, b, c);
        if (tsw1) travm = trav < travm ? trav : travm; else { travm = trav; tsw1 = 1; }
      }
    }
  TTN(iz, ix) = travm;
}

/* fouds2: mixed-order update (fm2d_ttime.f90:199-345) */
static void fouds2(fm_t* F, int iz, int ix) {
  int tsw1 = 0;
  double travm = 0, slown = 1.0 / VELN(iz, ix);
  const double dnx = F->dnx, dnz = F->dnz;
  F->n_update++;
  for (int j = ix - 1; j <= ix + 1; j += 2) {
    if (j < 1 || j > F->nnx) continue;
    int swj = -1, j2;
    if (j == ix - 1) { j2 = j - 1; if (j2 >= 1) { if (NSTS(iz, j2) == 0) swj = 0; } }
    else { j2 = j + 1; if (j2 <= F->nnx) { if (NSTS(iz, j2) == 0) swj = 0; } }
    if (NSTS(iz, j) == 0 && swj == 0) { swj = -1; if (TTN(iz, j) > TTN(iz, j2)) swj = 0; }
    else swj = -1;
    for (int k = iz - 1; k <= iz + 1; k += 2) {
      if (k < 1 || k > F->nnz) continue;
      int swk = -1, k2;
      if (k == iz - 1) { k2 = k - 1; if (k2 >= 1) { if (NSTS(k2, ix) == 0) swk = 0; } }
      else { k2 = k + 1; if (k2 <= F->nnz) { if (NSTS(k2, ix) == 0) swk = 0; } }
      if (NSTS(k, ix) == 0 && swk == 0) { swk = -1; if (TTN(k, ix) > TTN(k2, ix)) swk = 0; }
      else swk = -1;
      int swsol = 0;
      double a = 0, b = 0, c = 0, tref = 0, tdiv = 1.0, u, v, em;
      if (swj == 0) {
        swsol = 1;
        if (swk == 0) {
          u = 2.0 * dnx; v = 2.0 * dnz;
          em = 4.0 * TTN(iz, j) - TTN(iz, j2) - 4.0 * TTN(k, ix);
          em = em + TTN(k2, ix);
          a = v * v + u * u;
          b = 2.0 * em * (u * u);
          c = (u * u) * (em * em - (slown * slown) * (v * v));
          tref = 4.0 * TTN(iz, j) - TTN(iz, j2);
          tdiv = 3.0;
        } else if (NSTS(k, ix) == 0) {
          u = dnz; v = 2.0 * dnx;
          em = 3.0 * TTN(k, ix) - 4.0 * TTN(iz, j) + TTN(iz, j2);
          a = v * v + 9.0 * (u * u);
          b = 6.0 * em * (u * u);
          c = (u * u) * (em * em - (slown * slown) * (v * v));
          tref = TTN(k, ix);
          tdiv = 1.0;
        } else {
          u = 2.0 * dnx;
          a = 1.0; b = 0.0; c = -(u * u) * (slown * slown);
          tref = 4.0 * TTN(iz, j) - TTN(iz, j2);
          tdiv = 3.0;
        }
      } else if (NSTS(iz, j) == 0) {
        swsol = 1;
        if (swk == 0) {
          u = dnx; v = 2.0 * dnz;
          em = 3.0 * TTN(iz, j) - 4.0 * TTN(k, ix) + TTN(k2, ix);
          a = v * v + 9.0 * (u * u);
          b = 6.0 * em * (u * u);
          c = (u * u) * (em * em - (v * v) * (slown * slown));
          tref = TTN(iz, j);
          tdiv = 1.0;
        } else if (NSTS(k, ix) == 0) {
          u = dnx; v = dnz;
          em = TTN(k, ix) - TTN(iz, j);
          a = u * u + v * v;
          b = -2.0 * (u * u) * em;
          c = (u * u) * (em * em - (v * v) * (slown * slown));
          tref = TTN(iz, j);
          tdiv = 1.0;
        } else { a = 1.0; b = 0.0; c = -(slown * slown) * (dnx * dnx); tref = TTN(iz, j); tdiv = 1.0; }
      } else {
        if (swk == 0) {
          swsol = 1;
          u = 2.0 * dnz;
          a = 1.0; b = 0.0; c = -(u * u) * (slown * slown);
          tref = 4.0 * TTN(k, ix) - TTN(k2, ix);
          tdiv = 3.0;
        } else if (NSTS(k, ix) == 0) {
          swsol = 1;
          a = 1.0; b = 0.0; c = -(slown * slown) * (dnz * dnz);
          tref = TTN(k, ix);
          tdiv = 1.0;
        }
      }
      if (swsol) {
        double trav = (tref + qsolve(a, b, c)) / tdiv;
        if (tsw1) travm = trav < travm ? trav : travm; else { travm = trav; tsw1 = 1; }
      }
    }
  }
  TTN(iz, ix) = travm;
}

/* bilinear: fm2d_ttime.f90:441-455 */
static double bilinear(const fm_t* F, double nv[3][3], double dsx, double dsz) {
  double biv = 0.0;
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) {
      double produ = (1.0 - fabs(((i - 1) * F->dnx - dsx) / F->dnx)) * (1.0 - fabs(((j - 1) * F->dnz - dsz) / F->dnz));
      biv = biv + nv[i][j] * produ;
    }
  return biv;
}

static void update_neighbour(fm_t* F, int iz, int ix) {
  if (NSTS(iz, ix) == -1) {
    if (F->fom == 0) fouds1(F, iz, ix); else fouds2(F, iz, ix);
    addtree(F, iz, ix);
  } else if (NSTS(iz, ix) > 0) {
    if (F->fom == 0) fouds1(F, iz, ix); else fouds2(F, iz, ix);
    updtree(F, iz, ix);
  }
}

/* travel: fm2d_ttime.f90:27-136 */
static void travel(fm_t* F, double scx, double scz, int urg) {
  int isx = (int)((scx - F->gox) / F->dnx) + 1;
  int isz = (int)((scz - F->goz) / F->dnz) + 1;
  if (isx < 1 || isx > F->nnx || isz < 1 || isz > F->nnz) { F->error = 1; return; } /* STOP in the Fortran */
  if (isx == F->nnx) isx--;
  if (isz == F->nnz) isz--;
  F->ntr = 0;
  if (urg == 2) {
    for (int i = 1; i <= F->nnx; ++i)
      for (int j = 1; j <= F->nnz; ++j)
        if (NSTS(j, i) > 0) addtree(F, j, i);
  } else {
    for (size_t q = 0; q < F->cells; ++q) F->nsts[q] = -1; /* nsts = -1: the whole array */
    double vss[3][3];
    for (int i = 1; i <= 2; ++i) for (int j = 1; j <= 2; ++j) vss[i][j] = VELN(isz - 1 + j, isx - 1 + i);
    double dsx = (scx - F->gox) - (isx - 1) * F->dnx;
    double dsz = (scz - F->goz) - (isz - 1) * F->dnz;
    double vsrc = bilinear(F, vss, dsx, dsz);
    for (int i = 1; i <= 2; ++i)
      for (int j = 1; j <= 2; ++j) {
        double ds = sqrt(sq(dsx - (i - 1) * F->dnx) + sq(dsz - (j - 1) * F->dnz));
        TTN(isz - 1 + j, isx - 1 + i) = 2.0 * ds / (vss[i][j] + vsrc);
        addtree(F, isz - 1 + j, isx - 1 + i);
      }
  }
  while (F->ntr > 0 && !F->error) {
    int ix = F->bpx[1], iz = F->bpz[1];
    if (urg == 1) {
      int swrg = 0;
      if (ix == 1 && F->vnl != 1) swrg = 1;
      if (ix == F->nnx && F->vnr != F->nnx) swrg = 1; /* note: nnx is the REFINED extent here, vnr a coarse index (as in the Fortran) */
      if (iz == 1 && F->vnt != 1) swrg = 1;
      if (iz == F->nnz && F->vnb != F->nnz) swrg = 1;
      if (swrg) { NSTS(iz, ix) = 0; break; }
    }
    NSTS(iz, ix) = 0;
    F->n_accept++;
    downtree(F);
    for (int i = ix - 1; i <= ix + 1; i += 2) if (i >= 1 && i <= F->nnx) update_neighbour(F, iz, i);
    for (int i = iz - 1; i <= iz + 1; i += 2) if (i >= 1 && i <= F->nnz) update_neighbour(F, i, ix);
  }
}

/* test entry: the march alone, on the caller's arrays (column major, leading dimension nnz) -- what
 * tests/test_oracle_fm2d_vs_reference.py compares with the mechanical translation of fm2d_ttime.f90 (oracle/f90toc.py).
 * heap_pxpz receives (px, pz) of the heap entries left when the march ends (urg = 1 stops early), ntr_out their number. */
int orc_fm2d_travel(int nnx, int nnz, double gox, double goz, double dnx, double dnz, int fom, const double* veln, double* ttn, int* nsts,
                    int urg, int vnl, int vnr, int vnt, int vnb, double scx, double scz, int* heap_pxpz, int* ntr_out) {
  fm_t F;
  memset(&F, 0, sizeof F);
  F.nnx = nnx; F.nnz = nnz; F.gox = gox; F.goz = goz; F.dnx = dnx; F.dnz = dnz; F.fom = fom;
  F.vnl = vnl; F.vnr = vnr; F.vnt = vnt; F.vnb = vnb;
  F.ld = nnz; F.cells = (size_t)nnx * nnz;
  F.veln = (double*)veln; F.ttn = ttn; F.nsts = nsts;
  F.maxbt = nnx * nnz + 1;
  F.bpx = (int*)calloc((size_t)F.maxbt + 2, sizeof(int));
  F.bpz = (int*)calloc((size_t)F.maxbt + 2, sizeof(int));
  travel(&F, scx, scz, urg);
  for (int i = 0; i < F.ntr; ++i) { heap_pxpz[2 * i] = F.bpx[i + 1]; heap_pxpz[2 * i + 1] = F.bpz[i + 1]; }
  *ntr_out = F.ntr;
  free(F.bpx); free(F.bpz);
  return F.error;
}

/* B-spline basis of gridder / bsplrefine */
static inline void bspl(double u, double w[5]) {
  double um = 1.0 - u;
  w[1] = ((um * um) * um) / 6.0;
  w[2] = (4.0 - 6.0 * (u * u) + 3.0 * ((u * u) * u)) / 6.0;
  w[3] = (1.0 + 3.0 * u + 3.0 * (u * u) - 3.0 * ((u * u) * u)) / 6.0;
  w[4] = ((u * u) * u) / 6.0;
}

/* gridder: fm2dray_cartesian.f90:490-590 (velv is already in place) */
static void gridder(fm_t* F) {
  const int gdx = F->gdx, gdz = F->gdz, nvx = F->nvx, nvz = F->nvz;
  double (*ui)[5] = malloc(sizeof(double[5]) * (size_t)(gdx + 2));
  double (*vi)[5] = malloc(sizeof(double[5]) * (size_t)(gdz + 2));
  for (int i = 1; i <= gdx + 1; ++i) { double u = gdx; u = (i - 1) / u; bspl(u, ui[i]); }
  for (int i = 1; i <= gdz + 1; ++i) { double u = gdz; u = (i - 1) / u; bspl(u, vi[i]); }
  for (int i = 1; i <= nvz - 1; ++i) {
    int conz = gdz; if (i == nvz - 1) conz = gdz + 1;
    for (int j = 1; j <= nvx - 1; ++j) {
      int conx = gdx; if (j == nvx - 1) conx = gdx + 1;
      for (int l = 1; l <= conz; ++l) {
        int stz = gdz * (i - 1) + l;
        for (int m = 1; m <= conx; ++m) {
          int stx = gdx * (j - 1) + m;
          double sumi = 0.0;
          for (int i1 = 1; i1 <= 4; ++i1) {
            double sumj = 0.0;
            for (int j1 = 1; j1 <= 4; ++j1) sumj = sumj + ui[m][j1] * VELV(i - 2 + i1, j - 2 + j1);
            sumi = sumi + vi[l][i1] * sumj;
          }
          VELN(stz, stx) = sumi;
        }
      }
    }
  }
  free(ui); free(vi);
}

/* bsplrefine: fm2dray_cartesian.f90:598-668; nnx/nnz are the REFINED extents when it runs */
static void bsplrefine(fm_t* F) {
  const int nrxr = F->gdx * F->sgdl, nrzr = F->gdz * F->sgdl;
  /* ui(j,i,:) depends on j only, vi(j,i,:) on i only */
  double (*ui)[5] = malloc(sizeof(double[5]) * (size_t)(nrxr + 2));
  double (*vi)[5] = malloc(sizeof(double[5]) * (size_t)(nrzr + 2));
  for (int j = 1; j <= nrxr + 1; ++j) { double u = nrxr; u = (j - 1) / u; bspl(u, ui[j]); }
  for (int i = 1; i <= nrzr + 1; ++i) { double v = nrzr; v = (i - 1) / v; bspl(v, vi[i]); }
  const int origx = (F->vnl - 1) * F->sgdl + 1, origz = (F->vnt - 1) * F->sgdl + 1;
  for (int i = 1; i <= F->nvz - 1; ++i) {
    int conz = nrzr; if (i == F->nvz - 1) conz = nrzr + 1;
    for (int j = 1; j <= F->nvx - 1; ++j) {
      int conx = nrxr; if (j == F->nvx - 1) conx = nrxr + 1;
      for (int k = 1; k <= conz; ++k) {
        int st1 = F->gdz * (i - 1) + (k - 1) / F->sgdl + 1;
        if (st1 < F->vnt || st1 > F->vnb) continue;
        st1 = nrzr * (i - 1) + k;
        for (int l = 1; l <= conx; ++l) {
          int st2 = F->gdx * (j - 1) + (l - 1) / F->sgdl + 1;
          if (st2 < F->vnl || st2 > F->vnr) continue;
          st2 = nrxr * (j - 1) + l;
          double sum[5];
          for (int i1 = 1; i1 <= 4; ++i1) {
            sum[i1] = 0.0;
            for (int j1 = 1; j1 <= 4; ++j1) sum[i1] = sum[i1] + ui[l][j1] * VELV(i - 2 + i1, j - 2 + j1);
            sum[i1] = vi[k][i1] * sum[i1];
          }
          int idm1 = st1 - origz + 1, idm2 = st2 - origx + 1;
          if (idm1 < 1 || idm1 > F->nnz) continue;
          if (idm2 < 1 || idm2 > F->nnx) continue;
          VELN(idm1, idm2) = sum[1] + sum[2] + sum[3] + sum[4];
        }
      }
    }
  }
  free(ui); free(vi);
}

/* srtimes: fm2dray_cartesian.f90:676-770 */
static void srtimes(fm_t* F, double scx, double scz, int csid, int nrc, const double* rcx, const double* rcz, const int* srs, double* ttime) {
  for (int i = 1; i <= nrc; ++i) {
    if (srs[(size_t)(csid - 1) * nrc + (i - 1)] == 0) continue;
    int irx = (int)floor((rcx[i - 1] - F->gox) / F->dnx) + 1;
    int irz = (int)floor((rcz[i - 1] - F->goz) / F->dnz) + 1;
    int sw = 0;
    if (irx < 1 || irx > F->nnx || irz < 1 || irz > F->nnz) { F->error = 3; return; }
    if (irx == F->nnx) irx--;
    if (irz == F->nnz) irz--;
    int isx = (int)floor((scx - F->gox) / F->dnx) + 1;
    int isz = (int)floor((scz - F->goz) / F->dnz) + 1;
    double dpl = F->dnx, rd1 = F->dnz;
    if (rd1 < dpl) dpl = rd1;
    double sred = sq(scx - rcx[i - 1]);
    sred = sred + sq(scz - rcz[i - 1]);
    sred = sqrt(sred);
    if (sred < dpl) sw = 1;
    if (isx == irx && isz == irz) sw = 1;
    double trr;
    if (sw) {
      double vss[3][3];
      for (int k = 1; k <= 2; ++k) for (int l = 1; l <= 2; ++l) vss[k][l] = VELN(isz - 1 + l, isx - 1 + k);
      double drx = (scx - F->gox) - (isx - 1) * F->dnx, drz = (scz - F->goz) - (isz - 1) * F->dnz;
      double vels = bilinear(F, vss, drx, drz);
      for (int k = 1; k <= 2; ++k) for (int l = 1; l <= 2; ++l) vss[k][l] = VELN(irz - 1 + l, irx - 1 + k);
      drx = (rcx[i - 1] - F->gox) - (irx - 1) * F->dnx; drz = (rcz[i - 1] - F->goz) - (irz - 1) * F->dnz;
      double velr = bilinear(F, vss, drx, drz);
      trr = 2.0 * sred / (vels + velr);
    } else {
      double drx = (rcx[i - 1] - F->gox) - (irx - 1) * F->dnx, drz = (rcz[i - 1] - F->goz) - (irz - 1) * F->dnz;
      trr = 0.0;
      for (int k = 1; k <= 2; ++k)
        for (int l = 1; l <= 2; ++l) {
          double produ = (1.0 - fabs(((l - 1) * F->dnz - drz) / F->dnz)) * (1.0 - fabs(((k - 1) * F->dnx - drx) / F->dnx));
          trr = trr + TTN(irz - 1 + l, irx - 1 + k) * produ;
        }
    }
    ttime[(size_t)(csid - 1) * nrc + (i - 1)] = trr;
  }
}


/* test entries for gridder / bsplrefine / srtimes alone (compared with the mechanical translation of the Fortran by
 * tests/test_oracle_fm2d_vs_reference.py).  velv: (nvx+2, nvz+2) C-order = the Fortran's velvin(nvz+2, nvx+2); veln_out
 * (nnx, nnz) C-order = veln(nnz, nnx). */
int orc_fm2d_gridder(int nvx, int nvz, int gdx, int gdz, const double* velv, double* veln_out) {
  fm_t F;
  memset(&F, 0, sizeof F);
  F.nvx = nvx; F.nvz = nvz; F.gdx = gdx; F.gdz = gdz;
  F.nnx = (nvx - 1) * gdx + 1; F.nnz = (nvz - 1) * gdz + 1; F.ld = F.nnz;
  F.velv = (double*)velv; F.veln = veln_out;
  gridder(&F);
  return 0;
}
int orc_fm2d_bsplrefine(int nvx, int nvz, int gdx, int gdz, int sgdl, int vnl, int vnr, int vnt, int vnb, const double* velv, int nnxr,
                        int nnzr, double* veln_out) {
  fm_t F;
  memset(&F, 0, sizeof F);
  F.nvx = nvx; F.nvz = nvz; F.gdx = gdx; F.gdz = gdz; F.sgdl = sgdl;
  F.vnl = vnl; F.vnr = vnr; F.vnt = vnt; F.vnb = vnb;
  F.nnx = nnxr; F.nnz = nnzr; F.ld = nnzr;
  F.velv = (double*)velv; F.veln = veln_out;
  bsplrefine(&F);
  return 0;
}
int orc_fm2d_srtimes(int nnx, int nnz, double gox, double goz, double dnx, double dnz, const double* veln, const double* ttn, double scx,
                     double scz, int nrc, const double* rcx, const double* rcz, const int* srs, double* ttime) {
  fm_t F;
  memset(&F, 0, sizeof F);
  F.nnx = nnx; F.nnz = nnz; F.ld = nnz; F.gox = gox; F.goz = goz; F.dnx = dnx; F.dnz = dnz;
  F.veln = (double*)veln; F.ttn = (double*)ttn;
  srtimes(&F, scx, scz, 1, nrc, rcx, rcz, srs, ttime);
  return F.error;
}

/* rpaths (fm2dray_cartesian.f90:773-1456) with cfd = 0 as the source hard-wires it (the Frechet block is dead code): every
 * receiver's ray is traced from the receiver towards the source in steps of dpl = half the smaller node spacing along
 * -grad T, the gradient taken by central (one-sided at the edges) differences at the lower-left node of the cell the point
 * is in -- of the refined source grid while all four nodes of that refined cell are alive (igref), of the coarse grid
 * otherwise.  Quirks kept: the coarse branch tests `ipzr == nnzr` (the REFINED indices) for its one-sided z difference
 * (:1092); receivers in the last cell column/row are refused here (ipx >= nnx) although srtimes accepts them.
 * Two places where the Fortran's behaviour is undefined are resolved and flagged instead: with asgr = 0 that `ipzr` is
 * never assigned (taken as "not equal"); a vanishing gradient (0/0) ends the ray as a crazy ray.
 * ray r of source csid goes to slot srsv(r) - 1: npts[slot], pts[slot*2*cap + 2*k + {0,1}], len[slot] = T_RAY%length. */
static void rpaths(fm_t* F, int asgr, double scx, double scz, int csid, int nrc, const double* rcx, const double* rcz, const int* srs,
                   const int* srsv, int nslots, int cap, int* ray_npts, double* ray_pts, double* ray_len, int* crazy) {
  const int maxrp = F->nnx * F->nnz;
  double* rgx = malloc(sizeof(double) * (size_t)(maxrp + 3));
  double* rgz = malloc(sizeof(double) * (size_t)(maxrp + 3));
  int isx, isz;
  if (asgr == 1) { isx = (int)floor((scx - F->goxr) / F->dnxr) + 1; isz = (int)floor((scz - F->gozr) / F->dnzr) + 1; }
  else { isx = (int)floor((scx - F->gox) / F->dnx) + 1; isz = (int)floor((scz - F->goz) / F->dnz) + 1; }
  double dpl = F->dnx, rd1 = F->dnz;
  if (rd1 < dpl) dpl = rd1;
  dpl = 0.5 * dpl;
  for (int i = 1; i <= nrc; ++i) {
    if (srs[(size_t)(csid - 1) * nrc + (i - 1)] == 0) continue;
    int ipx = (int)floor((rcx[i - 1] - F->gox) / F->dnx) + 1;
    int ipz = (int)floor((rcz[i - 1] - F->goz) / F->dnz) + 1;
    if (ipx < 1 || ipx >= F->nnx || ipz < 1 || ipz >= F->nnz) { F->error = 3; break; }
    int ipxr = 0, ipzr = 0, igref = 0, sw = 0, nrp = 1;
    rgx[1] = rcx[i - 1]; rgz[1] = rcz[i - 1];
    double sred = (scx - rgx[1]) * (scx - rgx[1]);
    sred = sred + (scz - rgz[1]) * (scz - rgz[1]);
    sred = sqrt(sred);
    if (sred < 2.0 * dpl) { rgx[2] = scx; rgz[2] = scz; nrp = 2; sw = 1; }
#define IGREF_AT(px, pz) do { \
      ipxr = (int)floor(((px) - F->goxr) / F->dnxr) + 1; ipzr = (int)floor(((pz) - F->gozr) / F->dnzr) + 1; igref = 1; \
      if (ipxr < 1 || ipxr >= F->nnxr) igref = 0; \
      if (ipzr < 1 || ipzr >= F->nnzr) igref = 0; \
      if (igref == 1) { \
        if (NSTSR(ipzr, ipxr) != 0 || NSTSR(ipzr + 1, ipxr) != 0) igref = 0; \
        if (NSTSR(ipzr, ipxr + 1) != 0 || NSTSR(ipzr + 1, ipxr + 1) != 0) igref = 0; \
      } } while (0)
    if (asgr == 1) IGREF_AT(rcx[i - 1], rcz[i - 1]); else igref = 0;
    if (sw == 0) {
      if (asgr == 1) { if (igref == 1 && ipxr == isx && ipzr == isz) { rgx[2] = scx; rgz[2] = scz; nrp = 2; sw = 1; } }
      else if (ipx == isx && ipz == isz) { rgx[2] = scx; rgz[2] = scz; nrp = 2; sw = 1; }
    }
    for (int j = 1; j <= maxrp; ++j) {
      if (sw == 1) break;
      double dtx, dtz;
      if (igref == 1) {
        if (ipxr == 1) { dtx = TTNR(ipzr, ipxr + 1) - TTNR(ipzr, ipxr); dtx = dtx / F->dnxr; }
        else if (ipxr == F->nnxr) { dtx = TTNR(ipzr, ipxr) - TTNR(ipzr, ipxr - 1); dtx = dtx / F->dnxr; }
        else { dtx = TTNR(ipzr, ipxr + 1) - TTNR(ipzr, ipxr - 1); dtx = dtx / (2.0 * F->dnxr); }
        if (ipzr == 1) { dtz = TTNR(ipzr + 1, ipxr) - TTNR(ipzr, ipxr); dtz = dtz / F->dnzr; }
        else if (ipzr == F->nnzr) { dtz = TTNR(ipzr, ipxr) - TTNR(ipzr - 1, ipxr); dtz = dtz / F->dnzr; }
        else { dtz = TTNR(ipzr + 1, ipxr) - TTNR(ipzr - 1, ipxr); dtz = dtz / (2.0 * F->dnzr); }
      } else {
        if (ipx == 1) { dtx = TTN(ipz, ipx + 1) - TTN(ipz, ipx); dtx = dtx / F->dnx; }
        else if (ipx == F->nnx) { dtx = TTN(ipz, ipx) - TTN(ipz, ipx - 1); dtx = dtx / F->dnx; }
        else { dtx = TTN(ipz, ipx + 1) - TTN(ipz, ipx - 1); dtx = dtx / (2.0 * F->dnx); }
        if (ipz == 1) { dtz = TTN(ipz + 1, ipx) - TTN(ipz, ipx); dtz = dtz / F->dnz; }
        else if (asgr == 1 && ipzr == F->nnzr) { dtz = TTN(ipz, ipx) - TTN(ipz - 1, ipx); dtz = dtz / F->dnz; } /* sic: ipzr, nnzr */
        else { dtz = TTN(ipz + 1, ipx) - TTN(ipz - 1, ipx); dtz = dtz / (2.0 * F->dnz); }
      }
      if (j + 2 > cap) { (*crazy)++; sw = 1; break; } /* output slots hold cap points: a ray still wandering after that many is
                                                         declared crazy here; the Fortran does so after nnx*nnz points */
      rd1 = sqrt(dtx * dtx + dtz * dtz);
      if (!(rd1 > 0.0)) { (*crazy)++; nrp = 1; sw = 1; break; } /* 0/0 in the Fortran */
      rgx[j + 1] = rgx[j] - dpl * dtx / rd1;
      rgz[j + 1] = rgz[j] - dpl * dtz / rd1;
      if (asgr == 1) IGREF_AT(rgx[j + 1], rgz[j + 1]); else igref = 0;
      ipx = (int)floor((rgx[j + 1] - F->gox) / F->dnx) + 1;
      ipz = (int)floor((rgz[j + 1] - F->goz) / F->dnz) + 1;
      sred = (scx - rgx[j + 1]) * (scx - rgx[j + 1]);
      sred = sred + (scz - rgz[j + 1]) * (scz - rgz[j + 1]);
      sred = sqrt(sred);
      sw = 0;
      if (sred < 2.0 * dpl) { rgx[j + 2] = scx; rgz[j + 2] = scz; nrp = j + 2; sw = 1; break; }
      if (asgr == 1) { if (igref == 1 && ipxr == isx && ipzr == isz) { rgx[j + 2] = scx; rgz[j + 2] = scz; nrp = j + 2; sw = 1; break; } }
      else if (ipx == isx && ipz == isz) { rgx[j + 2] = scx; rgz[j + 2] = scz; nrp = j + 2; sw = 1; break; }
      if (ipx < 1) { rgx[j + 1] = F->gox; ipx = 1; }
      if (ipx >= F->nnx) { rgx[j + 1] = F->gox + (F->nnx - 1) * F->dnx; ipx = F->nnx - 1; }
      if (ipz < 1) { rgz[j + 1] = F->goz; ipz = 1; }
      if (ipz >= F->nnz) { rgz[j + 1] = F->goz + (F->nnz - 1) * F->dnz; ipz = F->nnz - 1; }
      if (j == maxrp - 1 && sw == 0) { (*crazy)++; sw = 1; break; } /* nrp keeps its last value (1), as in the Fortran */
    }
#undef IGREF_AT
    const int slot = srsv[(size_t)(csid - 1) * nrc + (i - 1)] - 1;
    if (slot < 0 || slot >= nslots) { F->error = 5; break; }
    if (nrp > cap) { F->error = 4; break; }
    ray_npts[slot] = nrp;
    double len = 0;
    for (int k = 1; k <= nrp; ++k) {
      ray_pts[((size_t)slot * cap + (k - 1)) * 2 + 0] = rgx[k];
      ray_pts[((size_t)slot * cap + (k - 1)) * 2 + 1] = rgz[k];
      if (k >= 2) { double d = (rgx[k] - rgx[k - 1]) * (rgx[k] - rgx[k - 1]) + (rgz[k] - rgz[k - 1]) * (rgz[k] - rgz[k - 1]); len = len + sqrt(d); }
    }
    if (ray_len) ray_len[slot] = len;
  }
  free(rgx); free(rgz);
}

/*
 * modrays for ONE velocity map (one period), travel times only.
 *   scx,scz (nsrc), rcx,rcz (nrc): source / receiver coordinates (x = first grid axis, z = second);
 *   srs (nrc, nsrc) column-major: 1 where the pair carries data (raystat(:,1,period) reshaped);
 *   velv (nvz+2, nvx+2) column-major: like%vel(period,:,:) with its replicated edge;
 *   ttime (nrc, nsrc) column-major, entries without data untouched; field (optional, nnz*nnx per source) = ttn.
 * Returns 0, or 1 source outside the model, 2 narrow band larger than snb*nnx*nnz, 3 receiver outside the model.
 */
static int fm2d_core(int nsrc, const double* scx, const double* scz, int nrc, const double* rcx, const double* rcz, const int* srs, int nvx,
                     int nvz, double gox, double goz, double dvx, double dvz, const double* velv, int gdx, int gdz, int asgr, int sgdl, int sgs,
                     int fom, double snb, double* ttime, double* field, int64_t* counters, const int* srsv, int cap, int* ray_npts,
                     double* ray_pts, double* ray_len, int* crazy) {
  fm_t S, *F = &S;
  memset(F, 0, sizeof S);
  F->nvx = nvx; F->nvz = nvz; F->gdx = gdx; F->gdz = gdz; F->fom = fom; F->sgdl = sgdl;
  F->gox = gox; F->goz = goz; F->dvx = dvx; F->dvz = dvz;
  const int nnx0 = (nvx - 1) * gdx + 1, nnz0 = (nvz - 1) * gdz + 1;
  const int rmaxx = 2 * sgs * sgdl + 1, rmaxz = 2 * sgs * sgdl + 1; /* largest refined grid */
  const int mx = nnx0 > rmaxx ? nnx0 : rmaxx, mz = nnz0 > rmaxz ? nnz0 : rmaxz;
  F->ld = mz;
  const size_t cells = (size_t)mx * (size_t)mz;
  F->cells = cells;
  F->velv = (double*)velv;
  F->veln = calloc(cells, sizeof(double)); F->velnb = calloc(cells, sizeof(double));
  F->ttn = calloc(cells, sizeof(double)); F->ttnr = calloc(cells, sizeof(double));
  F->nsts = calloc(cells, sizeof(int)); F->nstsr = calloc(cells, sizeof(int));
  F->bpx = calloc(cells + 2, sizeof(int)); F->bpz = calloc(cells + 2, sizeof(int));
  F->nnx = nnx0; F->nnz = nnz0;
  F->dnx = dvx / gdx; F->dnz = dvz / gdz;
  gridder(F);
  F->maxbt = (int)lround(snb * nnx0 * nnz0);
  for (int i = 1; i <= nsrc && !F->error; ++i) {
    int any = 0;
    for (int r = 0; r < nrc; ++r) any += srs[(size_t)(i - 1) * nrc + r];
    if (any == 0 && i != 1) continue;
    const double x = scx[i - 1], z = scz[i - 1];
    int isx = (int)((x - F->gox) / F->dnx) + 1, isz = (int)((z - F->goz) / F->dnz) + 1;
    if (isx < 1 || isx > F->nnx || isz < 1 || isz > F->nnz) { F->error = 1; break; }
    if (isx == F->nnx) isx--;
    if (isz == F->nnz) isz--;
    F->vnl = isx - sgs; if (F->vnl < 1) F->vnl = 1;
    F->vnr = isx + sgs; if (F->vnr > F->nnx) F->vnr = F->nnx;
    F->vnt = isz - sgs; if (F->vnt < 1) F->vnt = 1;
    F->vnb = isz + sgs; if (F->vnb > F->nnz) F->vnb = F->nnz;
    if (asgr == 1) {
      /* back up the coarse grid */
      memcpy(F->velnb, F->veln, cells * sizeof(double));
      const int nnxb = F->nnx, nnzb = F->nnz;
      const double dnxb = F->dnx, dnzb = F->dnz, goxb = F->gox, gozb = F->goz;
      const int nrnx = (F->vnr - F->vnl) * sgdl + 1, nrnz = (F->vnb - F->vnt) * sgdl + 1;
      const double drnx = dvx / (double)(float)(gdx * sgdl), drnz = dvz / (double)(float)(gdz * sgdl); /* dvx/REAL(gdx*sgdl) */
      const double gorx = F->gox + F->dnx * (F->vnl - 1), gorz = F->goz + F->dnz * (F->vnt - 1);
      F->nnx = nrnx; F->nnz = nrnz; F->dnx = drnx; F->dnz = drnz; F->gox = gorx; F->goz = gorz;
      if (F->nnx > nnxb || F->nnz > nnzb) { /* the Fortran re-allocates here; the narrow-band array is re-sized too */
        const int idm1 = F->nnx > nnxb ? F->nnx : nnxb, idm2 = F->nnz > nnzb ? F->nnz : nnzb;
        F->maxbt = (int)lround(snb * idm1 * idm2);
      }
      bsplrefine(F);
      travel(F, x, z, 1);
      if (F->error) break;
      memcpy(F->ttnr, F->ttn, cells * sizeof(double));
      memcpy(F->nstsr, F->nsts, cells * sizeof(int));
      const int ogx = F->vnl, ogz = F->vnt;
      for (size_t q = 0; q < cells; ++q) F->nsts[q] = -1;
      for (int k = 1; k <= F->nnz; k += sgdl) {
        const int idm1 = ogz + (k - 1) / sgdl;
        for (int l = 1; l <= F->nnx; l += sgdl) {
          const int idm2 = ogx + (l - 1) / sgdl;
          NSTS(idm1, idm2) = NSTSR(k, l);
          if (NSTS(idm1, idm2) >= 0) TTN(idm1, idm2) = TTNR(k, l);
        }
      }
      F->nnxr = F->nnx; F->nnzr = F->nnz; F->goxr = F->gox; F->gozr = F->goz; F->dnxr = F->dnx; F->dnzr = F->dnz;
      F->nnx = nnxb; F->nnz = nnzb; F->dnx = dnxb; F->dnz = dnzb; F->gox = goxb; F->goz = gozb;
      for (int j = 1; j <= F->nnx; ++j) for (int k = 1; k <= F->nnz; ++k) VELN(k, j) = VELNB(k, j);
      for (int k = 1; k <= F->nnx; ++k)
        for (int l = 1; l <= F->nnz; ++l)
          if (NSTS(l, k) == 0) {
            if (l - 1 >= 1 && NSTS(l - 1, k) == -1) NSTS(l, k) = 1;
            if (l + 1 <= F->nnz && NSTS(l + 1, k) == -1) NSTS(l, k) = 1;
            if (k - 1 >= 1 && NSTS(l, k - 1) == -1) NSTS(l, k) = 1;
            if (k + 1 <= F->nnx && NSTS(l, k + 1) == -1) NSTS(l, k) = 1;
          }
      travel(F, x, z, 2);
    } else {
      travel(F, x, z, 0);
    }
    if (F->error) break;
    /* A march that dies at once (the refined-grid edge test of travel compares a REFINED extent with a COARSE index, so a source
     * cell in the model's last cell row/column counts as a refined edge) leaves ttn as the PREVIOUS source left it: the Fortran
     * shares one ttn array over the source loop, and so does this restatement. */
    if (g_unreached) { int u = 0; for (int j = 1; j <= F->nnx; ++j) for (int k = 1; k <= F->nnz; ++k) u += NSTS(k, j) != 0; g_unreached[i - 1] = u; }
    if (field) for (int j = 1; j <= F->nnx; ++j) for (int k = 1; k <= F->nnz; ++k) field[((size_t)(i - 1) * F->nnx + (j - 1)) * F->nnz + (k - 1)] = TTN(k, j);
    srtimes(F, x, z, i, nrc, rcx, rcz, srs, ttime);
    if (ray_npts && !F->error) rpaths(F, asgr, x, z, i, nrc, rcx, rcz, srs, srsv, nrc * nsrc, cap, ray_npts, ray_pts, ray_len, crazy); /* uar = 0 */
  }
  if (counters) { counters[0] += F->n_accept; counters[1] += F->n_update; }
  int err = F->error;
  free(F->veln); free(F->velnb); free(F->ttn); free(F->ttnr); free(F->nsts); free(F->nstsr); free(F->bpx); free(F->bpz);
  return err;
}

int orc_fm2d_times(int nsrc, const double* scx, const double* scz, int nrc, const double* rcx, const double* rcz, const int* srs, int nvx,
                   int nvz, double gox, double goz, double dvx, double dvz, const double* velv, int gdx, int gdz, int asgr, int sgdl, int sgs,
                   int fom, double snb, double* ttime, double* field, int64_t* counters) {
  return fm2d_core(nsrc, scx, scz, nrc, rcx, rcz, srs, nvx, nvz, gox, goz, dvx, dvz, velv, gdx, gdz, asgr, sgdl, sgs, fom, snb, ttime, field,
                   counters, 0, 0, 0, 0, 0, 0);
}
/* The same with uar = 0 (group-velocity data): ray geometry as well.  srsv (nrc, nsrc): raystat(:,2,period), the 1-based ray
 * slot of every pair; ray_npts[nrc*nsrc] (zeroed by the caller), ray_pts[nrc*nsrc][cap][2], ray_len[nrc*nsrc]; *crazy counts
 * the rays that ran out of points.  Error 4: a ray longer than cap; 5: a slot outside 1..nrc*nsrc. */
int orc_fm2d_rays(int nsrc, const double* scx, const double* scz, int nrc, const double* rcx, const double* rcz, const int* srs, const int* srsv,
                  int nvx, int nvz, double gox, double goz, double dvx, double dvz, const double* velv, int gdx, int gdz, int asgr, int sgdl,
                  int sgs, int fom, double snb, double* ttime, int cap, int* ray_npts, double* ray_pts, double* ray_len, int* crazy) {
  *crazy = 0;
  return fm2d_core(nsrc, scx, scz, nrc, rcx, rcz, srs, nvx, nvz, gox, goz, dvx, dvz, velv, gdx, gdz, asgr, sgdl, sgs, fom, snb, ttime, 0, 0,
                   srsv, cap, ray_npts, ray_pts, ray_len, crazy);
}
