/*
 * oracle/forward_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the glue between the gridded model and the 1-D solver:
 * vs2vp_3d / vp2rho_3d, check_model, convert_to_layer and the OpenMP column loop of
 * surf_likelihood / program modelling.  Used as the parity checker and as the timed
 * CPU baseline (bench.py cpu_baseline / --impl reference).  See the headers of
 * surfdisp96_ref.c (pinned on the mechanically translated surfdisp96.f) and kdtree2_ref.c (pinned on the reference's kdtree2.o).
 *
 * Reference lines restated (relative to /root/reference):
 *   src/utils.f90:102-112,125-134           vs2vp_3d, vp2rho_3d
 *   src/likelihood_surf.F90:631-646         check_model
 *   src/likelihood_surf.F90:523-629         convert_to_layer (EPS=1e-10, /scaling)
 *   src/forward_modelling.f90:72-175        convert_to_layer twin (EPS=1e-5, no scaling)
 *   src/likelihood_surf.F90:186-206         output presets + OpenMP column loop
 *   src/likelihood_surf_mmode.F90           same with np*nmodes outputs (surfmmodes)
 * Compile with -ffp-contract=off -fopenmp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../mctomo_b200/csrc/mct_math.h"

typedef struct {
  int nx, ny, nz;
  double xmin, ymin, zmin, dx, dy, dz;
  double waterDepth, scaling;
} orc_grid;

int orc_surfmodes(const double* thick, const double* vp, const double* vs, const double* rho, int n,
                  const double* freqs, int np, int modetype, int phaseGroup, int nmodes, double dc,
                  int math_mode, double* phase, double* group, int* ierr, int64_t* counters);

/* vs2vp_3d + vp2rho_3d: `vp = vs*POISSON` (POISSON = 1.730, a default-real literal stored in a
 * double) and `rho = 1.74*vp**0.25` (src/utils.f90:107-110,131-133). */
void orc_vs2vp_rho(const double* vs, double* vp, double* rho, int64_t n, int math_mode) {
  const double POISSON = (double)1.730f;
  const double c174 = (double)1.74f;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    vp[i] = vs[i] * POISSON;
    rho[i] = c174 * (math_mode ? mct_pow025(vp[i]) : pow(vp[i], 0.25));
  }
}

/* check_model: .true. (1) = model must be discarded */
int orc_check_model(const double* vs, int nx, int ny, int nz) {
  for (int64_t c = 0; c < (int64_t)nx * ny; ++c) {
    const double* col = vs + c * nz;
    for (int k = 1; k < nz; ++k)
      if (col[k] < col[0]) return 1;
  }
  return 0;
}

#define NMAXL 256 /* reference: NMAX = 100 (likelihood_surf.F90:30); anything above is UB there */

/*
 * convert_to_layer for one column.  layer_eps: 1e-10f (likelihood_surf.F90:37) or 1e-5f
 * (forward_modelling.f90:27) widened to double; water_thresh: EPS (likelihood_surf.F90:546)
 * or 0 (forward_modelling.f90:91).  Returns nlayers.
 */
int orc_convert_column(const double* vp, const double* vs, const double* rho, int nz, double dz,
                       double waterDepth, double scaling, double layer_eps, double water_thresh,
                       double* thick, double* alpha, double* beta, double* rho_k) {
  int nlayers;
  if (waterDepth > water_thresh) {
    nlayers = 1;
    alpha[0] = 1.5; /* waterVel */
    beta[0] = 0;
    rho_k[0] = 1; /* waterDensity */
    thick[0] = waterDepth;
  } else {
    nlayers = 0;
  }
  double last_vp = vp[0], last_vs = vs[0], last_rho = rho[0];
  int last_k = 1;
  for (int k = 2; k <= nz; ++k) {
    if (fabs(vs[k - 1] - last_vs) > layer_eps) {
      nlayers = nlayers + 1;
      if (nlayers >= NMAXL) return -1;
      alpha[nlayers - 1] = last_vp;
      beta[nlayers - 1] = last_vs;
      rho_k[nlayers - 1] = last_rho;
      thick[nlayers - 1] = (k - last_k) * dz;
      last_vp = vp[k - 1];
      last_vs = vs[k - 1];
      last_rho = rho[k - 1];
      last_k = k;
    }
  }
  nlayers = nlayers + 1;
  /* both branches of :576-603 (vbnd absent) end with the bottom cell's values and thick = 0 */
  alpha[nlayers - 1] = vp[nz - 1];
  beta[nlayers - 1] = vs[nz - 1];
  rho_k[nlayers - 1] = rho[nz - 1];
  thick[nlayers - 1] = 0;
  for (int i = 0; i < nlayers; ++i) thick[i] = thick[i] / scaling; /* :615 */
  if (waterDepth > 0) thick[0] = waterDepth;                       /* :617 */
  return nlayers;
}

/*
 * The dispersion block of surf_likelihood (likelihood_surf.F90:155-206) over the 1-based
 * inclusive column window ix0..ix1, iy0..iy1 (already clamped).
 *   pvel, gvel : (np*nm, iy0:iy1, ix0:ix1), nm = max(nmodes,1); preset to `preset`
 *                (100.0 in likelihood_surf.F90:188-189, 1000.0 in forward_modelling.f90:410-411)
 *   ierr       : (iy0:iy1, ix0:ix1); 0/1 from surfdisp96, 2 = column needs the GRT branch
 *   counters   : [0] dltar calls, [1] layer steps, summed over columns (may be NULL)
 * nmodes <= 0: surfmodes/surfdisp96; nmodes >= 1: surfmmodes/surfdisp_mmodes.
 * Returns the number of columns that would have taken the GRT branch.
 */
int orc_surf_dispersion(const double* vp, const double* vs, const double* rho, const orc_grid* g, int ix0,
                        int ix1, int iy0, int iy1, const double* freqs, int np, int raylov, int phaseGroup,
                        int nmodes, double dphase, double layer_eps, double water_thresh, double preset,
                        int math_mode, int nthreads, double* pvel, double* gvel, int* ierr,
                        int64_t* counters) {
  const int nm = nmodes <= 0 ? 1 : nmodes;
  const int wy = iy1 - iy0 + 1;
  int nunsup = 0;
  int64_t c0 = 0, c1 = 0;
  if (nthreads <= 0) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(+ : nunsup, c0, c1)
  for (int i = ix0; i <= ix1; ++i) {
    double thick[NMAXL], alpha[NMAXL], beta[NMAXL], rho_k[NMAXL];
    for (int j = iy0; j <= iy1; ++j) {
      size_t col = (size_t)(i - 1) * g->ny + (size_t)(j - 1);
      size_t oc = (size_t)(i - ix0) * wy + (size_t)(j - iy0);
      double* pv = pvel + oc * (size_t)np * nm;
      double* gv = gvel + oc * (size_t)np * nm;
      for (int q = 0; q < np * nm; ++q) { pv[q] = preset; gv[q] = preset; }
      ierr[oc] = 0;
      int n = orc_convert_column(vp + col * g->nz, vs + col * g->nz, rho + col * g->nz, g->nz, g->dz,
                                 g->waterDepth, g->scaling, layer_eps, water_thresh, thick, alpha, beta,
                                 rho_k);
      int64_t cnt[2] = {0, 0};
      int e = 0;
      int rc = (n < 0) ? 3 : orc_surfmodes(thick, alpha, beta, rho_k, n, freqs, np, raylov, phaseGroup,
                                           nmodes, dphase, math_mode, pv, gv, &e, cnt);
      if (rc == 0) ierr[oc] = e;
      else { ierr[oc] = rc; nunsup++; }
      c0 += cnt[0];
      c1 += cnt[1];
    }
  }
  if (counters) { counters[0] += c0; counters[1] += c1; }
  return nunsup;
}

/* Map assembly of surf_likelihood (likelihood_surf.F90:259-264): copy the window into the padded
 * (np, ny+2, nx+2) field and replicate the edges the window touches. */
void orc_assemble_vel(const double* pvel, int np, int nx, int ny, int ix0, int ix1, int iy0, int iy1,
                      double* vel) {
  const int wy = iy1 - iy0 + 1;
  const size_t sy = (size_t)np, sx = (size_t)np * (ny + 2);
  for (int i = ix0; i <= ix1; ++i)
    for (int j = iy0; j <= iy1; ++j)
      memcpy(vel + (size_t)i * sx + (size_t)j * sy, pvel + ((size_t)(i - ix0) * wy + (j - iy0)) * np,
             sizeof(double) * np);
  if (ix0 == 1) memcpy(vel, vel + sx, sizeof(double) * sx);
  if (ix1 == nx) memcpy(vel + (size_t)(nx + 1) * sx, vel + (size_t)nx * sx, sizeof(double) * sx);
  if (iy0 == 1)
    for (int i = 0; i < nx + 2; ++i) memcpy(vel + i * sx, vel + i * sx + sy, sizeof(double) * np);
  if (iy1 == ny)
    for (int i = 0; i < nx + 2; ++i)
      memcpy(vel + i * sx + (size_t)(ny + 1) * sy, vel + i * sx + (size_t)ny * sy, sizeof(double) * np);
}

/* GetVelocity (likelihood_surf.F90:496-521): bilinear interpolation of one period's (ny,nx) map at a point.
 * vel is the (np,ny,nx) map (element (i,iy,ix) at i + np*(iy + ny*ix), 0-based), ip the 0-based period. */
static double orc_get_velocity(const double* vel, int np, int ip, int nx, int ny, double xmin, double ymin, double dx,
                               double dy, double px, double py) {
  int ix = (int)floor((px - xmin) / dx) + 1;
  int iy = (int)floor((py - ymin) / dy) + 1;
  if (ix < 1) ix = 1;
  if (iy < 1) iy = 1;
  if (ix >= nx) ix = nx - 1;
  if (iy >= ny) iy = ny - 1;
  const double dsx = px - (xmin + (ix - 1) * dx);
  const double dsy = py - (ymin + (iy - 1) * dy);
  double qv = 0;
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) {
      const double weight = (1.0 - fabs((i - 1) * dx - dsx) / dx) * (1.0 - fabs((j - 1) * dy - dsy) / dy);
      qv = qv + weight * vel[(size_t)ip + (size_t)np * ((size_t)(iy + j - 2) + (size_t)ny * (size_t)(ix + i - 2))];
    }
  return qv;
}

/* CalGroupTime (likelihood_surf.F90:454-494): travel time of every ray through the map of its period, trapezoidal
 * in slowness-free form dist*2/(vhead+vtail), accumulated point by point.  Rays are packed: ray r of period ip owns
 * points off[ip*nrays + r] .. off[ip*nrays + r + 1]-1 of pts (x,y pairs); time is (nrays, np), ray index fastest. */
void orc_cal_group_time(const double* vel, int np, int nx, int ny, double xmin, double ymin, double dx, double dy,
                        const double* pts, const int64_t* off, int nrays, double* time) {
  for (int ip = 0; ip < np; ++ip)
    for (int r = 0; r < nrays; ++r) {
      const int64_t a = off[(size_t)ip * nrays + r], b = off[(size_t)ip * nrays + r + 1];
      double t = 0;
      if (b - a >= 2) {
        double vhead = orc_get_velocity(vel, np, ip, nx, ny, xmin, ymin, dx, dy, pts[2 * a], pts[2 * a + 1]);
        for (int64_t n = a + 1; n < b; ++n) {
          const double ex = pts[2 * n] - pts[2 * (n - 1)], ey = pts[2 * n + 1] - pts[2 * (n - 1) + 1];
          double dist = ex * ex + ey * ey;
          dist = sqrt(dist);
          const double vtail = orc_get_velocity(vel, np, ip, nx, ny, xmin, ymin, dx, dy, pts[2 * n], pts[2 * n + 1]);
          t = t + dist * 2 / (vhead + vtail);
          vhead = vtail;
        }
      }
      time[(size_t)ip * nrays + r] = t;
    }
}

/* The Gaussian misfit of surf_likelihood (likelihood_surf.F90:356-404).  Arrays in the Fortran layout:
 * time (nrr,np) = like%phaseTime(k,j,i) flattened over (k,j); ttime (nrr,3,np); raystat (nrr,2,np); srdist (nrr,np);
 * snoise0/1 (np).  out = {like, misfit, unweighted_misfit}; sigma (nrr,np) out.  Returns 6 where the reference raises
 * 'The noise level is 0!' (:387-390), else 0.  math_mode 1: log from mct_math.h (what the device uses), 0: libm. */
int orc_surf_misfit(const double* time, int nrr, int np, int sigdep, int nrays_total, const double* ttime, const int* raystat,
                    const double* snoise0, const double* snoise1, const double* srdist, int math_mode, double* out, double* sigma) {
  const double EPS = (double)1.0E-10f; /* real(ii10), parameter :: EPS = 1.0E-10 (:37) */
  const double PI2 = (double)6.283185f; /* :36 */
  int bad = 0;
  /* noise level, :357-373 */
  for (int i = 0; i < np; ++i)
    for (int r = 0; r < nrr; ++r) {
      const size_t t = (size_t)r + (size_t)nrr * i;
      if (sigdep != 0) {
        if (raystat[(size_t)r + (size_t)nrr * 2 * i] == 1) sigma[t] = snoise0[i] * srdist[t] + snoise1[i];
        else sigma[t] = 1.0;
      } else {
        sigma[t] = ttime[(size_t)r + (size_t)nrr * (1 + 3 * (size_t)i)];
      }
    }
  double like = 0, misfit = 0, unw = 0;
  for (int i = 0; i < np; ++i)
    for (int r = 0; r < nrr; ++r) { /* nrr runs over j (sources) then k (receivers), receiver fastest: :381-385 */
      const size_t t = (size_t)r + (size_t)nrr * i;
      if (raystat[(size_t)r + (size_t)nrr * 2 * i] == 1) {
        if (sigma[t] < EPS) bad = 1;
        const double d = time[t] - ttime[(size_t)r + (size_t)nrr * 3 * (size_t)i];
        like = like + (d * d) / (2 * (sigma[t] * sigma[t]));
        misfit = misfit + (d * d) / (sigma[t] * sigma[t]);
        unw = unw + (d * d);
      } else {
        sigma[t] = 1.0;
      }
    }
  double slog = 0;
  for (size_t t = 0; t < (size_t)nrr * np; ++t) slog = slog + (math_mode ? mct_log(sigma[t]) : log(sigma[t]));
  like = like + slog + (double)((float)nrays_total / 2.0f) * (math_mode ? mct_log(PI2) : log(PI2));
  out[0] = like; out[1] = misfit; out[2] = unw;
  return bad ? 6 : 0;
}

double orc_mct_log(double x) { return mct_log(x); }
