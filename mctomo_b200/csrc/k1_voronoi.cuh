// k1_voronoi.cuh -- stage 1 device code: exact nearest-nucleus assignment for every grid node,
// plus the small whole-grid kernels that sit between stage 1 and stage 2 (property maps,
// check_model, column -> layers, map assembly).
//
// The nearest-nucleus search must return what kdtree2 returns (reference src/kdtree2.f90:
// 1028-1069, 1369-1443, 1496-1599), including WHICH nucleus wins an exact tie -- that depends on
// the order in which kdtree2's traversal meets the candidates, not on the index.  So the device
// walks the very same tree (built on the host by kdtree_build.h with kdtree2's split rule) in the
// same order with the same comparisons; the recursion of the Fortran is replaced by an explicit
// stack of "far child still to be considered" entries.  Threads are mapped to consecutive z nodes
// of a column, so a warp's lanes follow nearly identical paths and the four output streams
// (sites_id, vp, vs, rho) are written fully coalesced.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "mct_math.h"

#ifndef MCT_MAX_LAYERS
#define MCT_MAX_LAYERS 200
#endif

struct KdNodeDev {
  double cut_val, cut_left, cut_right; // cut_val, cut_val_left, cut_val_right
  double lo[3], up[3];                 // node box
  int32_t cut_dim;                     // 0..2, -1 for a leaf
  int32_t left, right;                 // node ids, -1 = absent
  int32_t l, u;                        // 0-based inclusive range into the rearranged points
  int32_t pad;
};

// One entry per model of a batch (chains sharing a GPU): where its tree and nuclei start.
struct K1Model {
  long long node_off, pt_off;
  int32_t root, n;
};

struct K1Params {
  // batch form (k1_column_kernel, blockIdx.y = model): the four pointers below are the BASES of the packed
  // arrays and models[b] gives the offsets; outputs of model b start at b*model_stride.  NULL = single model.
  const K1Model* models;
  long long model_stride;
  const KdNodeDev* nodes;
  const double* rpts;    // rearranged points (3, n)
  const int32_t* ind;    // rearranged position -> original 1-based nucleus index
  const double* params;  // (3, n): vp, vs, rho per ORIGINAL nucleus
  int32_t root;
  int32_t n;             // number of nuclei
  // window (1-based inclusive) and grid geometry
  int32_t ix0, iy0, iz0, wx, wy, wz;
  double xmin, ymin, zmin, dx, dy, dz;
  // array geometry: element (k,j,i) lives at ((i-ia0)*ny_a + (j-ja0))*nz_a + (k-ka0)
  int32_t ia0, ja0, ka0, ny_a, nz_a;
  double* vp; double* vs; double* rho; int32_t* sites;
  int32_t use_pm; double pm_vp, pm_vs, pm_eps;
  int32_t* err; // set to 1 on traversal-stack overflow
};

#define K1_STACK 64

__device__ __forceinline__ double dis2_from_bnd_dev(double x, double amin, double amax) {
  if (x > amax) return (x - amax) * (x - amax);
  if (x < amin) return (amin - x) * (amin - x);
  return 0.0;
}

// kdtree2_n_nearest(nn=1) for one query; returns the ORIGINAL 1-based nucleus index.
__device__ __forceinline__ int kd_nearest_dev(const K1Params& P, double q0, double q1, double q2, int32_t* err) {
  const double q[3] = {q0, q1, q2};
  double ball = (double)FLT_MAX; // sr%ballsize = huge(1.0), kdtree2.f90:1038
  int best = 0;
  int32_t stack[K1_STACK];
  int sp = 0;
  int cur = P.root;
  for (;;) {
    // descend to a terminal node, remembering every internal node passed (search :1401-1413)
    for (;;) {
      const KdNodeDev* N = &P.nodes[cur];
      const int left = N->left, right = N->right;
      if (!(left >= 0 && right >= 0)) {
        // process_terminal_node (:1532-1593) with nn = 1
        const int l = N->l, u = N->u;
        for (int i = l; i <= u; ++i) {
          const double d0 = P.rpts[3 * i + 0] - q0;
          const double d1 = P.rpts[3 * i + 1] - q1;
          const double d2 = P.rpts[3 * i + 2] - q2;
          double sd = 0.0 + d0 * d0;
          if (sd > ball) continue;
          sd = sd + d1 * d1;
          if (sd > ball) continue;
          sd = sd + d2 * d2;
          if (sd > ball) continue;
          ball = sd; // first hit: pq_insert; later hits: pq_replace_max (accepts sd == ballsize)
          best = P.ind[i];
        }
        break;
      }
      if (sp >= K1_STACK) { *err = 1; return best; }
      stack[sp++] = cur;
      cur = (q[N->cut_dim] < N->cut_val) ? left : right;
    }
    // unwind: for each pending internal node decide whether its far child must be searched (:1416-1440)
    bool descend = false;
    while (sp > 0) {
      const int id = stack[--sp];
      const KdNodeDev* N = &P.nodes[id];
      const int cd = N->cut_dim;
      const double qval = q[cd];
      int farther;
      double dis;
      if (qval < N->cut_val) { farther = N->right; dis = (N->cut_right - qval) * (N->cut_right - qval); }
      else { farther = N->left; dis = (N->cut_left - qval) * (N->cut_left - qval); }
      if (dis <= ball) {
        bool pruned = false;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i != cd && !pruned) {
            dis = dis + dis2_from_bnd_dev(q[i], N->lo[i], N->up[i]);
            if (dis > ball) pruned = true;
          }
        }
        if (!pruned) { cur = farther; descend = true; break; }
      }
    }
    if (!descend) return best;
  }
}

// K1: one thread per grid node of the window, z fastest.
__global__ void __launch_bounds__(256) k1_voronoi_kernel(const __grid_constant__ K1Params P) {
  const long long total = (long long)P.wx * P.wy * P.wz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int kz = (int)(t % P.wz);
    const long long r = t / P.wz;
    const int jy = (int)(r % P.wy);
    const int ixx = (int)(r / P.wy);
    const int i = P.ix0 + ixx, j = P.iy0 + jy, k = P.iz0 + kz; // 1-based
    const size_t o = ((size_t)(i - P.ia0) * P.ny_a + (size_t)(j - P.ja0)) * P.nz_a + (size_t)(k - P.ka0);
    if (P.use_pm) { // mcmc_loc2.f90:2055-2056
      if (!(fabs(P.vs[o] - P.pm_vs) < P.pm_eps && fabs(P.vp[o] - P.pm_vp) < P.pm_eps)) continue;
    }
    const double qx = P.xmin + (double)(i - 1) * P.dx; // :2054
    const double qy = P.ymin + (double)(j - 1) * P.dy;
    const double qz = P.zmin + (double)(k - 1) * P.dz;
    const int idx = kd_nearest_dev(P, qx, qy, qz, P.err);
    P.sites[o] = idx;
    const double* pr = P.params + 3 * (size_t)(idx - 1);
    P.vp[o] = pr[0];
    P.vs[o] = pr[1];
    P.rho[o] = pr[2];
  }
}

// ---- nearest nucleus of ARBITRARY query points (not grid nodes) ------------------------------------------------------
// kdtree_locate (reference mcmc2d/mcmc.f90:1528-1551) and sites_locate (src/likelihood_body.F90:799-831): one thread
// per query walks kdtree2's traversal.  With d_sites != NULL the reference's shortcut is applied first: when the eight
// nodes around the point carry the same cell index, that index is the answer and the tree is not consulted.
__global__ void __launch_bounds__(128) k1_points_kernel(const __grid_constant__ K1Params P, const double* __restrict__ q, long long nq,
                                                        const int32_t* __restrict__ d_sites, int nx, int ny, int nz, double scaling,
                                                        int32_t* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  const double x = q[3 * t], y = q[3 * t + 1], z = q[3 * t + 2];
  if (d_sites) {
    // point2idx, src/likelihood_body.F90:1038-1054 (zmin and dz arrive scaled, as grid_setup stores them)
    int ix = (int)floor((x - P.xmin) / P.dx) + 1;
    int iy = (int)floor((y - P.ymin) / P.dy) + 1;
    int iz = (int)floor((z - P.zmin / scaling) / (P.dz / scaling)) + 1;
    if (ix < 1) ix = 1;
    if (iy < 1) iy = 1;
    if (iz < 1) iz = 1;
    if (ix >= nx) ix = nx - 1;
    if (iy >= ny) iy = ny - 1;
    if (iz >= nz) iz = nz - 1;
    const int32_t idx = d_sites[((size_t)(ix - 1) * ny + (size_t)(iy - 1)) * nz + (size_t)(iz - 1)];
    int num = 0;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) num += abs(d_sites[((size_t)(ix - 1 + i) * ny + (size_t)(iy - 1 + j)) * nz + (size_t)(iz - 1 + k)] - idx);
    if (num == 0) { out[t] = idx; return; }
  }
  out[t] = kd_nearest_dev(P, x, y, z, P.err);
}

#include "k1_column.cuh" // k1_column_kernel: culled brute force per column, tree walk only for (near-)ties (round-1 shape)
#include "k1_tile.cuh"   // k1_tile_kernel: per-(tile, z-segment) candidate lists, float64 walk, vector stores (mode 3)
#include "k1_box.cuh"    // k1_box_kernel: the same lists, warp per box, float32 walk with a rigorous error bound (production)

// vs2vp_3d + vp2rho_3d (src/utils.f90:107-110,131-133), elementwise over n values.
__global__ void __launch_bounds__(256) vs2vp_rho_kernel(const double* __restrict__ vs, double* __restrict__ vp,
                                                        double* __restrict__ rho, long long n) {
  const double POISSON = (double)1.730f;
  const double c174 = (double)1.74f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const double p = vs[t] * POISSON;
    vp[t] = p;
    rho[t] = c174 * mct_pow025(p);
  }
}

// Posterior accumulation of the post-processor (src/sample.f90:483-486): aveS += vs, stdS += vs**2, aveP += vp,
// stdP += vp**2, elementwise over the grid, after every kept sample's full-grid regrid.  Kept on the device so
// the thousands of regrids `program sample` issues never leave HBM.
__global__ void __launch_bounds__(256) accumulate_stats_kernel(const double* __restrict__ vs, const double* __restrict__ vp,
                                                               double* __restrict__ aveS, double* __restrict__ stdS,
                                                               double* __restrict__ aveP, double* __restrict__ stdP, long long n) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const double s = vs[t], p = vp[t];
    aveS[t] = aveS[t] + s;
    stdS[t] = stdS[t] + s * s;
    aveP[t] = aveP[t] + p;
    stdP[t] = stdP[t] + p * p;
  }
}

// check_model (src/likelihood_surf.F90:631-646) over the columns ix0..ix0+wx-1, iy0..iy0+wy-1 (1-based; the reference
// scans the whole grid = 1..nx, 1..ny): flag |= any(vs(2:,j,i) < vs(1,j,i)).  One warp per column, lanes strided over z
// (coalesced).  Batched form: model b = c / (wx*wy) reads its columns at vs + b*model_stride and raises flag[2*b].
__global__ void __launch_bounds__(256) check_model_kernel(const double* __restrict__ vs, int ix0, int iy0, int wx, int wy, int ny,
                                                          int nmodels, long long model_stride, int nz, int32_t* flag) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long cols_per_model = (long long)wx * wy;
  const long long ncols = cols_per_model * nmodels;
  for (long long c = warp; c < ncols; c += nwarps) {
    const long long b = c / cols_per_model, cm = c - b * cols_per_model;
    const long long i = ix0 + cm / wy, j = iy0 + cm % wy;
    const double* col = vs + b * model_stride + ((i - 1) * ny + (j - 1)) * nz;
    const double v1 = col[0];
    bool bad = false;
    for (int k = 1 + lane; k < nz; k += 32) bad |= (col[k] < v1);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&flag[2 * b], 1);
  }
}

struct LayParams {
  const double* vp; const double* vs; const double* rho; // whole-grid (nz,ny,nx) arrays
  int32_t ny, nz;
  int32_t ix0, iy0, wx, wy; // 1-based window origin and extent (per model)
  int32_t nmodels;          // models stacked along the slowest axis: model b starts at b*model_stride
  long long model_stride;   // nx*ny*nz
  double dz, waterDepth, scaling, layer_eps, water_thresh;
  int32_t modetype; // 1 Rayleigh, 0 Love (for the nlvls1 predicate)
  float4* lay; int32_t* nlay; int32_t* status; int32_t stride;
  int32_t* flags; // per model b: flags[2*b+1] = max status code
  int32_t grt_on; // low-velocity columns will be solved by the generalized R/T kernel: status 2 is not a condition to report
};

// convert_to_layer (src/likelihood_surf.F90:523-629 / forward_modelling.f90:72-175) + the real(.,4)
// narrowing of surfmodes.f90:81-83 + setup_grt's nlvls1 predicate (surfmodes.f90:320-410).
// One thread per column of the window; column index c = (i-ix0)*wy + (j-iy0).
__global__ void __launch_bounds__(128) layerize_kernel(const __grid_constant__ LayParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int cpm = P.wx * P.wy;
  if (c >= cpm * P.nmodels) return;
  const int b = c / cpm, cm = c - b * cpm;
  const int i = P.ix0 + cm / P.wy, j = P.iy0 + cm % P.wy;
  const size_t base = (size_t)b * (size_t)P.model_stride + ((size_t)(i - 1) * P.ny + (size_t)(j - 1)) * P.nz;
  const double* vp = P.vp + base;
  const double* vs = P.vs + base;
  const double* rho = P.rho + base;
  float4* lay = P.lay + c;
  const double grt_eps = (double)1e-6f; // surfmodes.f90:18
  int nl = 0;
  int st = 0;
  // running facts for the nlvls1 predicate
  int ifs = 0, nlvl1 = 0, nlvls_w = 0;
  double vs1 = 0.0, vss1 = 0.0, L1vp = 0.0; // vs1/vss1 as in setup_grt; L1vp = vp(1)
  double pvp2 = 0.0, pvp1 = 0.0, pvs1 = 0.0; // vp(i-1), vp(i), vs(i) for the layer i = nl-1 awaiting its lower neighbour

  auto emit = [&](double th, double a, double b, double r, bool last) {
    nl = nl + 1;
    if (nl > MCT_MAX_LAYERS) { st = 3; return; }
    // thick/scaling (:615), first-layer water override (:617), then real(.,4)
    double thk = th / P.scaling;
    if (nl == 1 && P.waterDepth > 0) thk = P.waterDepth;
    lay[(size_t)(nl - 1) * P.stride] = make_float4((float)thk, (float)a, (float)b, (float)r);
    // fluid bookkeeping (surfmodes.f90:335-348)
    if (!(fabs(b) > grt_eps)) {
      if (nl > 1) { if (st < 4) st = 4; }
      else ifs = 1;
    }
    if (nl == 1) { L1vp = a; if (ifs == 0) vs1 = b; }
    if (nl == 1 + ifs) {
      if (P.modetype == 1) { if (ifs > 0) { vs1 = L1vp; vss1 = b; } }
      else vs1 = b;
    }
    // evaluate the predicate for layer ii = nl-1 now that its lower neighbour (this layer) is known
    if (nl >= 3) {
      const int ii = nl - 1;
      if (ii > ifs && pvs1 < vss1) nlvls_w++;
      if (pvp1 < a && pvp1 < pvp2) {
        if (ifs == 0) { if (pvs1 < vs1) nlvl1++; }
        else if (P.modetype == 1) { if (pvp1 < vs1) nlvl1++; }
        else { if (pvs1 > 0.) { if (pvs1 < vs1) nlvl1++; } }
      }
    }
    pvp2 = pvp1; pvp1 = a; pvs1 = b;
    (void)last;
  };

  if (P.waterDepth > P.water_thresh) emit(P.waterDepth, 1.5, 0.0, 1.0, false); // waterVel, waterDensity (:31-32,546-551)
  double last_vp = vp[0], last_vs = vs[0], last_rho = rho[0];
  int last_k = 1;
  for (int k = 2; k <= P.nz; ++k) {
    const double v = vs[k - 1];
    if (fabs(v - last_vs) > P.layer_eps) {
      emit((double)(k - last_k) * P.dz, last_vp, last_vs, last_rho, false);
      last_vp = vp[k - 1];
      last_vs = v;
      last_rho = rho[k - 1];
      last_k = k;
    }
  }
  emit(0.0, vp[P.nz - 1], vs[P.nz - 1], rho[P.nz - 1], true); // half-space = bottom cell (:575-603)
  int nlvls1 = (ifs == 0 || P.modetype == 0) ? nlvl1 : nlvls_w;
  if (st == 0 && nlvls1 != 0) st = 2;
  P.nlay[c] = nl > MCT_MAX_LAYERS ? MCT_MAX_LAYERS : nl;
  P.status[c] = st;
  if (st != 0 && !(st == 2 && P.grt_on) && P.flags) atomicMax(&P.flags[2 * b + 1], st);
}

// Layering of pre-layered columns (mct_surfmodes_batch): narrow to float4, evaluate the same predicate.
struct PreLayParams {
  const double* thick; const double* vp; const double* vs; const double* rho; const long long* offsets;
  int32_t ncol, modetype, stride, grt_on;
  float4* lay; int32_t* nlay; int32_t* status; int32_t* flags;
};
__global__ void __launch_bounds__(128) prelayered_kernel(const __grid_constant__ PreLayParams P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.ncol) return;
  const long long o = P.offsets[c];
  const int n = (int)(P.offsets[c + 1] - o);
  const double grt_eps = (double)1e-6f;
  int st = 0;
  if (n < 1) st = 1;
  if (n > MCT_MAX_LAYERS) st = 3;
  int ifs = 0;
  if (st == 0) {
    for (int i = 0; i < n; ++i) {
      P.lay[(size_t)i * P.stride + c] = make_float4((float)P.thick[o + i], (float)P.vp[o + i], (float)P.vs[o + i], (float)P.rho[o + i]);
      if (!(fabs(P.vs[o + i]) > grt_eps)) { if (i > 0) st = 4; else ifs = 1; }
    }
  }
  if (st == 0) {
    double vs1, vss1 = 0.0;
    if (P.modetype == 1) { if (ifs > 0) { vs1 = P.vp[o]; vss1 = P.vs[o + ifs]; } else vs1 = P.vs[o]; }
    else vs1 = P.vs[o + ifs];
    int nlvls_w = 0, nlvl1 = 0;
    for (int i = 2; i <= n - 1; ++i) {
      const double vpi = P.vp[o + i - 1], vsi = P.vs[o + i - 1];
      if (i > ifs && vsi < vss1) nlvls_w++;
      if (vpi < P.vp[o + i] && vpi < P.vp[o + i - 2]) {
        if (ifs == 0) { if (vsi < vs1) nlvl1++; }
        else if (P.modetype == 1) { if (vpi < vs1) nlvl1++; }
        else { if (vsi > 0.) { if (vsi < vs1) nlvl1++; } }
      }
    }
    const int nlvls1 = (ifs == 0 || P.modetype == 0) ? nlvl1 : nlvls_w;
    if (nlvls1 != 0) st = 2;
  }
  P.nlay[c] = n < 1 ? 1 : (n > MCT_MAX_LAYERS ? MCT_MAX_LAYERS : n);
  P.status[c] = st;
  if (st != 0 && !(st == 2 && P.grt_on) && P.flags) atomicMax(&P.flags[1], st);
}

// like%vel(:, iy0+1:iy1+1, ix0+1:ix1+1) = pvel, then edge replication (likelihood_surf.F90:259-264).
// Pass 1 scatters the window; pass 2 (separate launch) replicates the touched edges.
__global__ void assemble_scatter_kernel(const double* __restrict__ pvel, int np, int ny, int ix0, int iy0, int wx, int wy,
                                        double* __restrict__ vel) {
  const long long total = (long long)wx * wy * np;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % np);
    const long long c = t / np;
    const int j = iy0 + (int)(c % wy), i = ix0 + (int)(c / wy);
    vel[((size_t)i * (ny + 2) + (size_t)j) * np + p] = pvel[t];
  }
}
__global__ void assemble_edges_kernel(double* vel, int np, int nx, int ny, int ex0, int ex1, int ey0, int ey1, int phase) {
  // phase 0: x edges (whole (np,ny+2) slabs); phase 1: y edges over all nx+2 slabs -- two launches,
  // in the Fortran's order, because the y pass re-reads the corners the x pass wrote.
  const long long sx = (long long)np * (ny + 2);
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stp = (long long)gridDim.x * blockDim.x;
  if (phase == 0) {
    for (long long t = t0; t < sx; t += stp) {
      if (ex0) vel[t] = vel[sx + t];
      if (ex1) vel[(long long)(nx + 1) * sx + t] = vel[(long long)nx * sx + t];
    }
  } else {
    const long long total = (long long)(nx + 2) * np;
    for (long long t = t0; t < total; t += stp) {
      const int p = (int)(t % np);
      const long long i = t / np;
      if (ey0) vel[i * sx + p] = vel[i * sx + np + p];
      if (ey1) vel[i * sx + (long long)(ny + 1) * np + p] = vel[i * sx + (long long)ny * np + p];
    }
  }
}
