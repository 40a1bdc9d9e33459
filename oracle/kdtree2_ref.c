/*
 * oracle/kdtree2_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the parts of M. Kennel's kdtree2 that MCTomo's
 * Voronoi->grid conversion uses (build + 1-nearest-neighbour search), and of
 * kdtree_to_grid itself, written as the parity checker for the CUDA nearest-nucleus
 * kernel.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * PARITY STATUS: pinned.  tests/test_oracle_kdtree.py checks this restatement
 * (index AND squared distance, bit for bit, including engineered ties) against the
 * reference's own pre-built object utils/libutils.a:kdtree2.o, run through
 * oracle/_ref (oracle/build_ref.sh), and against fixtures generated from that object
 * (tests/golden/kdtree2_*.npz, script tools/make_golden_kdtree.py).
 *
 * Reference lines restated (relative to /root/reference):
 *   src/kdtree2.f90:609-689   kdtree2_create (rearrange=.true., sort=.false.)
 *   src/kdtree2.f90:691-840   build_tree / build_tree_for_range
 *   src/kdtree2.f90:842-901   select_on_coordinate_value
 *   src/kdtree2.f90:936-983   spread_in_coordinate
 *   src/kdtree2.f90:1028-1069 kdtree2_n_nearest (nn = 1)
 *   src/kdtree2.f90:1369-1443 search
 *   src/kdtree2.f90:1446-1462 dis2_from_bnd
 *   src/kdtree2.f90:1496-1599 process_terminal_node (+ pq_insert / pq_replace_max with nn=1)
 *   src/mcmc_loc2.f90:2002-2082 kdtree_to_grid
 * Compile with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BUCKET 12 /* kdtree2.f90:516 */
#define DIM 3

typedef struct {
  int cut_dim;                      /* 1-based as in the Fortran, 0 for a leaf */
  double cut_val, cut_val_left, cut_val_right;
  int l, u;                         /* 1-based inclusive range into ind */
  int left, right;                  /* node ids, -1 = not associated */
  double lo[DIM], up[DIM];          /* box */
} knode;

typedef struct {
  int n;
  const double* data; /* (3,n) column-major: the_data */
  int* ind;           /* 1-based permutation, ind[0] unused */
  double* rdata;      /* rearranged_data (3,n) */
  knode* nodes;
  int nnodes, cap;
  int root;
  int fail; /* set when the reference would recurse forever (> bucket identical points) */
} ktree;

#define V(T, c, i) ((T)->data[(size_t)((i)-1) * DIM + ((c)-1)]) /* the_data(c,i), 1-based */

/* spread_in_coordinate: kdtree2.f90:936-983 */
static void spread(const ktree* T, int c, int l, int u, double* lo, double* up) {
  double smin = V(T, c, T->ind[l]);
  double smax = smin;
  int i;
  for (i = l + 2; i <= u; i += 2) {
    double lmin = V(T, c, T->ind[i - 1]);
    double lmax = V(T, c, T->ind[i]);
    if (lmin > lmax) { double t = lmin; lmin = lmax; lmax = t; }
    if (smin > lmin) smin = lmin;
    if (smax < lmax) smax = lmax;
  }
  if (i == u + 1) {
    double last = V(T, c, T->ind[u]);
    if (smin > last) smin = last;
    if (smax < last) smax = last;
  }
  *lo = smin;
  *up = smax;
}

/* select_on_coordinate_value: kdtree2.f90:842-901 */
static int select_value(ktree* T, int c, double alpha, int li, int ui) {
  int lb = li, rb = ui;
  while (lb < rb) {
    if (V(T, c, T->ind[lb]) <= alpha) {
      lb = lb + 1;
    } else {
      int tmp = T->ind[lb]; T->ind[lb] = T->ind[rb]; T->ind[rb] = tmp;
      rb = rb - 1;
    }
  }
  return (V(T, c, T->ind[lb]) <= alpha) ? lb : lb - 1;
}

static int new_node(ktree* T) {
  if (T->nnodes == T->cap) {
    T->cap = T->cap ? 2 * T->cap : 64;
    T->nodes = (knode*)realloc(T->nodes, sizeof(knode) * (size_t)T->cap);
  }
  memset(&T->nodes[T->nnodes], 0, sizeof(knode));
  T->nodes[T->nnodes].left = T->nodes[T->nnodes].right = -1;
  return T->nnodes++;
}

/* build_tree_for_range: kdtree2.f90:704-840.  parent = -1 for the root. */
static int build_range(ktree* T, int l, int u, int parent, int chain) {
  if (u < l) return -1;
  int id = new_node(T);
  if ((u - l) <= BUCKET) {
    for (int i = 1; i <= DIM; ++i) spread(T, i, l, u, &T->nodes[id].lo[i - 1], &T->nodes[id].up[i - 1]);
    T->nodes[id].cut_dim = 0;
    T->nodes[id].cut_val = 0.0;
    T->nodes[id].l = l;
    T->nodes[id].u = u;
    return id;
  }
  for (int i = 1; i <= DIM; ++i) {
    int recompute = 1;
    if (parent >= 0 && i != T->nodes[parent].cut_dim) recompute = 0;
    if (recompute) {
      spread(T, i, l, u, &T->nodes[id].lo[i - 1], &T->nodes[id].up[i - 1]);
    } else {
      T->nodes[id].lo[i - 1] = T->nodes[parent].lo[i - 1];
      T->nodes[id].up[i - 1] = T->nodes[parent].up[i - 1];
    }
  }
  /* c = maxloc(upper-lower): first maximum */
  int c = 1;
  double best = T->nodes[id].up[0] - T->nodes[id].lo[0];
  for (int i = 2; i <= DIM; ++i) {
    double e = T->nodes[id].up[i - 1] - T->nodes[id].lo[i - 1];
    if (e > best) { best = e; c = i; }
  }
  /* average = sum(the_data(c,ind(l:u))) / real(u-l+1,kdkind) */
  double sum = 0.0;
  for (int i = l; i <= u; ++i) sum += V(T, c, T->ind[i]);
  double average = sum / (double)(u - l + 1);
  T->nodes[id].cut_val = average;
  int m = select_value(T, c, average, l, u);
  /* An empty side is legal (kdtree2.f90:818-826: the node keeps its one child and the child's box; the search
   * then treats it as terminal, :1388).  Only points coincident in EVERY dimension make the Fortran recurse
   * for ever: a chain of 32 one-child levels over the same range is reported instead of overflowing the stack. */
  int one_child = (m >= u || m < l);
  if (one_child && chain >= 32) {
    T->fail = 1;
    T->nodes[id].l = l;
    T->nodes[id].u = u;
    return id;
  }
  T->nodes[id].cut_dim = c;
  T->nodes[id].l = l;
  T->nodes[id].u = u;
  int left = build_range(T, l, m, id, one_child ? chain + 1 : 0);
  int right = T->fail ? -1 : build_range(T, m + 1, u, id, one_child ? chain + 1 : 0);
  if (T->fail) return id;
  knode* N = &T->nodes[id]; /* (re-fetch: realloc may have moved the array) */
  N->left = left;
  N->right = right;
  if (right < 0) {
    memcpy(N->lo, T->nodes[left].lo, sizeof N->lo);
    memcpy(N->up, T->nodes[left].up, sizeof N->up);
    N->cut_val_left = T->nodes[left].up[c - 1];
    N->cut_val = N->cut_val_left;
  } else if (left < 0) {
    memcpy(N->lo, T->nodes[right].lo, sizeof N->lo);
    memcpy(N->up, T->nodes[right].up, sizeof N->up);
    N->cut_val_right = T->nodes[right].lo[c - 1];
    N->cut_val = N->cut_val_right;
  } else {
    N->cut_val_right = T->nodes[right].lo[c - 1];
    N->cut_val_left = T->nodes[left].up[c - 1];
    N->cut_val = (N->cut_val_left + N->cut_val_right) / 2;
    for (int i = 0; i < DIM; ++i) {
      N->up[i] = fmax(T->nodes[left].up[i], T->nodes[right].up[i]);
      N->lo[i] = fmin(T->nodes[left].lo[i], T->nodes[right].lo[i]);
    }
  }
  return id;
}

void* orc_kdtree2_create(const double* points, int n) {
  ktree* T = (ktree*)calloc(1, sizeof(ktree));
  T->n = n;
  T->data = points;
  T->ind = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  for (int j = 1; j <= n; ++j) T->ind[j] = j;
  T->root = build_range(T, 1, n, -1, 0);
  if (T->fail) { free(T->ind); free(T->nodes); free(T); return NULL; }
  T->rdata = (double*)malloc(sizeof(double) * DIM * (size_t)n);
  for (int i = 1; i <= n; ++i)
    for (int k = 0; k < DIM; ++k) T->rdata[(size_t)(i - 1) * DIM + k] = points[(size_t)(T->ind[i] - 1) * DIM + k];
  return T;
}

void orc_kdtree2_destroy(void* tp) {
  ktree* T = (ktree*)tp;
  free(T->ind); free(T->rdata); free(T->nodes); free(T);
}

typedef struct { const double* qv; double ballsize; int nfound; double dis; int idx; } ksearch;

static inline double dis2_from_bnd(double x, double amin, double amax) {
  if (x > amax) return (x - amax) * (x - amax);
  if (x < amin) return (amin - x) * (amin - x);
  return 0.0;
}

/* process_terminal_node with nn = 1: kdtree2.f90:1496-1599 */
static void terminal(const ktree* T, const knode* N, ksearch* sr) {
  double ballsize = sr->ballsize;
  for (int i = N->l; i <= N->u; ++i) {
    double sd = 0.0;
    int skip = 0;
    for (int k = 0; k < DIM; ++k) {
      double d = T->rdata[(size_t)(i - 1) * DIM + k] - sr->qv[k];
      sd = sd + d * d;
      if (sd > ballsize) { skip = 1; break; }
    }
    if (skip) continue;
    /* nfound < nn: pq_insert; else pq_replace_max -- with nn=1 both store (sd, ind(i)) */
    sr->nfound = 1;
    sr->dis = sd;
    sr->idx = T->ind[i];
    ballsize = sd;
  }
  sr->ballsize = ballsize;
}

/* search: kdtree2.f90:1369-1443 */
static void search(const ktree* T, int id, ksearch* sr) {
  const knode* N = &T->nodes[id];
  if (!(N->left >= 0 && N->right >= 0)) {
    terminal(T, N, sr);
    return;
  }
  int cut_dim = N->cut_dim;
  double qval = sr->qv[cut_dim - 1];
  int ncloser, nfarther;
  double dis;
  if (qval < N->cut_val) {
    ncloser = N->left; nfarther = N->right;
    dis = (N->cut_val_right - qval) * (N->cut_val_right - qval);
  } else {
    ncloser = N->right; nfarther = N->left;
    dis = (N->cut_val_left - qval) * (N->cut_val_left - qval);
  }
  if (ncloser >= 0) search(T, ncloser, sr);
  if (nfarther >= 0) {
    double ballsize = sr->ballsize;
    if (dis <= ballsize) {
      for (int i = 1; i <= DIM; ++i) {
        if (i != cut_dim) {
          dis = dis + dis2_from_bnd(sr->qv[i - 1], N->lo[i - 1], N->up[i - 1]);
          if (dis > ballsize) return;
        }
      }
      search(T, nfarther, sr);
    }
  }
}

/* kdtree2_n_nearest with nn=1: returns 1-based index, *dis = squared distance */
int orc_kdtree2_nearest(const void* tp, const double* qv, double* dis) {
  const ktree* T = (const ktree*)tp;
  ksearch sr;
  sr.qv = qv;
  sr.ballsize = (double)FLT_MAX; /* huge(1.0): single-precision huge, kdtree2.f90:1038 */
  sr.nfound = 0;
  sr.dis = 0.0;
  sr.idx = 0;
  search(T, T->root, &sr);
  if (dis) *dis = sr.dis;
  return sr.idx;
}

void orc_kdtree2_nearest_batch(const void* tp, const double* q, int64_t nq, int* idx, double* dis) {
  for (int64_t i = 0; i < nq; ++i) idx[i] = orc_kdtree2_nearest(tp, q + 3 * i, dis ? dis + i : NULL);
}

/* grid description mirroring T_GRID (src/settings.f90:20-28) */
typedef struct {
  int nx, ny, nz;
  double xmin, ymin, zmin, dx, dy, dz;
  double waterDepth, scaling;
} orc_grid;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* index window of kdtree_to_grid: src/mcmc_loc2.f90:2034-2045 */
void orc_box_window(const orc_grid* g, const double box[6], int w[6]) {
  w[0] = (int)floor((box[0] - g->xmin) / g->dx) + 1;
  w[1] = (int)floor((box[3] - g->xmin) / g->dx) + 1;
  w[2] = (int)floor((box[1] - g->ymin) / g->dy) + 1;
  w[3] = (int)floor((box[4] - g->ymin) / g->dy) + 1;
  w[4] = (int)floor((box[2] - g->zmin) / g->dz) + 1;
  w[5] = (int)floor((box[5] - g->zmin) / g->dz) + 1;
  if (w[0] < 1) w[0] = 1;
  if (w[2] < 1) w[2] = 1;
  if (w[4] < 1) w[4] = 1;
  if (w[1] > g->nx) w[1] = g->nx;
  if (w[3] > g->ny) w[3] = g->ny;
  if (w[5] > g->nz) w[5] = g->nz;
  (void)clampi;
}

/*
 * kdtree_to_grid: src/mcmc_loc2.f90:2002-2082.
 * points (3,ncells), params (3,ncells) = (vp,vs,rho); box = {x0,y0,z0,x1,y1,z1};
 * pm = NULL or {vp,vs,rho}; arrays (nz,ny,nx) column-major, updated in place inside
 * the window.  OpenMP over x like the reference (:2049).
 */
int orc_kdtree_to_grid(const double* points, const double* params, int ncells, const orc_grid* g,
                        const double box[6], const double* pm, double* vp, double* vs, double* rho,
                        int* sites_id) {
  const double eps = (double)1e-8f; /* real(kind=ii10), parameter :: eps = 1e-8, mcmc_loc2.f90:51 */
  void* T = orc_kdtree2_create(points, ncells);
  if (!T) return 1;
  int w[6];
  orc_box_window(g, box, w);
#pragma omp parallel for schedule(static)
  for (int i = w[0]; i <= w[1]; ++i) {
    for (int j = w[2]; j <= w[3]; ++j) {
      for (int k = w[4]; k <= w[5]; ++k) {
        double qv[3] = {g->xmin + (i - 1) * g->dx, g->ymin + (j - 1) * g->dy, g->zmin + (k - 1) * g->dz};
        size_t o = ((size_t)(i - 1) * g->ny + (size_t)(j - 1)) * g->nz + (size_t)(k - 1);
        if (pm) {
          if (!(fabs(vs[o] - pm[1]) < eps && fabs(vp[o] - pm[0]) < eps)) continue;
        }
        int idx = orc_kdtree2_nearest(T, qv, NULL);
        sites_id[o] = idx;
        vp[o] = params[(size_t)(idx - 1) * 3 + 0];
        vs[o] = params[(size_t)(idx - 1) * 3 + 1];
        rho[o] = params[(size_t)(idx - 1) * 3 + 2];
      }
    }
  }
  orc_kdtree2_destroy(T);
  return 0;
}

/* tree export for white-box tests of the product's own builder */
int orc_kdtree2_nnodes(const void* tp) { return ((const ktree*)tp)->nnodes; }
void orc_kdtree2_ind(const void* tp, int* ind) {
  const ktree* T = (const ktree*)tp;
  for (int i = 1; i <= T->n; ++i) ind[i - 1] = T->ind[i];
}

/* sites_locate: src/likelihood_body.F90:799-831 with point2idx :1038-1054.  queries (3,nq); out 1-based cell index. */
int orc_sites_locate(const double* points, int ncells, const int* sites_id, const orc_grid* g, const double* q, int64_t nq, int* out) {
  void* T = orc_kdtree2_create(points, ncells);
  if (!T) return 1;
  for (int64_t t = 0; t < nq; ++t) {
    const double* p = q + 3 * t;
    int ix = (int)floor((p[0] - g->xmin) / g->dx) + 1;
    int iy = (int)floor((p[1] - g->ymin) / g->dy) + 1;
    int iz = (int)floor((p[2] - g->zmin / g->scaling) / (g->dz / g->scaling)) + 1;
    if (ix < 1) ix = 1;
    if (iy < 1) iy = 1;
    if (iz < 1) iz = 1;
    if (ix >= g->nx) ix = g->nx - 1;
    if (iy >= g->ny) iy = g->ny - 1;
    if (iz >= g->nz) iz = g->nz - 1;
#define SID(k, j, i) sites_id[((size_t)((i) - 1) * g->ny + (size_t)((j) - 1)) * g->nz + (size_t)((k) - 1)]
    const int idx = SID(iz, iy, ix);
    int num = 0;
    for (int i = 1; i <= 2; ++i)
      for (int j = 1; j <= 2; ++j)
        for (int k = 1; k <= 2; ++k) num = num + abs(SID(iz + k - 1, iy + j - 1, ix + i - 1) - idx);
#undef SID
    out[t] = (num == 0) ? idx : orc_kdtree2_nearest(T, p, NULL);
  }
  orc_kdtree2_destroy(T);
  return 0;
}
