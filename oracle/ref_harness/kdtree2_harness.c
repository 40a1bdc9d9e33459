/*
 * oracle/ref_harness/kdtree2_harness.c -- TEST INFRASTRUCTURE.
 *
 * Driver for the reference's OWN pre-built kd-tree object (utils/libutils.a:kdtree2.o,
 * GCC 4.8.5 gfortran; source utils/kdtree2.f90 == src/kdtree2.f90).  It builds hand-made
 * gfortran-4.8 array descriptors and calls
 *     kdtree2_create(points(:,1:n), sort=.false., rearrange=.true.)   mcmc_loc2.f90:2029
 *     kdtree2_n_nearest(tp=tree, qv=qv, nn=1, results=results)        mcmc_loc2.f90:2057
 *     kdtree2_destroy(tree)                                           mcmc_loc2.f90:2078
 * exactly as kdtree_to_grid does.  No reference source is copied: the object is linked
 * where it lies (see oracle/build_ref.sh); the executable lands in oracle/_ref/.
 *
 * Usage: kdtree2_ref <in.bin> <out.bin> [dim]     dim = 3 (default) or 2 (the 2-D product, mcmc2d/mcmc.f90:1469-1556)
 *   in : int64 n, int64 nq, double points[dim*n], double queries[dim*nq]
 *   out: int32 idx[nq] (1-based), double dis[nq]
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef struct { ptrdiff_t stride, lbound, ubound; } gf_dim;
typedef struct { void* base; ptrdiff_t offset; ptrdiff_t dtype; gf_dim dim[2]; } gf_desc2;
typedef struct { void* base; ptrdiff_t offset; ptrdiff_t dtype; gf_dim dim[1]; } gf_desc1;
typedef struct { double dis; int idx; int pad; } kd_result; /* type kdtree2_result, 16 B */

/* dtype = rank | type<<3 | elemsize<<6 ; real = 3, derived = 5 */
#define DT_R8_RANK2 ((ptrdiff_t)(2 | (3 << 3) | (8 << 6)))
#define DT_R8_RANK1 ((ptrdiff_t)(1 | (3 << 3) | (8 << 6)))
#define DT_DERIVED16_RANK1 ((ptrdiff_t)(1 | (5 << 3) | (16 << 6)))

extern void* __kdtree2_module_MOD_kdtree2_create(gf_desc2* input, int* dim, int* sort, int* rearrange);
extern void __kdtree2_module_MOD_kdtree2_n_nearest(void** tp, gf_desc1* qv, int* nn, gf_desc1* results);
extern void __kdtree2_module_MOD_kdtree2_destroy(void** tp);

int main(int argc, char** argv) {
  if (argc != 3 && argc != 4) { fprintf(stderr, "usage: %s in.bin out.bin [dim]\n", argv[0]); return 2; }
  const int D = argc == 4 ? atoi(argv[3]) : 3;
  if (D != 2 && D != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("in"); return 2; }
  int64_t n, nq;
  if (fread(&n, 8, 1, f) != 1 || fread(&nq, 8, 1, f) != 1) return 2;
  double* pts = malloc(sizeof(double) * D * (size_t)n);
  double* q = malloc(sizeof(double) * D * (size_t)nq);
  if (fread(pts, 8, D * (size_t)n, f) != D * (size_t)n) return 2;
  if (fread(q, 8, D * (size_t)nq, f) != D * (size_t)nq) return 2;
  fclose(f);

  gf_desc2 d;
  d.base = pts;
  d.dtype = DT_R8_RANK2;
  d.dim[0].stride = 1; d.dim[0].lbound = 1; d.dim[0].ubound = D;
  d.dim[1].stride = D; d.dim[1].lbound = 1; d.dim[1].ubound = n;
  d.offset = -(1 * 1 + 1 * D);
  int sort = 0, rearr = 1;
  void* tree = __kdtree2_module_MOD_kdtree2_create(&d, NULL, &sort, &rearr);

  int32_t* idx = malloc(sizeof(int32_t) * (size_t)nq);
  double* dis = malloc(sizeof(double) * (size_t)nq);
  for (int64_t i = 0; i < nq; ++i) {
    kd_result res[1];
    gf_desc1 dq, dr;
    dq.base = q + D * i; dq.dtype = DT_R8_RANK1; dq.offset = -1;
    dq.dim[0].stride = 1; dq.dim[0].lbound = 1; dq.dim[0].ubound = D;
    dr.base = res; dr.dtype = DT_DERIVED16_RANK1; dr.offset = -1;
    dr.dim[0].stride = 1; dr.dim[0].lbound = 1; dr.dim[0].ubound = 1;
    int nn = 1;
    __kdtree2_module_MOD_kdtree2_n_nearest(&tree, &dq, &nn, &dr);
    idx[i] = res[0].idx;
    dis[i] = res[0].dis;
  }
  __kdtree2_module_MOD_kdtree2_destroy(&tree);
  f = fopen(argv[2], "wb");
  if (!f) { perror("out"); return 2; }
  fwrite(idx, 4, (size_t)nq, f);
  fwrite(dis, 8, (size_t)nq, f);
  fclose(f);
  return 0;
}
