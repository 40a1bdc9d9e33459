// k1_column.cuh -- nearest-nucleus assignment by culled brute force: a block per tile of 8x8 grid
// columns, a warp per column, lanes along z.
//
// kdtree2's answer for a node is "the nucleus with the smallest squared distance sd = ((0+dx^2)+dy^2)+dz^2"
// (reference src/kdtree2.f90:1532-1538) whenever that minimum is unique by more than rounding noise;
// only for exact or few-ulp ties does the tree's traversal order decide.  So:
//   0. TILE CULL (block): one pass over all nuclei against the tile's xy-rectangle gives, per z-segment
//      s of the columns, a squared distance U_s that some nucleus is guaranteed not to exceed anywhere in
//      the tile (U_s = min_i (p_hi_i + max dz^2 over the segment)); nuclei whose closest approach p_lo_i
//      to the rectangle exceeds max_s U_s cannot win anywhere in the tile and are dropped.  What is left
//      (tens to a few hundred of possibly thousands) is the tile's candidate list in shared memory.
//   1. COLUMN CULL (warp): all nz nodes of a column share x and y, hence the partial sum
//      p_i = (0+dx^2)+dy^2 of every candidate -- computed once per column with exactly the reference's
//      operations and reused; the same bound, now with the column's exact p_i, leaves a few dozen
//      survivors, staged in the warp's slice of shared memory as (p_i, z_i, i).
//   2. NODES (lanes along z, so the four output streams are written as contiguous rows per warp): each
//      lane scans the staged survivors for its nodes, sd = p_i + dz^2, keeping the best and the
//      second-best.
//   3. A node whose two best distances are closer than 1e-12 relative is handed to the exact replay of
//      kdtree2's traversal (kd_nearest_dev): exact ties, duplicated nuclei, and the one-ulp corner in which
//      kdtree2's own pruning can discard the mathematically nearest point.
// Every bound is monotone in the rounded quantities it is built from and is inflated by 1e-9 relative,
// six orders of magnitude above any rounding error, so culling never removes a winner or a near-tie
// partner; the ORDER of candidates is irrelevant because anything order-dependent goes to step 3.
#pragma once

#define K1C_MAXC 192      // staged survivors per column
#define K1C_WARPS 16      // warps (columns in flight) per block
#define K1C_TILE 8        // tile = 8 x 8 columns
#define K1C_MAXT 2048     // tile candidates
#define K1C_MAXSEG 8
#define K1C_SMEM_BYTES ((2 * sizeof(double) + sizeof(int32_t)) * K1C_WARPS * K1C_MAXC + sizeof(int32_t) * K1C_MAXT)

__device__ __forceinline__ unsigned long long k1c_bits(double v) { return (unsigned long long)__double_as_longlong(v); }

__global__ void __launch_bounds__(32 * K1C_WARPS) k1_column_kernel(const __grid_constant__ K1Params P0) {
  // batch form: blockIdx.y selects the model; rebase the tree / nuclei / output pointers once
  K1Params P = P0;
  if (P0.models) {
    const K1Model M = P0.models[blockIdx.y];
    P.nodes = P0.nodes + M.node_off;
    P.rpts = P0.rpts + 3 * M.pt_off;
    P.ind = P0.ind + M.pt_off;
    P.params = P0.params + 3 * M.pt_off;
    P.root = M.root;
    P.n = M.n;
    const long long o = (long long)blockIdx.y * P0.model_stride;
    P.vp = P0.vp + o; P.vs = P0.vs + o; P.rho = P0.rho + o; P.sites = P0.sites + o;
  }
  // dynamic shared memory (K1C_SMEM_BYTES, above the 48 KB static limit): per-warp stages + the tile list
  extern __shared__ __align__(16) unsigned char k1c_smem[];
  double (*s_p)[K1C_MAXC] = reinterpret_cast<double (*)[K1C_MAXC]>(k1c_smem);
  double (*s_z)[K1C_MAXC] = reinterpret_cast<double (*)[K1C_MAXC]>(k1c_smem + sizeof(double) * K1C_WARPS * K1C_MAXC);
  int32_t (*s_id)[K1C_MAXC] = reinterpret_cast<int32_t (*)[K1C_MAXC]>(k1c_smem + 2 * sizeof(double) * K1C_WARPS * K1C_MAXC);
  int32_t* t_id = reinterpret_cast<int32_t*>(k1c_smem + (2 * sizeof(double) + sizeof(int32_t)) * K1C_WARPS * K1C_MAXC);
  __shared__ unsigned long long t_useg[K1C_MAXSEG];
  __shared__ int t_cnt;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int tiles_y = (P.wy + K1C_TILE - 1) / K1C_TILE;
  const int tiles_x = (P.wx + K1C_TILE - 1) / K1C_TILE;
  const int nseg = min(K1C_MAXSEG, max(1, P.wz / 6));
  const int seglen = (P.wz + nseg - 1) / nseg;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int tx0 = (tile / tiles_y) * K1C_TILE, ty0 = (tile % tiles_y) * K1C_TILE; // 0-based inside the window
    const int tw = min(K1C_TILE, P.wx - tx0), th = min(K1C_TILE, P.wy - ty0);
    // ---- 0. tile cull -------------------------------------------------------------------------------
    __syncthreads(); // previous tile fully consumed
    if (threadIdx.x < K1C_MAXSEG) t_useg[threadIdx.x] = k1c_bits(1.0e300);
    if (threadIdx.x == 0) t_cnt = 0;
    __syncthreads();
    const double rx0 = P.xmin + (double)(P.ix0 + tx0 - 1) * P.dx, rx1 = P.xmin + (double)(P.ix0 + tx0 + tw - 2) * P.dx;
    const double ry0 = P.ymin + (double)(P.iy0 + ty0 - 1) * P.dy, ry1 = P.ymin + (double)(P.iy0 + ty0 + th - 2) * P.dy;
    {
      double umin[K1C_MAXSEG];
#pragma unroll
      for (int s = 0; s < K1C_MAXSEG; ++s) umin[s] = 1.0e300;
      for (int n = threadIdx.x; n < P.n; n += blockDim.x) {
        const double X = __ldg(&P.rpts[3 * n + 0]), Y = __ldg(&P.rpts[3 * n + 1]), zn = __ldg(&P.rpts[3 * n + 2]);
        const double ax = fabs(X - rx0), bx = fabs(X - rx1), ay = fabs(Y - ry0), by = fabs(Y - ry1);
        const double hx = fmax(ax, bx), hy = fmax(ay, by);
        const double phi = (0.0 + hx * hx) + hy * hy; // >= p of this nucleus for every column of the tile
#pragma unroll
        for (int s = 0; s < K1C_MAXSEG; ++s) {
          if (s < nseg) {
            const int k0 = P.iz0 + s * seglen, k1 = min(P.iz0 + P.wz - 1, k0 + seglen - 1);
            const double za = P.zmin + (double)(k0 - 1) * P.dz, zb = P.zmin + (double)(k1 - 1) * P.dz;
            const double da = zn - za, db = zn - zb;
            umin[s] = fmin(umin[s], phi + fmax(da * da, db * db));
          }
        }
      }
#pragma unroll
      for (int s = 0; s < K1C_MAXSEG; ++s) {
        if (s < nseg) {
          double u = umin[s];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) u = fmin(u, __shfl_xor_sync(0xffffffffu, u, o));
          if (lane == 0) atomicMin(&t_useg[s], k1c_bits(u)); // positive doubles order like their bit patterns
        }
      }
    }
    __syncthreads();
    double tumax = 0.0;
    for (int s = 0; s < nseg; ++s) tumax = fmax(tumax, __longlong_as_double((long long)t_useg[s]));
    const double tbound = tumax * (1.0 + 1.0e-9) + 1.0e-300;
    for (int n0 = 0; n0 < P.n; n0 += blockDim.x) {
      const int n = n0 + threadIdx.x;
      bool keep = false;
      if (n < P.n) {
        const double X = __ldg(&P.rpts[3 * n + 0]), Y = __ldg(&P.rpts[3 * n + 1]);
        // closest approach to the rectangle: 0 inside, else distance to the nearer edge (monotone in the rounded terms)
        const double lx = (X < rx0) ? (rx0 - X) : ((X > rx1) ? (X - rx1) : 0.0);
        const double ly = (Y < ry0) ? (ry0 - Y) : ((Y > ry1) ? (Y - ry1) : 0.0);
        const double plo = (0.0 + lx * lx) + ly * ly;
        keep = plo <= tbound;
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      int base = 0;
      if (lane == 0 && m) base = atomicAdd(&t_cnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      if (keep && pos < K1C_MAXT) t_id[pos] = n;
    }
    __syncthreads();
    const int tn_all = t_cnt;
    const bool tile_list = tn_all <= K1C_MAXT; // else: too many tile candidates, columns scan all nuclei
    const int nlist = tile_list ? tn_all : P.n;
    // ---- columns of the tile: one per warp at a time -----------------------------------------------------
    for (int cidx = wib; cidx < tw * th; cidx += K1C_WARPS) {
      const int i = P.ix0 + tx0 + cidx / th, j = P.iy0 + ty0 + cidx % th;
      const double qx = P.xmin + (double)(i - 1) * P.dx; // mcmc_loc2.f90:2054
      const double qy = P.ymin + (double)(j - 1) * P.dy;
      const size_t obase = ((size_t)(i - P.ia0) * P.ny_a + (size_t)(j - P.ja0)) * P.nz_a + (size_t)(P.iz0 - P.ka0);
      // 1a. U_s with the column's exact p
      double umin[K1C_MAXSEG];
#pragma unroll
      for (int s = 0; s < K1C_MAXSEG; ++s) umin[s] = 1.0e300;
      for (int t = lane; t < nlist; t += 32) {
        const int n = tile_list ? t_id[t] : t;
        const double dx = __ldg(&P.rpts[3 * n + 0]) - qx;
        const double dy = __ldg(&P.rpts[3 * n + 1]) - qy;
        const double zn = __ldg(&P.rpts[3 * n + 2]);
        const double p = (0.0 + dx * dx) + dy * dy;
#pragma unroll
        for (int s = 0; s < K1C_MAXSEG; ++s) {
          if (s < nseg) {
            const int k0 = P.iz0 + s * seglen, k1 = min(P.iz0 + P.wz - 1, k0 + seglen - 1);
            const double za = P.zmin + (double)(k0 - 1) * P.dz, zb = P.zmin + (double)(k1 - 1) * P.dz;
            const double da = zn - za, db = zn - zb;
            umin[s] = fmin(umin[s], p + fmax(da * da, db * db));
          }
        }
      }
      double umax = 0.0;
#pragma unroll
      for (int s = 0; s < K1C_MAXSEG; ++s) {
        if (s < nseg) {
          double u = umin[s];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) u = fmin(u, __shfl_xor_sync(0xffffffffu, u, o));
          umax = fmax(umax, u);
        }
      }
      const double bound = umax * (1.0 + 1.0e-9) + 1.0e-300;
      // 1b. stage the survivors
      int ncand = 0;
      for (int t0 = 0; t0 < nlist; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        double p = 0.0, zn = 0.0;
        int n = 0;
        if (t < nlist) {
          n = tile_list ? t_id[t] : t;
          const double dx = __ldg(&P.rpts[3 * n + 0]) - qx;
          const double dy = __ldg(&P.rpts[3 * n + 1]) - qy;
          zn = __ldg(&P.rpts[3 * n + 2]);
          p = (0.0 + dx * dx) + dy * dy;
          keep = p <= bound;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = ncand + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < K1C_MAXC) {
          s_p[wib][pos] = p;
          s_z[wib][pos] = zn;
          s_id[wib][pos] = n;
        }
        ncand += __popc(m);
      }
      __syncwarp();
      const bool overflow = ncand > K1C_MAXC; // too many survivors to stage: exact tree walk for the whole column
      // 2. nodes: lanes along z
      for (int kz = lane; kz < P.wz; kz += 32) {
        const int k = P.iz0 + kz;
        const size_t o = obase + (size_t)kz;
        if (P.use_pm) { // mcmc_loc2.f90:2055-2056
          if (!(fabs(P.vs[o] - P.pm_vs) < P.pm_eps && fabs(P.vp[o] - P.pm_vp) < P.pm_eps)) continue;
        }
        const double qz = P.zmin + (double)(k - 1) * P.dz;
        int idx = 0;
        bool exact = overflow;
        if (!overflow) {
          double b1 = 1.0e300, b2 = 1.0e300;
          int i1 = 0;
          for (int c = 0; c < ncand; ++c) {
            const double dz = s_z[wib][c] - qz;
            const double sd = s_p[wib][c] + dz * dz;
            if (sd < b1) { b2 = b1; b1 = sd; i1 = c; }
            else if (sd < b2) b2 = sd;
          }
          exact = !(b2 > b1 * (1.0 + 1.0e-12) + 1.0e-300); // (near-)tie: let kdtree2's traversal decide
          idx = __ldg(&P.ind[s_id[wib][i1]]);
        }
        if (exact) idx = kd_nearest_dev(P, qx, qy, qz, P.err); // 3.
        P.sites[o] = idx;
        const double* pr = P.params + 3 * (size_t)(idx - 1);
        P.vp[o] = __ldg(&pr[0]);
        P.vs[o] = __ldg(&pr[1]);
        P.rho[o] = __ldg(&pr[2]);
      }
      __syncwarp();
    }
  }
}
