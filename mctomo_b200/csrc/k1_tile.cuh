// k1_tile.cuh -- nearest-nucleus assignment, round-2 shape: per-(tile, z-segment) candidate lists in shared memory,
// nodes flattened over the tile's columns, 16-byte vector stores.
//
// What must come out is kdtree2's answer (reference src/kdtree2.f90:1028-1069,1369-1443,1496-1599): the nucleus with
// the smallest sd = ((0+dx^2)+dy^2)+dz^2, computed with exactly those operations, and -- when two candidates are
// within rounding noise of each other -- whichever of them kdtree2's traversal meets last.  As in k1_column.cuh the
// second case is detected (best and second-best within 1e-12 relative) and handed to kd_nearest_dev, the replay of
// kdtree2's traversal; everything else is a minimum over a candidate set that provably contains every nucleus within
// 1e-9 relative of the minimum.  What changed is how small that set is before a node looks at it (the round-1 kernel
// let every node scan all survivors of its whole column: ~57 warp instructions per node, 5.7 % of the HBM rate):
//
//   0. BOX LISTS (block = a tile of tx x ty columns, cut into <= 32 z-segments): for every (tile, segment) box B,
//      U_B = min_i max_{q in B} |q - x_i|^2 bounds the nearest distance of every node in B from above, so only nuclei
//      with min_{q in B} |q - x_i|^2 <= U_B can win -- or tie -- there.  Three passes over the nuclei: (a) the nucleus
//      nearest to the whole tile in the plane, (b) U_B from the nuclei within ~2.5 spacings of it (any subset yields a
//      valid upper bound; the far ones are spared the segment loop), (c) the filter, with a plane-only pre-test.  The
//      culling arithmetic is float32 on coordinates relative to the tile's centre, every bound pushed 1e-4 the safe
//      way.  The union is staged in shared memory as (X, Y, Z, index) records; each box keeps a list of positions.
//      The host sizes tile and segment to about half the nucleus spacing, which leaves ~5-10 candidates per box.
//   1. NODES: the tile's nodes flattened as (column, z-pair) (index arithmetic by multiply-shift: the divisors are
//      kernel-uniform); a thread resolves two consecutive z against its box's list with the reference's float64
//      arithmetic, tracking best and second best, and writes vp/vs/rho as double2 and sites_id as int2.  Columns
//      adjacent in y are adjacent in memory, so a warp's stores are contiguous runs whatever nz is; odd nz (columns
//      starting on odd element offsets) only shifts the pairing by one node per column.
// A box whose list overflows, or a tile whose union does, takes the exact tree walk for its nodes: no input can produce
// a wrong cell, only a slower one.
#pragma once

#define K1T_MAXT 512   // staged candidates per tile (union over its boxes)
#define K1T_MAXSEG 32  // z-segments per tile
#define K1T_L 48       // entries of a box list
#define K1T_MAXN 1024  // nuclei near the tile in the plane

struct __align__(16) K1TShared {
  double4 rec[K1T_MAXT];                         // X, Y, Z, original 1-based nucleus index (as bits)
  unsigned short list[K1T_MAXSEG][K1T_L];        // staged positions of each box's candidates
  unsigned int useg[K1T_MAXSEG];                 // U_B as float bits (positive floats order like their bits)
  float smid[K1T_MAXSEG], shalf_up[K1T_MAXSEG], shalf_dn[K1T_MAXSEG]; // segment centres / half lengths, relative to the column's middle
  int cnt[K1T_MAXSEG];
  unsigned int phimin;
  int tcnt, ncnt;
  float4 near[K1T_MAXN];                         // nuclei near the tile in the plane: tile-relative float coordinates, position
};

// x / d for 0 <= x < 2^16 and a kernel-uniform divisor d <= 2^15: one multiply, one shift.
// mul = ceil(2^32 / d): x * mul >> 32 == x / d for all x < 2^16 (error term x * (mul*d - 2^32) / d < 2^16 * d / d ... < 2^32/d).
struct K1TDiv { unsigned mul; };
__device__ __forceinline__ int k1t_div(int x, K1TDiv m) { return m.mul ? (int)__umulhi((unsigned)x, m.mul) : x; } // mul == 0: d == 1
static inline K1TDiv k1t_magic(int d) { return K1TDiv{d <= 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d)}; }
// (d == 1: mul would not fit 32 bits; encoded as 0)

// ttx, tty: tile size in columns; seglen: nodes per z-segment ((wz + seglen - 1) / seglen <= K1T_MAXSEG); near_r: ~2.5
// nucleus spacings (only steers which nuclei feed the upper bounds: any value is correct); dv_*: multiply-shift
// constants of npairs, tty and seglen.
__global__ void __launch_bounds__(256, 2) k1_tile_kernel(const __grid_constant__ K1Params P0, int ttx, int tty, int seglen, float near_r,
                                                         K1TDiv dv_pairs, K1TDiv dv_ty, K1TDiv dv_seg) {
  K1Params P = P0;
  if (P0.models) { // batch form: blockIdx.y selects the model
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
  __shared__ K1TShared S;
  const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
  const unsigned lt = (1u << lane) - 1u;
  const int tiles_y = (P.wy + tty - 1) / tty;
  const int tiles_x = (P.wx + ttx - 1) / ttx;
  const int nseg = (P.wz + seglen - 1) / seglen;
  const double zfirst = P.zmin + (double)(P.iz0 - 1) * P.dz, zlast = P.zmin + (double)(P.iz0 + P.wz - 2) * P.dz;
  const double zc = 0.5 * (zfirst + zlast);
  const int npairs = (P.wz + 2) / 2; // (column, pair) items; a pair is two z-adjacent nodes on a 16-byte boundary
  // segment ends are bracketed with a hair of slack instead of reproducing the nodes' own z = zmin + (k-1)*dz roundings
  const double zslack = 1.0e-9 * (fabs(P.zmin) + fabs(zlast) + P.dz) + 1.0e-300;
  // 16-byte stores need all four arrays in the same phase (true for any sane allocation; else scalar stores)
  const unsigned ph = (unsigned)(((uintptr_t)P.vp >> 3) & 1);
  const bool vec_ok = (((uintptr_t)P.vs >> 3) & 1) == ph && (((uintptr_t)P.rho >> 3) & 1) == ph && (((uintptr_t)P.sites >> 2) & 1) == ph &&
                      ((uintptr_t)P.vp & 7) == 0 && ((uintptr_t)P.vs & 7) == 0 && ((uintptr_t)P.rho & 7) == 0 && ((uintptr_t)P.sites & 3) == 0;
  if (tid < nseg) {
    const int k0 = tid * seglen, k1 = min(P.wz - 1, k0 + seglen - 1);
    const double za = zfirst + (double)k0 * P.dz - zslack, zb = zfirst + (double)k1 * P.dz + zslack;
    S.smid[tid] = (float)(0.5 * (za + zb) - zc);
    // the float centre is off by up to one float ulp of its magnitude: widen / narrow the half length accordingly
    const float e = 1.0e-6f * (float)(fabs(za - zc) + fabs(zb - zc) + (zb - za));
    S.shalf_up[tid] = (float)(0.5 * (zb - za)) * 1.0001f + e + 1e-30f;
    S.shalf_dn[tid] = fmaxf((float)(0.5 * (zb - za)) * 0.9999f - e, 0.f);
  }

  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int tx0 = (tile / tiles_y) * ttx, ty0 = (tile % tiles_y) * tty; // 0-based inside the window
    const int tw = min(ttx, P.wx - tx0), th = min(tty, P.wy - ty0);
    // ---- 0. box lists ---------------------------------------------------------------------------------------
    __syncthreads(); // previous tile fully consumed (and the segment table written)
    if (tid < K1T_MAXSEG) { S.useg[tid] = 0x7f000000u; S.cnt[tid] = 0; }
    if (tid == 0) { S.tcnt = 0; S.ncnt = 0; S.phimin = 0x7f000000u; }
    __syncthreads();
    const double rx0 = P.xmin + (double)(P.ix0 + tx0 - 1) * P.dx, rx1 = P.xmin + (double)(P.ix0 + tx0 + tw - 2) * P.dx;
    const double ry0 = P.ymin + (double)(P.iy0 + ty0 - 1) * P.dy, ry1 = P.ymin + (double)(P.iy0 + ty0 + th - 2) * P.dy;
    const double cx = 0.5 * (rx0 + rx1), cy = 0.5 * (ry0 + ry1);
    const float hxf = (float)(0.5 * (rx1 - rx0)) * 1.0001f + 1e-30f, hyf = (float)(0.5 * (ry1 - ry0)) * 1.0001f + 1e-30f; // half extents, rounded up
    const float hxl = (float)(0.5 * (rx1 - rx0)) * 0.9999f, hyl = (float)(0.5 * (ry1 - ry0)) * 0.9999f;                 // ... rounded down
    // a. the smallest "farthest corner" distance: the nucleus that is certainly near every column of the tile
    {
      float pm = 3.0e38f;
      for (int n = tid; n < P.n; n += nthr) {
        const float ax = fabsf((float)(__ldg(&P.rpts[3 * n + 0]) - cx)) + hxf, ay = fabsf((float)(__ldg(&P.rpts[3 * n + 1]) - cy)) + hyf;
        pm = fminf(pm, ax * ax + ay * ay);
      }
      const unsigned hmin = __reduce_min_sync(0xffffffffu, __float_as_uint(pm));
      if (lane == 0) atomicMin(&S.phimin, hmin);
    }
    __syncthreads();
    // a'. compact the nuclei whose nearest approach to the rectangle is within ~1.25 spacings of that one: the only ones
    //     the segment loops below look at (relative float coordinates + position, 16 bytes each)
    const float r1 = sqrtf(__uint_as_float(S.phimin)) * 1.0001f + near_r;
    const float phi_cut = r1 * r1;
    for (int n0 = 0; n0 < P.n; n0 += nthr) {
      const int n = n0 + tid;
      bool near = false;
      float rx = 0.f, ry = 0.f;
      if (n < P.n) {
        rx = (float)(__ldg(&P.rpts[3 * n + 0]) - cx); ry = (float)(__ldg(&P.rpts[3 * n + 1]) - cy);
        const float lx = fmaxf(fabsf(rx) - hxl, 0.f), ly = fmaxf(fabsf(ry) - hyl, 0.f);
        near = (lx * lx + ly * ly) * 0.9999f <= phi_cut;
      }
      const unsigned m = __ballot_sync(0xffffffffu, near);
      if (m == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&S.ncnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int pos = base + __popc(m & lt);
      if (near && pos < K1T_MAXN) S.near[pos] = make_float4(rx, ry, (float)(__ldg(&P.rpts[3 * n + 2]) - zc), __int_as_float(n));
    }
    __syncthreads();
    const int N = S.ncnt;
    const bool near_ok = N <= K1T_MAXN;
    // b. U_B per segment: lane = segment, the warps share the list (any subset of the nuclei gives a valid upper bound;
    //    the nucleus of step a is in this one, so every U_B is finite)
    if (near_ok) {
      // up to 16 segments: the two half-warps split the list between them (lane & 15 = segment)
      const int split = nseg <= 16 ? 2 : 1;
      const int sub = split == 2 ? (lane >> 4) : 0;
      const int wib = (tid >> 5) * split + sub, nw = (nthr >> 5) * split;
      for (int s = (split == 2 ? (lane & 15) : lane); s < nseg; s += 32) {
        const float mid = S.smid[s], hu = S.shalf_up[s];
        float um = 3.0e38f;
        for (int e = wib; e < N; e += nw) {
          const float4 r = S.near[e];
          const float ax = fabsf(r.x) + hxf, ay = fabsf(r.y) + hyf;
          const float phi = (ax * ax + ay * ay) * 1.0001f; // >= (0+dx^2)+dy^2 of this nucleus for every column of the tile
          const float dh = fabsf(r.z - mid) + hu;          // farthest end of the segment
          um = fminf(um, fmaf(dh, dh, phi));
        }
        if (um < 1.0e38f) atomicMin(&S.useg[s], __float_as_uint(um));
      }
    }
    __syncthreads();
    // c. the filter: nearest approach to the box against U_B
    {
      float umax = 0.f;
      for (int s = 0; s < nseg; ++s) umax = fmaxf(umax, __uint_as_float(S.useg[s]) * 1.0002f);
      const bool from_list = near_ok && umax <= phi_cut; // else (a tile much wider than the nucleus spacing): scan them all
      const int total = from_list ? N : P.n;
      for (int e0 = 0; e0 < total; e0 += nthr) {
        const int e = e0 + tid;
        unsigned mask = 0;
        int n = 0;
        if (e < total) {
          float rx, ry, rz;
          if (from_list) { const float4 r = S.near[e]; rx = r.x; ry = r.y; rz = r.z; n = __float_as_int(r.w); }
          else {
            n = e;
            rx = (float)(__ldg(&P.rpts[3 * n + 0]) - cx); ry = (float)(__ldg(&P.rpts[3 * n + 1]) - cy); rz = (float)(__ldg(&P.rpts[3 * n + 2]) - zc);
          }
          const float lx = fmaxf(fabsf(rx) - hxl, 0.f), ly = fmaxf(fabsf(ry) - hyl, 0.f);
          const float plo = (lx * lx + ly * ly) * 0.9999f; // closest approach to the tile's rectangle, rounded down
          if (plo <= umax) {
            for (int s = 0; s < nseg; ++s) {
              const float dl = fmaxf(fabsf(rz - S.smid[s]) - S.shalf_dn[s], 0.f);
              if (fmaf(dl * 0.9999f, dl, plo) <= __uint_as_float(S.useg[s]) * 1.0002f) mask |= 1u << s;
            }
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, mask != 0);
        if (m == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&S.tcnt, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(m & lt);
        if (mask != 0 && pos < K1T_MAXT) {
          S.rec[pos] = make_double4(__ldg(&P.rpts[3 * n + 0]), __ldg(&P.rpts[3 * n + 1]), __ldg(&P.rpts[3 * n + 2]),
                                    __longlong_as_double((long long)__ldg(&P.ind[n])));
          while (mask) {
            const int s = __ffs(mask) - 1;
            mask &= mask - 1;
            const int q = atomicAdd(&S.cnt[s], 1);
            if (q < K1T_L) S.list[s][q] = (unsigned short)pos;
          }
        }
      }
    }
    __syncthreads();
    const bool tile_ok = S.tcnt <= K1T_MAXT; // else: every node of the tile takes the exact tree walk
    // ---- 1. nodes: (column of the FULL tile shape, z-pair) items; columns outside a clipped edge tile are skipped -----
    const int nitems = tw * tty * npairs;
    for (int item = tid; item < nitems; item += nthr) {
      const int c = k1t_div(item, dv_pairs), mpair = item - c * npairs;
      const int ci = k1t_div(c, dv_ty), cj = c - ci * tty;
      if (cj >= th) continue;
      const int i = P.ix0 + tx0 + ci, j = P.iy0 + ty0 + cj;
      const double qx = P.xmin + (double)(i - 1) * P.dx; // mcmc_loc2.f90:2054
      const double qy = P.ymin + (double)(j - 1) * P.dy;
      const size_t obase = ((size_t)(i - P.ia0) * P.ny_a + (size_t)(j - P.ja0)) * P.nz_a + (size_t)(P.iz0 - P.ka0);
      const int par = (int)((obase + ph) & 1); // pairs start where the absolute address is a multiple of 16
      const int e0 = 2 * mpair - par;          // elements e0, e0+1 of the column's window (0-based)
      int idx[2] = {0, 0};
      bool on[2];
      double qz[2];
      int sg[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kz = e0 + h;
        on[h] = kz >= 0 && kz < P.wz;
        qz[h] = P.zmin + (double)(P.iz0 + kz - 1) * P.dz;
        sg[h] = on[h] ? k1t_div(kz, dv_seg) : 0;
        if (on[h] && P.use_pm) { // mcmc_loc2.f90:2055-2056: only nodes still carrying the moved cell's old values
          const size_t o = obase + (size_t)kz;
          if (!(fabs(P.vs[o] - P.pm_vs) < P.pm_eps && fabs(P.vp[o] - P.pm_vp) < P.pm_eps)) on[h] = false;
        }
      }
      if (!on[0] && !on[1]) continue;
      bool exact[2] = {false, false};
      if (on[0] && on[1] && sg[0] == sg[1]) {
        // both nodes in one box (the usual case: segments are an even number of nodes long): one walk over its list
        const int cnt = S.cnt[sg[0]];
        if (!tile_ok || cnt > K1T_L) exact[0] = exact[1] = true;
        else {
          double a1 = 1.0e300, a2 = 1.0e300, c1 = 1.0e300, c2 = 1.0e300;
          const unsigned short* L = S.list[sg[0]];
          int pos = cnt > 0 ? L[0] : 0;
          for (int q = 0; q < cnt; ++q) {
            const double4 r = S.rec[pos];
            if (q + 1 < cnt) pos = L[q + 1];
            const int id = (int)__double_as_longlong(r.w);
            const double dx = r.x - qx, dy = r.y - qy, dza = r.z - qz[0], dzb = r.z - qz[1];
            double p = 0.0 + dx * dx; // kdtree2.f90:1534-1538, left to right
            p = p + dy * dy;
            const double sa = p + dza * dza, sb = p + dzb * dzb;
            if (sa < a1) { a2 = a1; a1 = sa; idx[0] = id; } else if (sa < a2) a2 = sa;
            if (sb < c1) { c2 = c1; c1 = sb; idx[1] = id; } else if (sb < c2) c2 = sb;
          }
          exact[0] = !(a2 > a1 * (1.0 + 1.0e-12) + 1.0e-300); // (near-)tie: kdtree2's traversal decides
          exact[1] = !(c2 > c1 * (1.0 + 1.0e-12) + 1.0e-300);
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!on[h]) continue;
          const int cnt = S.cnt[sg[h]];
          if (!tile_ok || cnt > K1T_L) { exact[h] = true; continue; }
          double b1 = 1.0e300, b2 = 1.0e300;
          const unsigned short* L = S.list[sg[h]];
          for (int q = 0; q < cnt; ++q) {
            const double4 r = S.rec[L[q]];
            const double dx = r.x - qx, dy = r.y - qy, dz = r.z - qz[h];
            double sd = 0.0 + dx * dx;
            sd = sd + dy * dy;
            sd = sd + dz * dz;
            if (sd < b1) { b2 = b1; b1 = sd; idx[h] = (int)__double_as_longlong(r.w); }
            else if (sd < b2) b2 = sd;
          }
          exact[h] = !(b2 > b1 * (1.0 + 1.0e-12) + 1.0e-300);
        }
      }
      double pvp[2], pvs[2], prho[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!on[h]) continue;
        if (exact[h]) idx[h] = kd_nearest_dev(P, qx, qy, qz[h], P.err);
        const double* pr = P.params + 3 * (size_t)(idx[h] - 1);
        pvp[h] = __ldg(&pr[0]); pvs[h] = __ldg(&pr[1]); prho[h] = __ldg(&pr[2]);
      }
      const size_t o0 = obase + (size_t)(e0 < 0 ? 0 : e0);
      if (on[0] && on[1] && vec_ok) { // 16-byte stores (8 for sites_id); aligned by construction of `par`
        *reinterpret_cast<int2*>(P.sites + o0) = make_int2(idx[0], idx[1]);
        *reinterpret_cast<double2*>(P.vp + o0) = make_double2(pvp[0], pvp[1]);
        *reinterpret_cast<double2*>(P.vs + o0) = make_double2(pvs[0], pvs[1]);
        *reinterpret_cast<double2*>(P.rho + o0) = make_double2(prho[0], prho[1]);
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!on[h]) continue;
          const size_t o = obase + (size_t)(e0 + h);
          P.sites[o] = idx[h]; P.vp[o] = pvp[h]; P.vs[o] = pvs[h]; P.rho[o] = prho[h];
        }
      }
    }
  }
}
