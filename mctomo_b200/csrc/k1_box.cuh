// k1_box.cuh -- nearest-nucleus assignment, third shape (round 2, second half): the box lists of k1_tile.cuh, but
//   * a warp's lanes work on ONE box (items ordered segment-major), so the walk over the box's candidate list has a
//     warp-uniform trip count and every record load is a shared-memory broadcast (k1_tile_kernel: 19 of 32 lanes
//     active, the lanes of a warp spanning 10-20 boxes with different list lengths);
//   * a thread resolves NPT = 4 (or 2) z-adjacent nodes of its column in one walk: the column's index arithmetic and the
//     in-plane part of every distance are shared between them;
//   * the walk itself is float32: tile-relative coordinates (node coordinates from per-tile / per-block tables in shared
//     memory), smallest and second-smallest squared distance tracked with three FMNMX per node.  The float32 winner is
//     kdtree2's winner whenever the gap to the runner-up exceeds a rigorous bound of the float32 evaluation error
//     (below); the node then never touches the FP64 pipe.  Everything else -- true near-ties, and the ~1e-5 of the
//     nodes that sit within float32 noise of a cell wall -- takes kd_nearest_dev, the replay of kdtree2's own
//     traversal (reference src/kdtree2.f90:1028-1069,1369-1443,1496-1599), which is exact by construction;
//   * one persistent grid over (model, tile) pairs; everything kernel-uniform is computed on the host.
//
// Error bound.  Let a = float(nucleus - centre), q = float(node - centre) (each one rounding of an exactly formed
// float64 difference, |a|,|q| <= B per coordinate), D = fl(a - q).  |D - true| <= 2^-24 (|a| + |q| + |D|), so for
// s = fl(sum D^2) against the reference's float64 sd: |s - sd| <= 2^-24 (7 B sqrt(sd) + 7.5 sd) <= 2^-24 (3.5 B^2 + 11 sd).
// The float32 argmin is the float64 argmin if for every other candidate s_j - s_1 > err_1 + err_j, which holds for all
// j once it holds for the runner-up m2: we demand m2 - m1 > 2^-20 (B^2 + m1 + m2) (16 x 2^-24: margin over 7 and 11).
// The float64 tie rule of the reference (1e-12 relative) is far inside that band, so exact ties always reach the walk.
#pragma once

#define K1B_MAXT 512  // staged candidates per tile (union over its boxes)
#define K1B_MAXN 768  // nuclei near the tile in the plane
#define K1B_MAXZ 512  // nodes of a column window (larger windows use the column kernel)
#define K1B_MAXC 64   // tile extent in x or y

struct K1BGeom { // kernel-uniform quantities, all from the host
  int ttx, tty, seglen, nseg, tiles_x, tiles_y, tiles_per_model, total_tiles;
  int nunits;    // NPT-node units per column ((wz + NPT) / NPT: one spare for the shifted pairing)
  int ups;       // units per segment (seglen / NPT)
  int ngroups;   // unit groups per column (ceil(nunits / ups))
  int per_group; // ttx * tty * ups
  float near_r;
  K1TDiv dv_group, dv_ups, dv_ty, dv_seg, dv_tpm, dv_tiles_y;
};

struct __align__(16) K1BShared {
  float4 rec[K1B_MAXT];                          // tile-relative X, Y, Z (float32), ORIGINAL 1-based nucleus index (as bits)
  float4 near[K1B_MAXN];
  unsigned short list[K1T_MAXSEG][K1T_L];        // staged positions of each box's candidates
  unsigned long long useg[K1T_MAXSEG];           // U_B as float bits (high word) and the near-list entry that attains it
  float smid[K1T_MAXSEG], shalf_up[K1T_MAXSEG], shalf_dn[K1T_MAXSEG];
  int cnt[K1T_MAXSEG];
  float fz[K1B_MAXZ];                            // node z - column middle, per block
  float fx[K1B_MAXC], fy[K1B_MAXC];              // node x, y - tile centre, per tile
  unsigned int phimin, bmax;
  int tcnt, ncnt;
};

// one node against one box list (units that straddle two segments, clipped units)
__device__ __forceinline__ void k1b_scan1(const K1BShared& S, int seg, int cnt, float fx, float fy, float fz, float& m1, float& m2, int& pos) {
  m1 = 3.0e38f; m2 = 3.0e38f; pos = 0;
  const unsigned short* L = S.list[seg];
  for (int q = 0; q < cnt; ++q) {
    const int e = L[q];
    const float4 r = S.rec[e];
    const float dx = r.x - fx, dy = r.y - fy, dz = r.z - fz;
    const float s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    pos = s < m1 ? e : pos;
    m2 = fminf(m2, fmaxf(m1, s));
    m1 = fminf(m1, s);
  }
}

// the exact tree walk of one node (cold path); the per-model pointers come in explicitly, the geometry from the launch
// parameters
__device__ __noinline__ int k1b_walk(const K1Params& P0, const KdNodeDev* nodes, const double* rpts, const int32_t* ind, int root, int n,
                                     int i, int j, int kz) {
  K1Params Q = P0;
  Q.nodes = nodes; Q.rpts = rpts; Q.ind = ind; Q.root = root; Q.n = n;
  const double qx = Q.xmin + (double)(i - 1) * Q.dx; // mcmc_loc2.f90:2054
  const double qy = Q.ymin + (double)(j - 1) * Q.dy;
  const double qz = Q.zmin + (double)(Q.iz0 + kz - 1) * Q.dz;
  return kd_nearest_dev(Q, qx, qy, qz, Q.err);
}

template <int NPT>
__global__ void __launch_bounds__(256, 4) k1_box_kernel(const __grid_constant__ K1Params P0, const __grid_constant__ K1BGeom G) {
  __shared__ K1BShared S;
  const K1Params& P = P0; // geometry and window: kernel-uniform (constant bank); per-model pointers are the m_* locals
  const KdNodeDev* m_nodes = P0.nodes;
  const double* m_rpts = P0.rpts;
  const int32_t* m_ind = P0.ind;
  const double* m_params = P0.params;
  int m_root = P0.root, m_n = P0.n;
  double *m_vp = P0.vp, *m_vs = P0.vs, *m_rho = P0.rho;
  int32_t* m_sites = P0.sites;
  const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
  const unsigned lt = (1u << lane) - 1u;
  const int ttx = G.ttx, tty = G.tty, seglen = G.seglen, nseg = G.nseg;
  const double zfirst = P.zmin + (double)(P.iz0 - 1) * P.dz, zlast = P.zmin + (double)(P.iz0 + P.wz - 2) * P.dz;
  const double zc = 0.5 * (zfirst + zlast);
  const double zslack = 1.0e-9 * (fabs(P.zmin) + fabs(zlast) + P.dz) + 1.0e-300;
  const float hzf = (float)(0.5 * (zlast - zfirst)) * 1.0001f + 1e-30f;
  if (tid < nseg) {
    const int k0 = tid * seglen, k1 = min(P.wz - 1, k0 + seglen - 1);
    const double za = zfirst + (double)k0 * P.dz - zslack, zb = zfirst + (double)k1 * P.dz + zslack;
    S.smid[tid] = (float)(0.5 * (za + zb) - zc);
    const float e = 1.0e-6f * (float)(fabs(za - zc) + fabs(zb - zc) + (zb - za));
    S.shalf_up[tid] = (float)(0.5 * (zb - za)) * 1.0001f + e + 1e-30f;
    S.shalf_dn[tid] = fmaxf((float)(0.5 * (zb - za)) * 0.9999f - e, 0.f);
  }
  for (int kz = tid; kz < P.wz; kz += nthr) S.fz[kz] = (float)((P.zmin + (double)(P.iz0 + kz - 1) * P.dz) - zc);
  int cur_model = -1;
  bool vec_ok = false;
  unsigned ph = 0;

  for (int t = blockIdx.x; t < G.total_tiles; t += gridDim.x) {
    const int model = P0.models ? k1t_div(t, G.dv_tpm) : 0;
    const int tile = t - model * G.tiles_per_model;
    if (model != cur_model) {
      cur_model = model;
      if (P0.models) { // batch form
        const K1Model M = P0.models[model];
        m_nodes = P0.nodes + M.node_off;
        m_rpts = P0.rpts + 3 * M.pt_off;
        m_ind = P0.ind + M.pt_off;
        m_params = P0.params + 3 * M.pt_off;
        m_root = M.root;
        m_n = M.n;
        const long long o = (long long)model * P0.model_stride;
        m_vp = P0.vp + o; m_vs = P0.vs + o; m_rho = P0.rho + o; m_sites = P0.sites + o;
      }
      // 16-byte stores need all four arrays in the same phase (true for any sane allocation; else scalar stores)
      ph = (unsigned)(((uintptr_t)m_vp >> 3) & 1);
      vec_ok = (((uintptr_t)m_vs >> 3) & 1) == ph && (((uintptr_t)m_rho >> 3) & 1) == ph && (((uintptr_t)m_sites >> 2) & 1) == ph &&
               ((uintptr_t)m_vp & 7) == 0 && ((uintptr_t)m_vs & 7) == 0 && ((uintptr_t)m_rho & 7) == 0 && ((uintptr_t)m_sites & 3) == 0;
    }
    const int txi = k1t_div(tile, G.dv_tiles_y);
    const int tx0 = txi * ttx, ty0 = (tile - txi * G.tiles_y) * tty;
    const int tw = min(ttx, P.wx - tx0), th = min(tty, P.wy - ty0);
    // ---- 0. box lists (as k1_tile.cuh; the staged records are float32) ------------------------------------------
    __syncthreads();
    if (tid < K1T_MAXSEG) { S.useg[tid] = 0x7f000000ull << 32; S.cnt[tid] = 0; }
    if (tid == 0) { S.tcnt = 0; S.ncnt = 0; S.phimin = 0x7f000000u; S.bmax = 0u; }
    const double rx0 = P.xmin + (double)(P.ix0 + tx0 - 1) * P.dx, rx1 = P.xmin + (double)(P.ix0 + tx0 + tw - 2) * P.dx;
    const double ry0 = P.ymin + (double)(P.iy0 + ty0 - 1) * P.dy, ry1 = P.ymin + (double)(P.iy0 + ty0 + th - 2) * P.dy;
    const double cx = 0.5 * (rx0 + rx1), cy = 0.5 * (ry0 + ry1);
    if (tid < tw) S.fx[tid] = (float)((P.xmin + (double)(P.ix0 + tx0 + tid - 1) * P.dx) - cx);
    else if (tid >= 64 && tid - 64 < th) S.fy[tid - 64] = (float)((P.ymin + (double)(P.iy0 + ty0 + (tid - 64) - 1) * P.dy) - cy);
    __syncthreads();
    const float hxf = (float)(0.5 * (rx1 - rx0)) * 1.0001f + 1e-30f, hyf = (float)(0.5 * (ry1 - ry0)) * 1.0001f + 1e-30f;
    const float hxl = (float)(0.5 * (rx1 - rx0)) * 0.9999f, hyl = (float)(0.5 * (ry1 - ry0)) * 0.9999f;
    // a. the smallest "farthest corner" distance in the plane
    {
      float pm = 3.0e38f;
      for (int n = tid; n < m_n; n += nthr) {
        const float ax = fabsf((float)(__ldg(&m_rpts[3 * n + 0]) - cx)) + hxf, ay = fabsf((float)(__ldg(&m_rpts[3 * n + 1]) - cy)) + hyf;
        pm = fminf(pm, ax * ax + ay * ay);
      }
      const unsigned hmin = __reduce_min_sync(0xffffffffu, __float_as_uint(pm));
      if (lane == 0) atomicMin(&S.phimin, hmin);
    }
    __syncthreads();
    // a'. the nuclei near the tile in the plane
    const float r1 = sqrtf(__uint_as_float(S.phimin)) * 1.0001f + G.near_r;
    const float phi_cut = r1 * r1;
    for (int n0 = 0; n0 < m_n; n0 += nthr) {
      const int n = n0 + tid;
      bool near = false;
      float rx = 0.f, ry = 0.f;
      if (n < m_n) {
        rx = (float)(__ldg(&m_rpts[3 * n + 0]) - cx); ry = (float)(__ldg(&m_rpts[3 * n + 1]) - cy);
        const float lx = fmaxf(fabsf(rx) - hxl, 0.f), ly = fmaxf(fabsf(ry) - hyl, 0.f);
        near = (lx * lx + ly * ly) * 0.9999f <= phi_cut;
      }
      const unsigned m = __ballot_sync(0xffffffffu, near);
      if (m == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&S.ncnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int pos = base + __popc(m & lt);
      if (near && pos < K1B_MAXN) S.near[pos] = make_float4(rx, ry, (float)(__ldg(&m_rpts[3 * n + 2]) - zc), __int_as_float(n));
    }
    __syncthreads();
    const int N = S.ncnt;
    const bool near_ok = N <= K1B_MAXN;
    // b. U_B per segment
    if (near_ok) {
      const int split = nseg <= 16 ? 2 : 1;
      const int sub = split == 2 ? (lane >> 4) : 0;
      const int wib = (tid >> 5) * split + sub, nw = (nthr >> 5) * split;
      for (int s = (split == 2 ? (lane & 15) : lane); s < nseg; s += 32) {
        const float mid = S.smid[s], hu = S.shalf_up[s];
        float um = 3.0e38f;
        int ue = 0;
        for (int e = wib; e < N; e += nw) {
          const float4 r = S.near[e];
          const float ax = fabsf(r.x) + hxf, ay = fabsf(r.y) + hyf;
          const float phi = (ax * ax + ay * ay) * 1.0001f;
          const float dh = fabsf(r.z - mid) + hu;
          const float v = fmaf(dh, dh, phi);
          if (v < um) { um = v; ue = e; }
        }
        if (um < 1.0e38f) atomicMin(&S.useg[s], ((unsigned long long)__float_as_uint(um) << 32) | (unsigned)ue);
      }
    }
    __syncthreads();
    // c. the filter
    {
      float umax = 0.f;
      for (int s = 0; s < nseg; ++s) umax = fmaxf(umax, __uint_as_float((unsigned)(S.useg[s] >> 32)) * 1.0002f);
      const bool from_list = near_ok && umax <= phi_cut;
      const int total = from_list ? N : m_n;
      float bm = 0.f;
      for (int e0 = 0; e0 < total; e0 += nthr) {
        const int e = e0 + tid;
        unsigned mask = 0;
        int n = 0;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (e < total) {
          if (from_list) { const float4 r = S.near[e]; rx = r.x; ry = r.y; rz = r.z; n = __float_as_int(r.w); }
          else {
            n = e;
            rx = (float)(__ldg(&m_rpts[3 * n + 0]) - cx); ry = (float)(__ldg(&m_rpts[3 * n + 1]) - cy); rz = (float)(__ldg(&m_rpts[3 * n + 2]) - zc);
          }
          const float lx = fmaxf(fabsf(rx) - hxl, 0.f), ly = fmaxf(fabsf(ry) - hyl, 0.f);
          const float plo = (lx * lx + ly * ly) * 0.9999f;
          if (plo <= umax) {
            const float ni = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
            for (int s = 0; s < nseg; ++s) {
              const unsigned long long us = S.useg[s];
              const float mid = S.smid[s];
              const float dl = fmaxf(fabsf(rz - mid) - S.shalf_dn[s], 0.f);
              if (fmaf(dl * 0.9999f, dl, plo) <= __uint_as_float((unsigned)(us >> 32)) * 1.0002f) {
                // Dominance: J, the nucleus that attains U_B, is strictly nearer than this one at EVERY point q of the box
                // iff max_q (|q-J|^2 - |q-x|^2) < 0; the difference is linear in q (2 q.(x-J) + |J|^2 - |x|^2), so its
                // maximum over the box is taken at a corner: 2 (mid dz + |dx| hx + |dy| hy + |dz| hz) + |J|^2 - |x|^2.
                // Float32 on tile-relative coordinates, half extents rounded up, and a margin of 3e-5 of the magnitudes
                // involved (evaluation error < 1e-6 of them): only nuclei that lose everywhere by that margin are dropped.
                bool keep = true;
                if (near_ok) {
                  const float4 J = S.near[(unsigned)us];
                  const float dx = rx - J.x, dy = ry - J.y, dz = rz - J.z;
                  const float hu = S.shalf_up[s];
                  const float nJ = fmaf(J.z, J.z, fmaf(J.y, J.y, J.x * J.x));
                  const float lin = fmaf(mid, dz, fmaf(fabsf(dx), hxf, fmaf(fabsf(dy), hyf, fabsf(dz) * hu)));
                  const float val = fmaf(2.f, lin, nJ - ni);
                  const float scale = ni + nJ + fmaf(mid, mid, fmaf(hu, hu, fmaf(hyf, hyf, hxf * hxf)));
                  keep = !(val < -3.0e-5f * scale);
                }
                if (keep) mask |= 1u << s;
              }
            }
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, mask != 0);
        if (m == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&S.tcnt, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(m & lt);
        if (mask != 0 && pos < K1B_MAXT) {
          S.rec[pos] = make_float4(rx, ry, rz, __int_as_float(__ldg(&m_ind[n])));
          bm = fmaxf(bm, fmaxf(fabsf(rx), fmaxf(fabsf(ry), fabsf(rz))));
          while (mask) {
            const int s = __ffs(mask) - 1;
            mask &= mask - 1;
            const int q = atomicAdd(&S.cnt[s], 1);
            if (q < K1T_L) S.list[s][q] = (unsigned short)pos;
          }
        }
      }
      const unsigned bw = __reduce_max_sync(0xffffffffu, __float_as_uint(bm));
      if (lane == 0 && bw) atomicMax(&S.bmax, bw);
    }
    __syncthreads();
    const bool tile_ok = S.tcnt <= K1B_MAXT; // else: every node of the tile takes the exact tree walk
    // per-coordinate magnitude bound of everything the float32 walk subtracts (candidates and the tile's own nodes)
    const float Bm = fmaxf(fmaxf(__uint_as_float(S.bmax), hzf), fmaxf(hxf, hyf)) * 1.0001f;
    const float B2 = Bm * Bm;
    // ---- 1. nodes: items ordered (unit group = segment, column, unit inside the segment) --------------------------
    const int nitems = G.ngroups * G.per_group;
    for (int item = tid; item < nitems; item += nthr) {
      const int gsel = k1t_div(item, G.dv_group), rem = item - gsel * G.per_group;
      const int c = k1t_div(rem, G.dv_ups), unit = gsel * G.ups + (rem - c * G.ups);
      const int ci = k1t_div(c, G.dv_ty), cj = c - ci * tty;
      if (unit >= G.nunits || ci >= tw || cj >= th) continue;
      const int i = P.ix0 + tx0 + ci, j = P.iy0 + ty0 + cj;
      const size_t obase = ((size_t)((i - P.ia0) * P.ny_a + (j - P.ja0))) * (size_t)P.nz_a + (size_t)(P.iz0 - P.ka0);
      const int par = (int)((obase + ph) & 1); // pairs start where the absolute address is a multiple of 16
      const int e0 = NPT * unit - par;         // elements e0 .. e0+NPT-1 of the column's window (0-based)
      int idx[NPT];
      bool on[NPT];
      bool all_on = true;
#pragma unroll
      for (int h = 0; h < NPT; ++h) {
        const int kz = e0 + h;
        on[h] = kz >= 0 && kz < P.wz;
        if (on[h] && P.use_pm) { // mcmc_loc2.f90:2055-2056: only nodes still carrying the moved cell's old values
          const size_t o = obase + (size_t)kz;
          if (!(fabs(m_vs[o] - P.pm_vs) < P.pm_eps && fabs(m_vp[o] - P.pm_vp) < P.pm_eps)) on[h] = false;
        }
        all_on = all_on && on[h];
        idx[h] = 0;
      }
      const float fx = S.fx[ci], fy = S.fy[cj];
      bool exact[NPT];
      int pos[NPT];
#pragma unroll
      for (int h = 0; h < NPT; ++h) { exact[h] = false; pos[h] = 0; }
      const int sg0 = all_on ? k1t_div(e0, G.dv_seg) : 0;
      if (all_on && sg0 == k1t_div(e0 + NPT - 1, G.dv_seg)) {
        // all nodes of the unit in one box (the usual case): one walk over its list
        const int cnt = S.cnt[sg0];
        if (!tile_ok || cnt > K1T_L || cnt == 0) {
#pragma unroll
          for (int h = 0; h < NPT; ++h) exact[h] = true;
        } else {
          float m1[NPT], m2[NPT], fz[NPT];
#pragma unroll
          for (int h = 0; h < NPT; ++h) { m1[h] = 3.0e38f; m2[h] = 3.0e38f; fz[h] = S.fz[e0 + h]; }
          const unsigned short* L = S.list[sg0];
          for (int q = 0; q < cnt; ++q) {
            const int e = L[q];
            const float4 r = S.rec[e];
            const float dx = r.x - fx, dy = r.y - fy;
            const float p = fmaf(dy, dy, dx * dx);
#pragma unroll
            for (int h = 0; h < NPT; ++h) {
              const float dz = r.z - fz[h];
              const float s = fmaf(dz, dz, p);
              pos[h] = s < m1[h] ? e : pos[h];
              m2[h] = fminf(m2[h], fmaxf(m1[h], s));
              m1[h] = fminf(m1[h], s);
            }
          }
#pragma unroll
          for (int h = 0; h < NPT; ++h) exact[h] = !(m2[h] - m1[h] > 9.5367431640625e-7f * (B2 + m1[h] + m2[h]));
        }
      } else {
#pragma unroll
        for (int h = 0; h < NPT; ++h) {
          if (!on[h]) continue;
          const int sg = k1t_div(e0 + h, G.dv_seg);
          const int cnt = S.cnt[sg];
          if (!tile_ok || cnt > K1T_L || cnt == 0) { exact[h] = true; continue; }
          float m1, m2;
          k1b_scan1(S, sg, cnt, fx, fy, S.fz[e0 + h], m1, m2, pos[h]);
          exact[h] = !(m2 - m1 > 9.5367431640625e-7f * (B2 + m1 + m2));
        }
      }
      double pvp[NPT], pvs[NPT], prho[NPT];
#pragma unroll
      for (int h = 0; h < NPT; ++h) {
        if (!on[h]) continue;
        if (exact[h]) idx[h] = k1b_walk(P0, m_nodes, m_rpts, m_ind, m_root, m_n, i, j, e0 + h);
        else idx[h] = __float_as_int(S.rec[pos[h]].w);
        const double* pr = m_params + 3 * (size_t)(idx[h] - 1);
        pvp[h] = __ldg(&pr[0]); pvs[h] = __ldg(&pr[1]); prho[h] = __ldg(&pr[2]);
      }
#pragma unroll
      for (int hp = 0; hp < NPT; hp += 2) {
        const size_t o0 = obase + (size_t)(e0 + hp < 0 ? 0 : e0 + hp);
        if (on[hp] && on[hp + 1] && vec_ok) { // 16-byte stores (8 for sites_id); aligned by construction of `par`
          *reinterpret_cast<int2*>(m_sites + o0) = make_int2(idx[hp], idx[hp + 1]);
          *reinterpret_cast<double2*>(m_vp + o0) = make_double2(pvp[hp], pvp[hp + 1]);
          *reinterpret_cast<double2*>(m_vs + o0) = make_double2(pvs[hp], pvs[hp + 1]);
          *reinterpret_cast<double2*>(m_rho + o0) = make_double2(prho[hp], prho[hp + 1]);
        } else {
#pragma unroll
          for (int h = hp; h < hp + 2; ++h) {
            if (!on[h]) continue;
            const size_t o = obase + (size_t)(e0 + h);
            m_sites[o] = idx[h]; m_vp[o] = pvp[h]; m_vs[o] = pvs[h]; m_rho[o] = prho[h];
          }
        }
      }
    }
  }
}
