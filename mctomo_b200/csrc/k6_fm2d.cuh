// k6_fm2d.cuh -- SURVEY 8(f)2: the reference's 2-D fast-marching eikonal solver on the device: `modrays` as surf_likelihood
// drives it (src/likelihood_surf.F90:295-336): gridder, bsplrefine, the source loop with source-grid refinement
// (fm2d/fm2dray_cartesian.f90:67-478,490-668), srtimes (:676-770), rpaths with cfd = 0 (:773-1456; uar = 0, group-velocity
// data), travel / fouds1 / fouds2 / addtree / downtree / updtree / bilinear (fm2d/fm2d_ttime.f90); and the curved-ray
// branch of surf_likelihood on a session's resident maps (mct_session_likelihood_fm2d).
//
// The unit of parallelism is the reference's own: every (period, source) pair is an independent eikonal problem
// (the Fortran loops over sources inside an OpenMP loop over periods).  One block of two warps per problem, all of them in
// one launch.  Fast marching accepts nodes strictly in the order of a binary heap, and the reference's travel times depend
// on that order (which neighbours are alive when a node is updated), so the march keeps the Fortran's heap, stencils and
// operation order; inside one accepted node the (up to) 4 x 4 stencil quadrants are solved by 16 lanes of the stencil warp
// while lane 0 of the heap warp re-orders the heap (fm_travel / fm_stencil_warp, two named barriers per node).  What is
// data-parallel inside a problem -- B-spline velocities of the propagation grid and of the refined source grid, the
// narrow-band completion sweep, the receiver interpolation, the ray tracing of rpaths (one receiver per lane) -- is spread
// over the lanes of warp 0.  Results are bit-identical to oracle/fm2d_ref.c: receiver times, the whole field, node /
// stencil counts, every ray point.  Throughput comes from the number of problems in flight (np x nsrc: 88 in example1,
// thousands in the multi-mode configurations); one problem's latency is the heap warp's dependent instructions.
#pragma once

struct FmParams {
  int nmaps, nsrc, nrc;
  const double *scx, *scz, *rcx, *rcz;
  const int32_t* srs;        // (nrc, nsrc, nmaps): raystat(:,1,period) as the Fortran holds it
  int nvx, nvz;
  double gox, goz, dvx, dvz;
  const double* velv;        // like%vel with its replicated edge: element (a, b) of map m at velv[m*vel_ms + (b*(nvz+2) + a)*vel_es]
  long long vel_es, vel_ms;  // (1, (nvz+2)*(nvx+2)) for maps stored one after the other; (nmaps, 1) for the Fortran's like%vel(np, ny+2, nx+2)
  long long srs_ms;          // map stride of srs: nrc*nsrc, or 2*nrc*nsrc for dat%raystat(nrev*nsrc, 2, np)
  int gdx, gdz, asgr, sgdl, sgs, fom;
  double snb;
  int nnx, nnz, ldr;         // propagation grid; leading dimension of the refined arrays
  double* veln;              // (nnz, nnx, nmaps): gridder's output (fm2d_gridder_kernel)
  double* slow;              // (nnz, nnx, nmaps): 1 / veln, the slowness every stencil of a node starts from (same division, done once)
  char* scratch; size_t scratch_per_problem;
  size_t heap_bytes;         // bytes reserved per problem for the heap entries
  double* ttime;             // (nrc, nsrc, nmaps); entries without data untouched
  double* field;             // optional (nnz, nnx, nsrc, nmaps): ttn of every problem
  int32_t* err;              // per problem: 0, 1 source outside, 2 narrow band overflow, 3 receiver outside
  unsigned long long* counters; // [0] nodes accepted, [1] stencil updates
  // ray geometry (uar = 0, group-velocity data): null ray_npts = travel times only
  const int32_t* srsv; long long srsv_ms; // raystat(:,2,period): the 1-based slot of every pair's ray
  int ray_cap;                 // points per slot
  int32_t* ray_npts;           // (nrc*nsrc, nmaps), zeroed by the caller
  double* ray_pts;             // (2, ray_cap, nrc*nsrc, nmaps)
  double* ray_len;             // (nrc*nsrc, nmaps): T_RAY%length
  int32_t* crazy;              // per problem: rays that ran out of points (crazyrp)
};

__device__ __forceinline__ void fm_bspl(double u, double w[5]) {
  const double um = 1.0 - u;
  w[1] = ((um * um) * um) / 6.0;
  w[2] = (4.0 - 6.0 * (u * u) + 3.0 * ((u * u) * u)) / 6.0;
  w[3] = (1.0 + 3.0 * u + 3.0 * (u * u) - 3.0 * ((u * u) * u)) / 6.0;
  w[4] = ((u * u) * u) / 6.0;
}

// gridder (fm2dray_cartesian.f90:490-590): one thread per propagation node, the Fortran's (i,j,l,m) recovered from it
__global__ void fm2d_gridder_kernel(const __grid_constant__ FmParams P) {
  const long long total = (long long)P.nnx * P.nnz * P.nmaps;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int map = (int)(t / ((long long)P.nnx * P.nnz));
    const int r = (int)(t - (long long)map * P.nnx * P.nnz);
    const int stx = r / P.nnz + 1, stz = r % P.nnz + 1;
    int i = (stz - 1) / P.gdz + 1; if (i > P.nvz - 1) i = P.nvz - 1;
    int j = (stx - 1) / P.gdx + 1; if (j > P.nvx - 1) j = P.nvx - 1;
    const int l = stz - P.gdz * (i - 1), m = stx - P.gdx * (j - 1);
    double ui[5], vi[5], u;
    u = P.gdx; u = (m - 1) / u; fm_bspl(u, ui);
    u = P.gdz; u = (l - 1) / u; fm_bspl(u, vi);
    const double* velv = P.velv + (size_t)map * P.vel_ms;
    double sumi = 0.0;
    for (int i1 = 1; i1 <= 4; ++i1) {
      double sumj = 0.0;
      for (int j1 = 1; j1 <= 4; ++j1) sumj = sumj + ui[j1] * velv[((size_t)(j - 2 + j1) * (P.nvz + 2) + (i - 2 + i1)) * P.vel_es];
      sumi = sumi + vi[i1] * sumj;
    }
    P.veln[(size_t)map * P.nnx * P.nnz + (size_t)(stx - 1) * P.nnz + (stz - 1)] = sumi;
    P.slow[(size_t)map * P.nnx * P.nnz + (size_t)(stx - 1) * P.nnz + (stz - 1)] = 1.0 / sumi;
  }
}

// one marching grid: arrays with leading dimension ld (rows = z), 1-based accessors
struct __align__(16) FmEnt { double key; int32_t pos; int32_t pad; }; // one heap entry: the node's travel time beside its linear index
struct FmGrid {
  int nnx, nnz, ld;
  unsigned ld_magic; // ceil(2^32 / ld): a linear node index is split into (ix, iz) without an integer division
  double gox, goz, dnx, dnz;
  const double* veln;
  const double* slow;
  double* ttn;
  int32_t* nsts;
  FmEnt* heap; // 1-based entries
  int ntr, maxbt, fom;
  int vnl, vnr, vnt, vnb;
  int error;
  unsigned n_accept, n_update;
};
#define FVELN(G, k, j) (G).veln[(size_t)((j) - 1) * (G).ld + ((k) - 1)]
#define FTTN(G, k, j) (G).ttn[(size_t)((j) - 1) * (G).ld + ((k) - 1)]
#define FNSTS(G, k, j) (G).nsts[(size_t)((j) - 1) * (G).ld + ((k) - 1)]

// commands of the heap warp to the stencil warp (>= 0: the linear index of the node just accepted)
#define FM_CMD_GRID (-2)
#define FM_CMD_QUIT (-3)
// (barrier.sync, not bar.sync = barrier.sync.aligned: lane 0 of the heap warp arrives from its own branch)
__device__ __forceinline__ void fm_bar_a() { asm volatile("barrier.sync 1, 64;" ::: "memory"); }
__device__ __forceinline__ void fm_bar_b() { asm volatile("barrier.sync 2, 64;" ::: "memory"); }
__device__ __forceinline__ int fm_row(int h, int ld, unsigned magic) { // h / ld for 0 <= h < 2^31, ld >= 2
  int q = (int)__umulhi((unsigned)h, magic);
  if (q * ld > h) q--;
  return q;
}

// The narrow-band heap (addtree / updtree / downtree of fm2d_ttime.f90), lane 0 of the heap warp only.  Same comparisons in
// the same order as the Fortran, on a copy of each entry's travel time stored beside the entry (one 16-byte load per entry:
// the sift loops never chase ttn, and a level costs one load latency); sizes live in registers.
struct FmHeap {
  FmEnt* ent; int32_t* nsts; const double* ttn;
  int ntr, maxbt;
};
__device__ __forceinline__ FmEnt fm_ld(const FmEnt* p) {
  const int4 v = *reinterpret_cast<const int4*>(p);
  FmEnt e; e.key = __hiloint2double(v.y, v.x); e.pos = v.z; e.pad = 0;
  return e;
}
__device__ __forceinline__ void fm_st(FmEnt* p, double key, int pos) {
  int4 v; v.x = __double2loint(key); v.y = __double2hiint(key); v.z = pos; v.w = 0;
  *reinterpret_cast<int4*>(p) = v;
}
__device__ __forceinline__ void fm_sift_up(FmHeap& H, int self, double t, int tpc) {
  int tpp = tpc >> 1;
  while (tpp > 0) {
    const FmEnt p = fm_ld(H.ent + tpp);
    if (t < p.key) {
      H.nsts[p.pos] = tpc;
      fm_st(H.ent + tpc, p.key, p.pos);
      tpc = tpp;
      tpp = tpc >> 1;
    } else tpp = 0;
  }
  fm_st(H.ent + tpc, t, self);
  H.nsts[self] = tpc;
}
__device__ __forceinline__ bool fm_addtree(FmHeap& H, int a, double t) { // a: linear node index; false: the narrow band is full
  if (H.ntr + 1 > H.maxbt) return false;
  H.ntr++;
  fm_sift_up(H, a, t, H.ntr);
  return true;
}
__device__ __forceinline__ void fm_downtree(FmHeap& H) {
  if (H.ntr == 1) { H.ntr--; return; }
  const FmEnt last = fm_ld(H.ent + H.ntr);
  H.ntr--;
  const int ntr = H.ntr;
  int tpp = 1, tpc = 2;
  // the entry taken from the end sinks from the root: at each level the smaller child (the left one on a tie) moves up
  // while it is smaller than the sinking entry -- the Fortran's swaps, without writing the sinking entry at every level
  while (tpc < ntr) {
    FmEnt c = fm_ld(H.ent + tpc);
    const FmEnt r = fm_ld(H.ent + tpc + 1);
    if (c.key > r.key) { tpc = tpc + 1; c = r; }
    if (c.key < last.key) {
      H.nsts[c.pos] = tpp;
      fm_st(H.ent + tpp, c.key, c.pos);
      tpp = tpc;
      tpc = 2 * tpp;
    } else tpc = ntr + 1;
  }
  if (tpc == ntr) {
    const FmEnt c = fm_ld(H.ent + tpc);
    if (c.key < last.key) {
      H.nsts[c.pos] = tpp;
      fm_st(H.ent + tpp, c.key, c.pos);
      tpp = tpc;
    }
  }
  fm_st(H.ent + tpp, last.key, last.pos);
  H.nsts[last.pos] = tpp;
}
__device__ __forceinline__ double fm_qsolve(double a, double b, double c) {
  double rd1 = b * b - 4.0 * a * c;
  if (rd1 < 0.0) rd1 = 0.0;
  return (-b + sqrt(rd1)) / (2.0 * a);
}
// t / 3.0, correctly rounded, without computing a reciprocal: with y = RN(1/3), q = RN(t y) is within one ulp of t/3, the
// residual r = t - 3 q is exact in one fma, and RN(q + r y) is the correctly rounded quotient (Markstein's final step -- the
// same one the compiler's own division sequence ends with).  Outside the range where r cannot underflow: the division itself.
__device__ __forceinline__ double fm_div3(double t) {
  const double y = 1.0 / 3.0;
  const double m = fabs(t);
  if (m < 1.0e300 && m > 1.0e-280) {
    const double q = t * y;
    const double r = fma(-3.0, q, t);
    return fma(r, y, q);
  }
  return t / 3.0;
}
// what a stencil needs of the grid, in registers of the stencil warp
struct FmSten {
  int nnx, nnz, ld, fom;
  unsigned ld_magic;
  double dnx, dnz;
  const double* slow; double* ttn; const int32_t* nsts;
};
// ONE quadrant (the neighbour pair j = ix + dj, k = iz + dk) of fouds1 / fouds2 (fm2d_ttime.f90:138-197, 199-345) for the node
// (iz, ix) = linear index a: the candidate travel time of that stencil, or false when it has no solution.  The Fortran takes
// the minimum of the (up to) four candidates in the order (j, k) = (-,-), (-,+), (+,-), (+,+); a minimum does not depend on
// the order.
__device__ __forceinline__ bool fm_quadrant(const FmSten& G, int a, int iz, int ix, int dj, int dk, double* trav_out) {
  const int j = ix + dj, k = iz + dk;
  if (j < 1 || j > G.nnx || k < 1 || k > G.nnz) return false;
  const int aj = a + dj * G.ld, ak = a + dk; // (iz, j) and (k, ix)
  const double slown = G.slow[a], dnx = G.dnx, dnz = G.dnz;
  const int sj = G.nsts[aj], sk = G.nsts[ak];
  double a2 = 0, b = 0, c = 0, tref = 0, u, v, em;
  bool third = false; // the Fortran's final division by 3 (second-order operators)
  int swsol = 0;
  if (G.fom == 0) {
    if (sj == 0) {
      swsol = 1;
      const double tj = G.ttn[aj];
      if (sk == 0) {
        u = dnx; v = dnz; em = G.ttn[ak] - tj;
        a2 = u * u + v * v;
        b = -2.0 * (u * u) * em;
        c = (u * u) * (em * em - (v * v) * (slown * slown));
        tref = tj;
      } else { a2 = 1.0; b = 0.0; c = -(slown * slown) * (dnx * dnx); tref = tj; }
    } else if (sk == 0) {
      swsol = 1;
      const double sd = slown * dnz;
      a2 = 1.0; b = 0.0; c = -(sd * sd); tref = G.ttn[ak];
    }
  } else {
    int swj = -1, swk = -1;
    const int j2 = j + dj, k2 = k + dk;
    const int aj2 = aj + dj * G.ld, ak2 = ak + dk;
    if (j2 >= 1 && j2 <= G.nnx) { if (G.nsts[aj2] == 0) swj = 0; }
    const double tj = G.ttn[aj];
    double tj2 = 0;
    if (sj == 0 && swj == 0) { swj = -1; tj2 = G.ttn[aj2]; if (tj > tj2) swj = 0; }
    else swj = -1;
    if (k2 >= 1 && k2 <= G.nnz) { if (G.nsts[ak2] == 0) swk = 0; }
    const double tk = G.ttn[ak];
    double tk2 = 0;
    if (sk == 0 && swk == 0) { swk = -1; tk2 = G.ttn[ak2]; if (tk > tk2) swk = 0; }
    else swk = -1;
    if (swj == 0) {
      swsol = 1;
      if (swk == 0) {
        u = 2.0 * dnx; v = 2.0 * dnz;
        em = 4.0 * tj - tj2 - 4.0 * tk;
        em = em + tk2;
        a2 = v * v + u * u;
        b = 2.0 * em * (u * u);
        c = (u * u) * (em * em - (slown * slown) * (v * v));
        tref = 4.0 * tj - tj2;
        third = true;
      } else if (sk == 0) {
        u = dnz; v = 2.0 * dnx;
        em = 3.0 * tk - 4.0 * tj + tj2;
        a2 = v * v + 9.0 * (u * u);
        b = 6.0 * em * (u * u);
        c = (u * u) * (em * em - (slown * slown) * (v * v));
        tref = tk;
      } else {
        u = 2.0 * dnx;
        a2 = 1.0; b = 0.0; c = -(u * u) * (slown * slown);
        tref = 4.0 * tj - tj2;
        third = true;
      }
    } else if (sj == 0) {
      swsol = 1;
      if (swk == 0) {
        u = dnx; v = 2.0 * dnz;
        em = 3.0 * tj - 4.0 * tk + tk2;
        a2 = v * v + 9.0 * (u * u);
        b = 6.0 * em * (u * u);
        c = (u * u) * (em * em - (v * v) * (slown * slown));
        tref = tj;
      } else if (sk == 0) {
        u = dnx; v = dnz;
        em = tk - tj;
        a2 = u * u + v * v;
        b = -2.0 * (u * u) * em;
        c = (u * u) * (em * em - (v * v) * (slown * slown));
        tref = tj;
      } else { a2 = 1.0; b = 0.0; c = -(slown * slown) * (dnx * dnx); tref = tj; }
    } else {
      if (swk == 0) {
        swsol = 1;
        u = 2.0 * dnz;
        a2 = 1.0; b = 0.0; c = -(u * u) * (slown * slown);
        tref = 4.0 * tk - tk2;
        third = true;
      } else if (sk == 0) {
        swsol = 1;
        a2 = 1.0; b = 0.0; c = -(slown * slown) * (dnz * dnz);
        tref = tk;
      }
    }
  }
  if (!swsol) return false;
  const double t = tref + fm_qsolve(a2, b, c);
  *trav_out = third ? fm_div3(t) : t; // the divisor is 1 or 3; x / 1.0 == x
  return true;
}
__device__ __forceinline__ double fm_bilinear(const FmGrid& G, const double nv[3][3], double dsx, double dsz) {
  double biv = 0.0;
  for (int i = 1; i <= 2; ++i)
    for (int j = 1; j <= 2; ++j) {
      const double produ = (1.0 - fabs(((i - 1) * G.dnx - dsx) / G.dnx)) * (1.0 - fabs(((j - 1) * G.dnz - dsz) / G.dnz));
      biv = biv + nv[i][j] * produ;
    }
  return biv;
}
// The stencil warp of a problem (warp 1 of the block).  For every node the heap warp accepts it solves the (up to) four
// stencil quadrants of the (up to) four neighbours, one (neighbour, quadrant) per lane on lanes 0..15, takes the minimum per
// neighbour and writes the neighbour's new travel time -- on the state before the updates (a neighbour being updated is not
// alive, and the stencils only read alive nodes), WHILE the heap warp re-orders the heap after the pop (downtree only moves
// heap positions, i.e. positive status values, which no stencil distinguishes).  Barrier A: a command is ready; barrier B:
// the travel times are written.
__device__ void fm_stencil_warp(const FmGrid& G, volatile int* sh, int lane) {
  FmSten S;
  S.nnx = S.nnz = S.ld = S.fom = 0; S.ld_magic = 0; S.dnx = S.dnz = 0; S.slow = nullptr; S.ttn = nullptr; S.nsts = nullptr;
  const int n = (lane >> 2) & 3, q = lane & 3;
  const int dj = (q >> 1) ? 1 : -1, dk = (q & 1) ? 1 : -1;
  int off = 0;
  for (;;) {
    fm_bar_a();
    const int h = sh[0];
    if (h == FM_CMD_QUIT) return;
    if (h == FM_CMD_GRID) {
      S.nnx = G.nnx; S.nnz = G.nnz; S.ld = G.ld; S.fom = G.fom; S.ld_magic = G.ld_magic; S.dnx = G.dnx; S.dnz = G.dnz;
      S.slow = G.slow; S.ttn = G.ttn; S.nsts = G.nsts;
      off = n == 0 ? -S.ld : (n == 1 ? S.ld : (n == 2 ? -1 : 1));
      fm_bar_b();
      continue;
    }
    const int ix = fm_row(h, S.ld, S.ld_magic) + 1, iz = h - (ix - 1) * S.ld + 1;
    const int nix = n == 0 ? ix - 1 : (n == 1 ? ix + 1 : ix), niz = n == 2 ? iz - 1 : (n == 3 ? iz + 1 : iz);
    const int a = h + off;
    int st = 0; // the neighbour's status; 0 = nothing to do (alive or outside)
    double trav = 0;
    bool has = false;
    if (lane < 16 && nix >= 1 && nix <= S.nnx && niz >= 1 && niz <= S.nnz) st = S.nsts[a];
    if (st != 0) has = fm_quadrant(S, a, niz, nix, dj, dk, &trav);
    // minimum of the quadrants that have a solution
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const double ot = __shfl_xor_sync(0xffffffffu, trav, o);
      const bool oh = __shfl_xor_sync(0xffffffffu, has ? 1 : 0, o) != 0;
      if (oh && (!has || ot < trav)) trav = ot;
      has = has || oh;
    }
    // every stencil has read the state before any time is written: the lanes of a warp are past their loads here
    if (q == 0 && st != 0) S.ttn[a] = has ? trav : 0.0; // (no stencil solved: travm is undefined in the Fortran; cannot happen next to an alive node)
    fm_bar_b();
  }
}
// travel (fm2d_ttime.f90:27-136), the heap warp's side (warp 0; lane 0 owns the heap).  The march order is the heap's, so
// nodes are accepted one at a time; after the stencil warp has written the neighbours' times, lane 0 sifts the neighbours in
// the reference's order (x-1, x+1, z-1, z+1).  urg 0/1: nsts must already be -1 everywhere (the caller's lanes fill it).
// G lives in shared memory.
__device__ void fm_travel(FmGrid& G, double scx, double scz, int urg, int lane, volatile int* sh) {
  FmHeap H;
  H.ent = G.heap; H.nsts = G.nsts; H.ttn = G.ttn; H.ntr = 0; H.maxbt = G.maxbt;
  const int nnx = G.nnx, nnz = G.nnz, ld = G.ld;
  const unsigned magic = G.ld_magic;
  int error = 0;
  unsigned n_accept = 0, n_update = 0;
  if (lane == 0) sh[0] = FM_CMD_GRID;
  fm_bar_a();
  if (lane == 0) {
    int isx = (int)((scx - G.gox) / G.dnx) + 1;
    int isz = (int)((scz - G.goz) / G.dnz) + 1;
    if (isx < 1 || isx > nnx || isz < 1 || isz > nnz) error = 1;
    else {
      if (isx == nnx) isx--;
      if (isz == nnz) isz--;
      if (urg == 2) {
        // every node with a positive status, x outer, z inner; only the window mapped from the refined grid can hold one
        for (int i = G.vnl; i <= G.vnr && !error; ++i)
          for (int j = G.vnt; j <= G.vnb; ++j) {
            const int a = (i - 1) * ld + (j - 1);
            if (H.nsts[a] > 0) { if (!fm_addtree(H, a, H.ttn[a])) { error = 2; break; } }
          }
      } else {
        double vss[3][3];
        for (int i = 1; i <= 2; ++i) for (int j = 1; j <= 2; ++j) vss[i][j] = FVELN(G, isz - 1 + j, isx - 1 + i);
        const double dsx = (scx - G.gox) - (isx - 1) * G.dnx;
        const double dsz = (scz - G.goz) - (isz - 1) * G.dnz;
        const double vsrc = fm_bilinear(G, vss, dsx, dsz);
        for (int i = 1; i <= 2; ++i)
          for (int j = 1; j <= 2; ++j) {
            const double ex = dsx - (i - 1) * G.dnx, ez = dsz - (j - 1) * G.dnz;
            const double ds = sqrt(ex * ex + ez * ez);
            const double t0 = 2.0 * ds / (vss[i][j] + vsrc);
            const int a = (isx - 1 + i - 1) * ld + (isz - 1 + j - 1);
            G.ttn[a] = t0;
            if (!fm_addtree(H, a, t0)) error = 2;
          }
      }
    }
  }
  fm_bar_b();
  for (;;) {
    int h = -1, ix = 0, iz = 0;
    if (lane == 0 && H.ntr > 0 && !error) {
      h = H.ent[1].pos;
      ix = fm_row(h, ld, magic) + 1; iz = h - (ix - 1) * ld + 1;
      H.nsts[h] = 0;
      if (urg == 1) {
        bool swrg = false;
        if (ix == 1 && G.vnl != 1) swrg = true;
        if (ix == nnx && G.vnr != nnx) swrg = true; // (refined extent against a coarse index, as the Fortran has it)
        if (iz == 1 && G.vnt != 1) swrg = true;
        if (iz == nnz && G.vnb != nnz) swrg = true;
        if (swrg) h = -1;
      }
      if (h >= 0) sh[0] = h;
    }
    if (__shfl_sync(0xffffffffu, h, 0) < 0) break;
    fm_bar_a();
    int s[4] = {0, 0, 0, 0};
    if (lane == 0) {
      n_accept++;
      fm_downtree(H);
      // the neighbours' status, read while the stencils run: its sign is all that matters and no sift changes a sign
      if (ix > 1) s[0] = H.nsts[h - ld];
      if (ix < nnx) s[1] = H.nsts[h + ld];
      if (iz > 1) s[2] = H.nsts[h - 1];
      if (iz < nnz) s[3] = H.nsts[h + 1];
    }
    fm_bar_b();
    if (lane == 0) {
      double t[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int a = h + (m == 0 ? -ld : (m == 1 ? ld : (m == 2 ? -1 : 1)));
        t[m] = s[m] != 0 ? H.ttn[a] : 0.0;
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        if (s[m] != 0 && !error) {
          const int a = h + (m == 0 ? -ld : (m == 1 ? ld : (m == 2 ? -1 : 1)));
          n_update++;
          if (s[m] == -1) { if (!fm_addtree(H, a, t[m])) error = 2; } else fm_sift_up(H, a, t[m], H.nsts[a]);
        }
      }
    }
  }
  if (lane == 0) { G.error = error; G.n_accept = n_accept; G.n_update = n_update; G.ntr = H.ntr; }
  __syncwarp();
}

// rpaths for ONE receiver (fm2dray_cartesian.f90:773-1456, cfd = 0): see oracle/fm2d_ref.c for the quirks that are kept.
struct FmRayCtx {
  int asgr, nnx, nnz, nnxr, nnzr, ldr;
  double gox, goz, dnx, dnz, goxr, gozr, dnxr, dnzr;
  const double* ttn;     // coarse (nnz, nnx)
  const double* ttnr;    // refined, leading dimension ldr
  const int32_t* nstsr;
};
__device__ int fm_trace_ray(const FmRayCtx& C, double scx, double scz, double rx, double rz, int cap, double* pts /* [cap][2] */, int* crazy,
                            int* err) {
#define RTTN(k, j) C.ttn[(size_t)((j) - 1) * C.nnz + ((k) - 1)]
#define RTTNR(k, j) C.ttnr[(size_t)((j) - 1) * C.ldr + ((k) - 1)]
#define RNSTSR(k, j) C.nstsr[(size_t)((j) - 1) * C.ldr + ((k) - 1)]
  int isx, isz;
  if (C.asgr == 1) { isx = (int)floor((scx - C.goxr) / C.dnxr) + 1; isz = (int)floor((scz - C.gozr) / C.dnzr) + 1; }
  else { isx = (int)floor((scx - C.gox) / C.dnx) + 1; isz = (int)floor((scz - C.goz) / C.dnz) + 1; }
  double dpl = C.dnx, rd1 = C.dnz;
  if (rd1 < dpl) dpl = rd1;
  dpl = 0.5 * dpl;
  int ipx = (int)floor((rx - C.gox) / C.dnx) + 1, ipz = (int)floor((rz - C.goz) / C.dnz) + 1;
  if (ipx < 1 || ipx >= C.nnx || ipz < 1 || ipz >= C.nnz) { *err = 3; return 0; }
  int ipxr = 0, ipzr = 0, igref = 0, sw = 0, nrp = 1;
  double cx = rx, cz = rz; // rgx(j), rgz(j)
  pts[0] = cx; pts[1] = cz;
  double sred = (scx - cx) * (scx - cx);
  sred = sred + (scz - cz) * (scz - cz);
  sred = sqrt(sred);
  if (sred < 2.0 * dpl) { pts[2] = scx; pts[3] = scz; nrp = 2; sw = 1; }
#define FM_IGREF(px, pz) do { \
    ipxr = (int)floor(((px) - C.goxr) / C.dnxr) + 1; ipzr = (int)floor(((pz) - C.gozr) / C.dnzr) + 1; igref = 1; \
    if (ipxr < 1 || ipxr >= C.nnxr) igref = 0; \
    if (ipzr < 1 || ipzr >= C.nnzr) igref = 0; \
    if (igref == 1) { \
      if (RNSTSR(ipzr, ipxr) != 0 || RNSTSR(ipzr + 1, ipxr) != 0) igref = 0; \
      if (RNSTSR(ipzr, ipxr + 1) != 0 || RNSTSR(ipzr + 1, ipxr + 1) != 0) igref = 0; \
    } } while (0)
  if (C.asgr == 1) FM_IGREF(rx, rz);
  if (sw == 0) {
    if (C.asgr == 1) { if (igref == 1 && ipxr == isx && ipzr == isz) { pts[2] = scx; pts[3] = scz; nrp = 2; sw = 1; } }
    else if (ipx == isx && ipz == isz) { pts[2] = scx; pts[3] = scz; nrp = 2; sw = 1; }
  }
  const int maxrp = C.nnx * C.nnz;
  for (int j = 1; j <= maxrp; ++j) {
    if (sw == 1) break;
    double dtx, dtz;
    if (igref == 1) {
      if (ipxr == 1) { dtx = RTTNR(ipzr, ipxr + 1) - RTTNR(ipzr, ipxr); dtx = dtx / C.dnxr; }
      else if (ipxr == C.nnxr) { dtx = RTTNR(ipzr, ipxr) - RTTNR(ipzr, ipxr - 1); dtx = dtx / C.dnxr; }
      else { dtx = RTTNR(ipzr, ipxr + 1) - RTTNR(ipzr, ipxr - 1); dtx = dtx / (2.0 * C.dnxr); }
      if (ipzr == 1) { dtz = RTTNR(ipzr + 1, ipxr) - RTTNR(ipzr, ipxr); dtz = dtz / C.dnzr; }
      else if (ipzr == C.nnzr) { dtz = RTTNR(ipzr, ipxr) - RTTNR(ipzr - 1, ipxr); dtz = dtz / C.dnzr; }
      else { dtz = RTTNR(ipzr + 1, ipxr) - RTTNR(ipzr - 1, ipxr); dtz = dtz / (2.0 * C.dnzr); }
    } else {
      if (ipx == 1) { dtx = RTTN(ipz, ipx + 1) - RTTN(ipz, ipx); dtx = dtx / C.dnx; }
      else if (ipx == C.nnx) { dtx = RTTN(ipz, ipx) - RTTN(ipz, ipx - 1); dtx = dtx / C.dnx; }
      else { dtx = RTTN(ipz, ipx + 1) - RTTN(ipz, ipx - 1); dtx = dtx / (2.0 * C.dnx); }
      if (ipz == 1) { dtz = RTTN(ipz + 1, ipx) - RTTN(ipz, ipx); dtz = dtz / C.dnz; }
      else if (C.asgr == 1 && ipzr == C.nnzr) { dtz = RTTN(ipz, ipx) - RTTN(ipz - 1, ipx); dtz = dtz / C.dnz; } // sic: the refined indices (:1092)
      else { dtz = RTTN(ipz + 1, ipx) - RTTN(ipz - 1, ipx); dtz = dtz / (2.0 * C.dnz); }
    }
    if (j + 2 > cap) { (*crazy)++; sw = 1; break; }
    rd1 = sqrt(dtx * dtx + dtz * dtz);
    if (!(rd1 > 0.0)) { (*crazy)++; nrp = 1; sw = 1; break; }
    double nx_ = cx - dpl * dtx / rd1, nz_ = cz - dpl * dtz / rd1; // rgx(j+1), rgz(j+1)
    if (C.asgr == 1) FM_IGREF(nx_, nz_); else igref = 0;
    ipx = (int)floor((nx_ - C.gox) / C.dnx) + 1;
    ipz = (int)floor((nz_ - C.goz) / C.dnz) + 1;
    sred = (scx - nx_) * (scx - nx_);
    sred = sred + (scz - nz_) * (scz - nz_);
    sred = sqrt(sred);
    sw = 0;
    bool done = false;
    if (sred < 2.0 * dpl) done = true;
    else if (C.asgr == 1) { if (igref == 1 && ipxr == isx && ipzr == isz) done = true; }
    else if (ipx == isx && ipz == isz) done = true;
    if (done) {
      pts[2 * j] = nx_; pts[2 * j + 1] = nz_;
      pts[2 * (j + 1)] = scx; pts[2 * (j + 1) + 1] = scz;
      nrp = j + 2; sw = 1;
      break;
    }
    if (ipx < 1) { nx_ = C.gox; ipx = 1; }
    if (ipx >= C.nnx) { nx_ = C.gox + (C.nnx - 1) * C.dnx; ipx = C.nnx - 1; }
    if (ipz < 1) { nz_ = C.goz; ipz = 1; }
    if (ipz >= C.nnz) { nz_ = C.goz + (C.nnz - 1) * C.dnz; ipz = C.nnz - 1; }
    pts[2 * j] = nx_; pts[2 * j + 1] = nz_;
    cx = nx_; cz = nz_;
    if (j == maxrp - 1 && sw == 0) { (*crazy)++; sw = 1; break; }
  }
#undef FM_IGREF
#undef RTTN
#undef RTTNR
#undef RNSTSR
  return nrp;
}

// One (period, source) problem, by the heap warp (warp 0 of the block); the stencil warp serves fm_travel.
__device__ void fm2d_problem(const FmParams& P, FmGrid& G, volatile int& s_err, volatile int* s_sh, int lane) {
  const int prob = blockIdx.x;
  const int map = prob / P.nsrc, isrc = prob - map * P.nsrc; // 0-based
  const int32_t* srs = P.srs + (size_t)map * P.srs_ms + (size_t)isrc * P.nrc;
  int any = 0;
  for (int r = lane; r < P.nrc; r += 32) any += srs[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) any += __shfl_xor_sync(0xffffffffu, any, o);
  if (lane == 0) P.err[prob] = 0;
  if (any == 0 && isrc != 0) return; // no valid rays for this source (unless it is the first): cycle
  const double x = P.scx[isrc], z = P.scz[isrc];
  const double dnx0 = P.dvx / P.gdx, dnz0 = P.dvz / P.gdz;
  // scratch: coarse ttn | refined veln | refined slowness | refined ttn | coarse nsts | refined nsts | heap entries
  const size_t cc = (size_t)P.nnx * P.nnz, cr = (size_t)P.ldr * P.ldr;
  char* base = P.scratch + (size_t)prob * P.scratch_per_problem;
  double* ttn_c = (double*)base;
  double* veln_r = ttn_c + cc;
  double* slow_r = veln_r + cr;
  double* ttn_r = slow_r + cr;
  int32_t* nsts_c = (int32_t*)(ttn_r + cr);
  int32_t* nsts_r = nsts_c + cc;
  FmEnt* heap = (FmEnt*)(((uintptr_t)(nsts_r + cr) + 15) & ~(uintptr_t)15);
  int32_t* stage = (int32_t*)heap; // the (idle) heap doubles as staging between the two marches
  for (size_t q = lane; q < cc; q += 32) ttn_c[q] = 0.0; // nodes the march never reaches read 0 (see the check after the march)
  int isx = (int)((x - P.gox) / dnx0) + 1, isz = (int)((z - P.goz) / dnz0) + 1;
  if (isx < 1 || isx > P.nnx || isz < 1 || isz > P.nnz) { if (lane == 0) P.err[prob] = 1; return; }
  if (isx == P.nnx) isx--;
  if (isz == P.nnz) isz--;
  int vnl = isx - P.sgs; if (vnl < 1) vnl = 1;
  int vnr = isx + P.sgs; if (vnr > P.nnx) vnr = P.nnx;
  int vnt = isz - P.sgs; if (vnt < 1) vnt = 1;
  int vnb = isz + P.sgs; if (vnb > P.nnz) vnb = P.nnz;
  const double* veln_c = P.veln + (size_t)map * cc;
  const double* slow_c = P.slow + (size_t)map * cc;
  const int maxbt0 = (int)floor(P.snb * P.nnx * P.nnz + 0.5); // NINT of a positive value
  unsigned nacc = 0, nupd = 0;
  int nrnx = 0, nrnz = 0;
  double drnx = 0, drnz = 0, gorx = 0, gorz = 0;
  if (P.asgr == 1) {
    nrnx = (vnr - vnl) * P.sgdl + 1; nrnz = (vnb - vnt) * P.sgdl + 1;
    drnx = P.dvx / (double)(float)(P.gdx * P.sgdl); drnz = P.dvz / (double)(float)(P.gdz * P.sgdl);
    gorx = P.gox + dnx0 * (vnl - 1); gorz = P.goz + dnz0 * (vnt - 1);
    // bsplrefine (fm2dray_cartesian.f90:598-668): every refined node, the Fortran's (i,j,k,l) recovered from it
    const int nrxr = P.gdx * P.sgdl, nrzr = P.gdz * P.sgdl;
    const int origx = (vnl - 1) * P.sgdl + 1, origz = (vnt - 1) * P.sgdl + 1;
    const double* velv = P.velv + (size_t)map * P.vel_ms;
    for (int t = lane; t < nrnx * nrnz; t += 32) {
      const int idm2 = t / nrnz + 1, idm1 = t % nrnz + 1;
      const int st1 = idm1 + origz - 1, st2 = idm2 + origx - 1;
      int i = (st1 - 1) / nrzr + 1; if (i > P.nvz - 1) i = P.nvz - 1;
      int j = (st2 - 1) / nrxr + 1; if (j > P.nvx - 1) j = P.nvx - 1;
      const int k = st1 - nrzr * (i - 1), l = st2 - nrxr * (j - 1);
      double ui[5], vi[5], u;
      u = nrxr; u = (l - 1) / u; fm_bspl(u, ui);
      u = nrzr; u = (k - 1) / u; fm_bspl(u, vi);
      double sum[5];
      for (int i1 = 1; i1 <= 4; ++i1) {
        sum[i1] = 0.0;
        for (int j1 = 1; j1 <= 4; ++j1) sum[i1] = sum[i1] + ui[j1] * velv[((size_t)(j - 2 + j1) * (P.nvz + 2) + (i - 2 + i1)) * P.vel_es];
        sum[i1] = vi[i1] * sum[i1];
      }
      const double vr = sum[1] + sum[2] + sum[3] + sum[4];
      veln_r[(size_t)(idm2 - 1) * P.ldr + (idm1 - 1)] = vr;
      slow_r[(size_t)(idm2 - 1) * P.ldr + (idm1 - 1)] = 1.0 / vr;
    }
    for (size_t q = lane; q < cr; q += 32) nsts_r[q] = -1;
    __syncwarp();
    if (lane == 0) {
      G.nnx = nrnx; G.nnz = nrnz; G.ld = P.ldr; G.gox = gorx; G.goz = gorz; G.dnx = drnx; G.dnz = drnz;
      G.ld_magic = 0xFFFFFFFFu / (unsigned)P.ldr + 1u;
      G.veln = veln_r; G.slow = slow_r; G.ttn = ttn_r; G.nsts = nsts_r; G.heap = heap; G.fom = P.fom;
      G.vnl = vnl; G.vnr = vnr; G.vnt = vnt; G.vnb = vnb; G.error = 0; G.n_accept = 0; G.n_update = 0;
      int mb = maxbt0;
      if (nrnx > P.nnx || nrnz > P.nnz) { const int a = nrnx > P.nnx ? nrnx : P.nnx, b = nrnz > P.nnz ? nrnz : P.nnz; mb = (int)floor(P.snb * a * b + 0.5); }
      G.maxbt = mb;
    }
    __syncwarp();
    fm_travel(G, x, z, 1, lane, s_sh);
    if (lane == 0) { s_err = G.error; nacc = G.n_accept; nupd = G.n_update; }
    __syncwarp();
    if (s_err) { if (lane == 0) P.err[prob] = s_err; return; }
    // map the refined grid onto the coarse one (:341-372), then complete the narrow band (:398-417)
    for (size_t q = lane; q < cc; q += 32) nsts_c[q] = -1;
    __syncwarp();
    const int nk = (nrnz - 1) / P.sgdl + 1, nl = (nrnx - 1) / P.sgdl + 1;
    for (int t = lane; t < nk * nl; t += 32) {
      const int kk = t % nk, ll = t / nk;
      const int k = 1 + kk * P.sgdl, l = 1 + ll * P.sgdl;
      const int idm1 = vnt + kk, idm2 = vnl + ll;
      const int s = nsts_r[(size_t)(l - 1) * P.ldr + (k - 1)];
      nsts_c[(size_t)(idm2 - 1) * P.nnz + (idm1 - 1)] = s;
      if (s >= 0) ttn_c[(size_t)(idm2 - 1) * P.nnz + (idm1 - 1)] = ttn_r[(size_t)(l - 1) * P.ldr + (k - 1)];
    }
    __syncwarp();
    // alive nodes with a far neighbour become close: decided on the mapped state (a node set to 1 is neither 0 nor -1, so the
    // Fortran's in-place sweep sees the same neighbours); only the mapped window can hold alive nodes
    for (int t = lane; t < nk * nl; t += 32) {
      const int l = vnt + t % nk, k = vnl + t / nk; // l: z index, k: x index (the Fortran's names)
      if (nsts_c[(size_t)(k - 1) * P.nnz + (l - 1)] == 0) {
        bool far = false;
        if (l - 1 >= 1 && nsts_c[(size_t)(k - 1) * P.nnz + (l - 2)] == -1) far = true;
        if (l + 1 <= P.nnz && nsts_c[(size_t)(k - 1) * P.nnz + l] == -1) far = true;
        if (k - 1 >= 1 && nsts_c[(size_t)(k - 2) * P.nnz + (l - 1)] == -1) far = true;
        if (k + 1 <= P.nnx && nsts_c[(size_t)k * P.nnz + (l - 1)] == -1) far = true;
        stage[1 + t] = far ? 1 : 0; // applied after every lane has looked
      } else stage[1 + t] = 0;
    }
    __syncwarp();
    for (int t = lane; t < nk * nl; t += 32) {
      const int l = vnt + t % nk, k = vnl + t / nk;
      if (stage[1 + t]) nsts_c[(size_t)(k - 1) * P.nnz + (l - 1)] = 1;
    }
    __syncwarp();
    if (lane == 0) {
      G.nnx = P.nnx; G.nnz = P.nnz; G.ld = P.nnz; G.gox = P.gox; G.goz = P.goz; G.dnx = dnx0; G.dnz = dnz0;
      G.ld_magic = 0xFFFFFFFFu / (unsigned)P.nnz + 1u;
      G.veln = veln_c; G.slow = slow_c; G.ttn = ttn_c; G.nsts = nsts_c; G.n_accept = 0; G.n_update = 0;
    }
    __syncwarp();
    fm_travel(G, x, z, 2, lane, s_sh);
    if (lane == 0) { s_err = G.error; nacc += G.n_accept; nupd += G.n_update; }
  } else {
    for (size_t q = lane; q < cc; q += 32) nsts_c[q] = -1;
    __syncwarp();
    if (lane == 0) {
      G.nnx = P.nnx; G.nnz = P.nnz; G.ld = P.nnz; G.gox = P.gox; G.goz = P.goz; G.dnx = dnx0; G.dnz = dnz0;
      G.ld_magic = 0xFFFFFFFFu / (unsigned)P.nnz + 1u;
      G.veln = veln_c; G.slow = slow_c; G.ttn = ttn_c; G.nsts = nsts_c; G.heap = heap; G.fom = P.fom; G.maxbt = maxbt0;
      G.vnl = vnl; G.vnr = vnr; G.vnt = vnt; G.vnb = vnb; G.error = 0; G.n_accept = 0; G.n_update = 0;
    }
    __syncwarp();
    fm_travel(G, x, z, 0, lane, s_sh);
    if (lane == 0) { s_err = G.error; nacc = G.n_accept; nupd = G.n_update; }
  }
  __syncwarp();
  if (lane == 0 && P.counters) { atomicAdd(&P.counters[0], (unsigned long long)nacc); atomicAdd(&P.counters[1], (unsigned long long)nupd); }
  if (s_err) { if (lane == 0) P.err[prob] = s_err; return; }
  // Every node must have been accepted.  The Fortran's refined march stops at a refined-grid edge that it takes for an
  // interior one (its test compares the REFINED extent with a COARSE index, fm2d_ttime.f90:76-87): for a source in the
  // model's last cell row or column that is the source cell itself, the march dies after one or two nodes, and the Fortran
  // then returns whatever the shared ttn array held from the PREVIOUS source.  That result is not a function of this
  // problem's inputs; the problem is flagged (condition 6) and its unreached nodes read 0 here.
  {
    int unreached = 0;
    for (size_t q = lane; q < cc; q += 32) unreached += nsts_c[q] != 0 ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) unreached += __shfl_xor_sync(0xffffffffu, unreached, o);
    if (unreached > 0 && lane == 0) P.err[prob] = 6;
  }
  if (P.field) {
    double* f = P.field + (size_t)prob * cc;
    for (size_t q = lane; q < cc; q += 32) f[q] = ttn_c[q];
  }
  // srtimes (fm2dray_cartesian.f90:676-770): one receiver per lane
  double* tt = P.ttime + ((size_t)map * P.nsrc + isrc) * P.nrc;
  for (int r = lane; r < P.nrc; r += 32) {
    if (srs[r] == 0) continue;
    const double rx = P.rcx[r], rz = P.rcz[r];
    int irx = (int)floor((rx - P.gox) / dnx0) + 1, irz = (int)floor((rz - P.goz) / dnz0) + 1;
    if (irx < 1 || irx > P.nnx || irz < 1 || irz > P.nnz) { P.err[prob] = 3; continue; }
    if (irx == P.nnx) irx--;
    if (irz == P.nnz) irz--;
    const int jsx = (int)floor((x - P.gox) / dnx0) + 1, jsz = (int)floor((z - P.goz) / dnz0) + 1;
    double dpl = dnx0;
    if (dnz0 < dpl) dpl = dnz0;
    double sred = (x - rx) * (x - rx);
    sred = sred + (z - rz) * (z - rz);
    sred = sqrt(sred);
    int sw = 0;
    if (sred < dpl) sw = 1;
    if (jsx == irx && jsz == irz) sw = 1;
    double trr;
    if (sw) {
      FmGrid H; H.dnx = dnx0; H.dnz = dnz0;
      double vss[3][3];
      for (int k = 1; k <= 2; ++k) for (int l = 1; l <= 2; ++l) vss[k][l] = veln_c[(size_t)(jsx - 1 + k - 1) * P.nnz + (jsz - 1 + l - 1)];
      double drx = (x - P.gox) - (jsx - 1) * dnx0, drz = (z - P.goz) - (jsz - 1) * dnz0;
      const double vels = fm_bilinear(H, vss, drx, drz);
      for (int k = 1; k <= 2; ++k) for (int l = 1; l <= 2; ++l) vss[k][l] = veln_c[(size_t)(irx - 1 + k - 1) * P.nnz + (irz - 1 + l - 1)];
      drx = (rx - P.gox) - (irx - 1) * dnx0; drz = (rz - P.goz) - (irz - 1) * dnz0;
      const double velr = fm_bilinear(H, vss, drx, drz);
      trr = 2.0 * sred / (vels + velr);
    } else {
      const double drx = (rx - P.gox) - (irx - 1) * dnx0, drz = (rz - P.goz) - (irz - 1) * dnz0;
      trr = 0.0;
      for (int k = 1; k <= 2; ++k)
        for (int l = 1; l <= 2; ++l) {
          const double produ = (1.0 - fabs(((l - 1) * dnz0 - drz) / dnz0)) * (1.0 - fabs(((k - 1) * dnx0 - drx) / dnx0));
          trr = trr + ttn_c[(size_t)(irx - 1 + k - 1) * P.nnz + (irz - 1 + l - 1)] * produ;
        }
    }
    tt[r] = trr;
  }
  // rpaths (uar = 0): one receiver's ray per lane
  if (P.ray_npts) {
    FmRayCtx C;
    C.asgr = P.asgr; C.nnx = P.nnx; C.nnz = P.nnz; C.nnxr = nrnx; C.nnzr = nrnz; C.ldr = P.ldr;
    C.gox = P.gox; C.goz = P.goz; C.dnx = dnx0; C.dnz = dnz0; C.goxr = gorx; C.gozr = gorz; C.dnxr = drnx; C.dnzr = drnz;
    C.ttn = ttn_c; C.ttnr = ttn_r; C.nstsr = nsts_r;
    const int nrr = P.nrc * P.nsrc;
    const int32_t* srsv = P.srsv + (size_t)map * P.srsv_ms + (size_t)isrc * P.nrc;
    int crazy = 0;
    for (int r = lane; r < P.nrc; r += 32) {
      if (srs[r] == 0) continue;
      const int slot = srsv[r] - 1;
      if (slot < 0 || slot >= nrr) { P.err[prob] = 5; continue; }
      const size_t gs = (size_t)map * nrr + slot;
      double* pts = P.ray_pts + gs * (size_t)P.ray_cap * 2;
      int e = 0;
      const int n = fm_trace_ray(C, x, z, P.rcx[r], P.rcz[r], P.ray_cap, pts, &crazy, &e);
      if (e) { P.err[prob] = e; continue; }
      P.ray_npts[gs] = n;
      double len = 0;
      for (int k = 1; k < n; ++k) {
        const double ex = pts[2 * k] - pts[2 * (k - 1)], ez = pts[2 * k + 1] - pts[2 * (k - 1) + 1];
        const double d = ex * ex + ez * ez;
        len = len + sqrt(d);
      }
      P.ray_len[gs] = len;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) crazy += __shfl_xor_sync(0xffffffffu, crazy, o);
    if (lane == 0) P.crazy[prob] = crazy;
  }
}

// a block = one problem = two warps: warp 0 owns the heap and everything around the march, warp 1 solves the stencils
__global__ void __launch_bounds__(64) fm2d_kernel(const __grid_constant__ FmParams P) {
  __shared__ FmGrid G;
  __shared__ int s_err;
  __shared__ int s_sh[4];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) { fm_stencil_warp(G, s_sh, lane); return; }
  fm2d_problem(P, G, s_err, s_sh, lane);
  __syncwarp();
  if (lane == 0) s_sh[0] = FM_CMD_QUIT;
  fm_bar_a();
}

// ---- host side -------------------------------------------------------------------------------------------------------
namespace {
DevBuf fm_veln, fm_slow, fm_scratch, fm_err, fm_geo, fm_srs, fm_vel, fm_tt, fm_rays;

int fm2d_launch(FmParams& P, cudaStream_t st) {
  int rc;
  P.nnx = (P.nvx - 1) * P.gdx + 1; P.nnz = (P.nvz - 1) * P.gdz + 1;
  if (P.nnx > 32767 || P.nnz > 32767) return fail(MCT_E_INVALID_ARG, "fm2d: propagation grid larger than 32767 nodes along an axis");
  P.ldr = 2 * P.sgs * P.sgdl + 1;
  const size_t cc = (size_t)P.nnx * P.nnz, cr = (size_t)P.ldr * P.ldr;
  const int a = std::max(P.ldr, P.nnx), b = std::max(P.ldr, P.nnz);
  const size_t maxbt = (size_t)std::max(floor(P.snb * P.nnx * P.nnz + 0.5), floor(P.snb * a * b + 0.5)) + 4;
  const size_t win = (size_t)(2 * P.sgs + 1) * (2 * P.sgs + 1) + 4; // staging of the refined -> coarse mapping
  P.heap_bytes = (std::max(sizeof(FmEnt) * maxbt, 4 * win) + 15) & ~(size_t)15;
  size_t per = 8 * (cc + 3 * cr) + 4 * (cc + cr) + 16 + P.heap_bytes;
  per = (per + 15) & ~(size_t)15;
  P.scratch_per_problem = per;
  const int nprob = P.nmaps * P.nsrc;
  if ((rc = ensure(fm_veln, 8 * cc * (size_t)P.nmaps))) return rc;
  if ((rc = ensure(fm_slow, 8 * cc * (size_t)P.nmaps))) return rc;
  if ((rc = ensure(fm_scratch, per * (size_t)nprob))) return rc;
  P.veln = (double*)fm_veln.p; P.slow = (double*)fm_slow.p; P.scratch = (char*)fm_scratch.p;
  P.counters = g.count_on ? (unsigned long long*)g.counters.p + 10 : nullptr;
  ProfScope ps(2, st);
  fm2d_gridder_kernel<<<grid_blocks((long long)cc * P.nmaps, 256, 8), 256, 0, st>>>(P);
  fm2d_kernel<<<nprob, 64, 0, st>>>(P);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 2;
  return MCT_OK;
}
int fm2d_check(int nsrc, int nrc, int nmaps, int nvx, int nvz, double dvx, double dvz, const mct_fm2d_opts* o) {
  if (!o || nsrc < 1 || nrc < 1 || nmaps < 1 || nvx < 2 || nvz < 2 || !(dvx > 0) || !(dvz > 0)) return fail(MCT_E_INVALID_ARG, "fm2d: bad sizes");
  if (o->gridx < 1 || o->gridy < 1 || o->sgdic < 1 || o->sgext < 1 || (o->order != 0 && o->order != 1) || !(o->band > 0))
    return fail(MCT_E_INVALID_ARG, "fm2d: bad options");
  return MCT_OK;
}
void fm2d_fill(FmParams& P, int nsrc, int nrc, int nmaps, int nvx, int nvz, double gox, double goz, double dvx, double dvz, const mct_fm2d_opts* o) {
  memset(&P, 0, sizeof P);
  P.nmaps = nmaps; P.nsrc = nsrc; P.nrc = nrc; P.nvx = nvx; P.nvz = nvz;
  P.gox = gox; P.goz = goz; P.dvx = dvx; P.dvz = dvz;
  P.gdx = o->gridx; P.gdz = o->gridy; P.asgr = o->sgref ? 1 : 0; P.sgdl = o->sgdic; P.sgs = o->sgext; P.fom = o->order; P.snb = o->band;
}
} // namespace

extern "C" {

int mct_fm2d_times_dev(const double* d_src_xz, int nsrc, const double* d_rcv_xz, int nrc, const int32_t* d_srs, long long srs_map_stride,
                       const double* d_vel, long long vel_elem_stride, long long vel_map_stride, int nmaps, int nvx, int nvz, double gox,
                       double goz, double dvx, double dvz, const mct_fm2d_opts* o, double* d_ttime, int32_t* d_err, void* stream) {
  NEED_INIT();
  int rc;
  if ((rc = fm2d_check(nsrc, nrc, nmaps, nvx, nvz, dvx, dvz, o))) return rc;
  if (!d_src_xz || !d_rcv_xz || !d_srs || !d_vel || !d_ttime || !d_err) return fail(MCT_E_INVALID_ARG, "fm2d: NULL pointer");
  FmParams P;
  fm2d_fill(P, nsrc, nrc, nmaps, nvx, nvz, gox, goz, dvx, dvz, o);
  P.scx = d_src_xz; P.scz = d_src_xz + nsrc; P.rcx = d_rcv_xz; P.rcz = d_rcv_xz + nrc;
  P.srs = d_srs; P.srs_ms = srs_map_stride;
  P.velv = d_vel; P.vel_es = vel_elem_stride; P.vel_ms = vel_map_stride;
  P.ttime = d_ttime; P.err = d_err;
  return fm2d_launch(P, stream ? (cudaStream_t)stream : g.stream);
}

static int fm2d_host(const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z, int nrc, const int32_t* srs,
                     const int32_t* srsv, const double* vel, int nmaps, int nvx, int nvz, double gox, double goz, double dvx, double dvz,
                     const mct_fm2d_opts* o, double* ttime, double* field, int ray_cap, int32_t* ray_npts, double* ray_pts, double* ray_len,
                     int32_t* crazy) {
  int rc;
  if ((rc = fm2d_check(nsrc, nrc, nmaps, nvx, nvz, dvx, dvz, o))) return rc;
  if (!src_x || !src_z || !rcv_x || !rcv_z || !srs || !vel || !ttime) return fail(MCT_E_INVALID_ARG, "fm2d: NULL pointer");
  const bool rays = ray_npts != nullptr;
  if (rays && (!srsv || !ray_pts || !ray_len || !crazy || ray_cap < 4)) return fail(MCT_E_INVALID_ARG, "fm2d_rays: bad ray arguments");
  cudaStream_t st = g.stream;
  const size_t nv = (size_t)(nvz + 2) * (nvx + 2) * nmaps, nt = (size_t)nrc * nsrc * nmaps;
  const int nprob = nmaps * nsrc;
  if ((rc = ensure(fm_geo, 8 * (size_t)(2 * nsrc + 2 * nrc)))) return rc;
  if ((rc = ensure(fm_srs, 4 * nt * (rays ? 2 : 1)))) return rc;
  if ((rc = ensure(fm_vel, 8 * nv))) return rc;
  if ((rc = ensure(fm_tt, 8 * nt))) return rc;
  if ((rc = ensure(fm_err, 4 * (size_t)nprob * 2))) return rc;
  double* geo = (double*)fm_geo.p;
  CK(cudaMemcpyAsync(geo, src_x, 8 * (size_t)nsrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(geo + nsrc, src_z, 8 * (size_t)nsrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(geo + 2 * nsrc, rcv_x, 8 * (size_t)nrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(geo + 2 * nsrc + nrc, rcv_z, 8 * (size_t)nrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(fm_srs.p, srs, 4 * nt, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(fm_vel.p, vel, 8 * nv, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(fm_tt.p, ttime, 8 * nt, cudaMemcpyHostToDevice, st)); // entries without data keep the caller's values
  FmParams P;
  fm2d_fill(P, nsrc, nrc, nmaps, nvx, nvz, gox, goz, dvx, dvz, o);
  P.scx = geo; P.scz = geo + nsrc; P.rcx = geo + 2 * nsrc; P.rcz = geo + 2 * nsrc + nrc;
  P.srs = (const int32_t*)fm_srs.p; P.srs_ms = (long long)nrc * nsrc;
  P.velv = (const double*)fm_vel.p; P.vel_es = 1; P.vel_ms = (long long)(nvz + 2) * (nvx + 2);
  P.ttime = (double*)fm_tt.p; P.err = (int32_t*)fm_err.p;
  if (rays) {
    if ((rc = ensure(fm_rays, 8 * nt * ((size_t)ray_cap * 2 + 1) + 4 * nt))) return rc;
    CK(cudaMemcpyAsync((int32_t*)fm_srs.p + nt, srsv, 4 * nt, cudaMemcpyHostToDevice, st));
    P.srsv = (const int32_t*)fm_srs.p + nt; P.srsv_ms = (long long)nrc * nsrc;
    P.ray_cap = ray_cap;
    P.ray_pts = (double*)fm_rays.p; P.ray_len = P.ray_pts + nt * (size_t)ray_cap * 2; P.ray_npts = (int32_t*)(P.ray_len + nt);
    P.crazy = (int32_t*)fm_err.p + nprob;
    CK(cudaMemsetAsync(P.ray_len, 0, 8 * nt + 4 * nt, st)); // pairs without data: no ray (npoints = 0, length 0)
    CK(cudaMemsetAsync(P.crazy, 0, 4 * (size_t)nprob, st));
  }
  if (field) {
    const size_t nf = (size_t)((nvx - 1) * o->gridx + 1) * ((nvz - 1) * o->gridy + 1) * nprob;
    void* pf = nullptr;
    CK(cudaMalloc(&pf, 8 * nf));
    P.field = (double*)pf;
  }
  rc = fm2d_launch(P, st);
  std::vector<int32_t> herr((size_t)nprob * 2, 0);
  if (!rc) {
    cudaMemcpyAsync(ttime, fm_tt.p, 8 * nt, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(herr.data(), fm_err.p, 4 * (size_t)nprob * (rays ? 2 : 1), cudaMemcpyDeviceToHost, st);
    if (field) cudaMemcpyAsync(field, P.field, 8 * (size_t)P.nnx * P.nnz * nprob, cudaMemcpyDeviceToHost, st);
    if (rays) {
      cudaMemcpyAsync(ray_pts, P.ray_pts, 8 * nt * (size_t)ray_cap * 2, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(ray_len, P.ray_len, 8 * nt, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(ray_npts, P.ray_npts, 4 * nt, cudaMemcpyDeviceToHost, st);
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (P.field) cudaFree(P.field);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(MCT_E_CUDA, "fm2d: %s", cudaGetErrorString(e));
  if (rays) for (int m = 0; m < nmaps; ++m) { crazy[m] = 0; for (int i = 0; i < nsrc; ++i) crazy[m] += herr[(size_t)nprob + (size_t)m * nsrc + i]; }
  for (int p = 0; p < nprob; ++p)
    if (herr[p] && herr[p] != 6) return fail(MCT_E_INVALID_ARG, "fm2d: problem %d (period %d, source %d): %s", p, p / nsrc + 1, p % nsrc + 1,
                             herr[p] == 1 ? "source outside the model" : herr[p] == 2 ? "narrow band exceeds band*nx*ny" :
                             herr[p] == 3 ? "receiver outside the model" : "ray slot (raystat(:,2,:)) outside 1..nrev*nsrc");
  for (int p = 0; p < nprob; ++p)
    if (herr[p] == 6) return fail(MCT_E_FM2D_STALE, "fm2d: period %d, source %d lies in the model's last cell row/column: the reference's refined march "
                                  "stops at once there and returns the previous source's field; times of that source are not comparable", p / nsrc + 1, p % nsrc + 1);
  return MCT_OK;
}

int mct_fm2d_times(const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z, int nrc, const int32_t* srs,
                   const double* vel, int nmaps, int nvx, int nvz, double gox, double goz, double dvx, double dvz, const mct_fm2d_opts* o,
                   double* ttime, double* field) {
  NEED_INIT();
  return fm2d_host(src_x, src_z, nsrc, rcv_x, rcv_z, nrc, srs, nullptr, vel, nmaps, nvx, nvz, gox, goz, dvx, dvz, o, ttime, field, 0, nullptr,
                   nullptr, nullptr, nullptr);
}

int mct_fm2d_rays(const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z, int nrc, const int32_t* srs,
                  const int32_t* srsv, const double* vel, int nmaps, int nvx, int nvz, double gox, double goz, double dvx, double dvz,
                  const mct_fm2d_opts* o, double* ttime, int ray_cap, int32_t* ray_npts, double* ray_pts, double* ray_len, int32_t* crazy) {
  NEED_INIT();
  if (!ray_npts) return fail(MCT_E_INVALID_ARG, "fm2d_rays: NULL ray_npts");
  return fm2d_host(src_x, src_z, nsrc, rcv_x, rcv_z, nrc, srs, srsv, vel, nmaps, nvx, nvz, gox, goz, dvx, dvz, o, ttime, nullptr, ray_cap, ray_npts,
                   ray_pts, ray_len, crazy);
}

int mct_fm2d_stats(int64_t out2[2]) {
  NEED_INIT();
  if (!out2) return fail(MCT_E_INVALID_ARG, "fm2d_stats: NULL pointer");
  unsigned long long c[2];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(c, (unsigned long long*)g.counters.p + 10, sizeof c, cudaMemcpyDeviceToHost));
  out2[0] = (int64_t)c[0]; out2[1] = (int64_t)c[1];
  return MCT_OK;
}

} // extern "C"

// ---- session: the curved-ray likelihood of surf_likelihood on resident maps --------------------------------------------
// like%vel(np, ny+2, nx+2) from the session's phase map (+ the pending proposal's window): interior = the map, edges
// replicated (likelihood_surf.F90:259-264; a map kept consistent call after call is exactly the clamped copy).
__global__ void __launch_bounds__(256) fm2d_pad_kernel(const double* __restrict__ vel, VelOverlay ov, int np, int nx, int ny, double* __restrict__ out) {
  const long long n = (long long)np * (ny + 2) * (nx + 2);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % np);
    const long long r = t / np;
    int iy = (int)(r % (ny + 2)), ix = (int)(r / (ny + 2));
    iy = iy < 1 ? 1 : (iy > ny ? ny : iy);
    ix = ix < 1 ? 1 : (ix > nx ? nx : ix);
    out[t] = vel_at(vel, ov, np, p, ny, iy, ix);
  }
}

extern "C" {

// Sources, receivers and the fast-marching settings of a session (made resident once, like the straight rays).
int mct_session_set_fm2d(mct_session* s, const double* src_x, const double* src_z, int nsrc, const double* rcv_x, const double* rcv_z, int nrc,
                         const mct_fm2d_opts* o) {
  NEED_INIT();
  if (!s || !src_x || !src_z || !rcv_x || !rcv_z) return fail(MCT_E_INVALID_ARG, "session_set_fm2d: NULL pointer");
  int rc;
  if ((rc = fm2d_check(nsrc, nrc, s->np, s->gr.nx, s->gr.ny, s->gr.dx, s->gr.dy, o))) return rc;
  if (s->nout != s->np) return fail(MCT_E_INVALID_ARG, "session_set_fm2d: needs a single-mode session");
  cudaStream_t st = g.stream;
  if ((rc = ensure(s->f_geo, 8 * (size_t)(2 * nsrc + 2 * nrc)))) return rc;
  double* geo = (double*)s->f_geo.p;
  CK(cudaMemcpyAsync(geo, src_x, 8 * (size_t)nsrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(geo + nsrc, src_z, 8 * (size_t)nsrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(geo + 2 * nsrc, rcv_x, 8 * (size_t)nrc, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(geo + 2 * nsrc + nrc, rcv_z, 8 * (size_t)nrc, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  s->f_nsrc = nsrc; s->f_nrc = nrc;
  s->f_opt_i[0] = o->gridx; s->f_opt_i[1] = o->gridy; s->f_opt_i[2] = o->sgref; s->f_opt_i[3] = o->sgdic; s->f_opt_i[4] = o->sgext; s->f_opt_i[5] = o->order;
  s->f_band = o->band;
  s->f_have = true;
  return MCT_OK;
}

// surf_likelihood with curved rays (settings%isStraight == 0, likelihood_surf.F90:244-404) on the session's resident
// maps: like%vel assembled on the device from the phase map, every (period, source) marched in one launch; phase-velocity
// data: like%srdist = like%phaseTime; group-velocity data: the rays are traced as well (uar = 0), any crazy ray makes the
// likelihood huge, else CalGroupTime integrates like%gvel along them and like%srdist is their length; then the noise
// level and the Gaussian sums.  pending = 0: the current model, 1: the pending proposal.  phase_time, sigma: optional
// host outputs (nrr, np).
int mct_session_likelihood_fm2d(mct_session* s, int pending, const double* snoise0, const double* snoise1, double out[3], double* phase_time,
                                double* sigma) {
  NEED_INIT();
  if (!s || !out) return fail(MCT_E_INVALID_ARG, "session_likelihood_fm2d: NULL pointer");
  if (!s->f_have) return fail(MCT_E_INVALID_ARG, "session_likelihood_fm2d: no sources / receivers (call mct_session_set_fm2d first)");
  if (!s->mf.have) return fail(MCT_E_INVALID_ARG, "session_likelihood_fm2d: no data (call mct_session_set_data first)");
  const bool group = s->opt.phaseGroup == 1;
  const int nrr = s->f_nsrc * s->f_nrc;
  if (s->mf.nrr != nrr) return fail(MCT_E_INVALID_ARG, "session_likelihood_fm2d: the data hold %d source-receiver pairs, the geometry %d", s->mf.nrr, nrr);
  VelOverlay ov; // of like%gvel: the group window when phaseGroup == 1, else the phase window
  int rc = session_overlay(s, pending, ov);
  if (rc) return rc;
  VelOverlay ovp = ov; // of the phase map: rays are always bent by the phase velocity (like%vel = pvel, likelihood_surf.F90:259)
  if (ov.w) ovp.w = (const double*)s->w_pvel.p;
  cudaStream_t st = g.stream;
  const int np = s->np, nx = s->gr.nx, ny = s->gr.ny;
  const int nprob = np * s->f_nsrc;
  const size_t npad = (size_t)np * (ny + 2) * (nx + 2), nt = (size_t)np * nrr;
  if ((rc = ensure(s->f_vel, 8 * npad))) return rc;
  if ((rc = ensure(s->f_err, 4 * (size_t)nprob * 2))) return rc;
  if ((rc = ensure(s->time, 8 * nt))) return rc;
  {
    ProfScope ps(2, st);
    fm2d_pad_kernel<<<grid_blocks((long long)npad, 256, 8), 256, 0, st>>>((const double*)s->pvel.p, ovp, np, nx, ny, (double*)s->f_vel.p);
  }
  g.host_stats.n_launches += 1;
  CK(cudaMemsetAsync(s->time.p, 0, 8 * nt, st)); // pairs without data: never read by the misfit
  mct_fm2d_opts o{s->f_opt_i[0], s->f_opt_i[1], s->f_opt_i[2], s->f_opt_i[3], s->f_opt_i[4], s->f_opt_i[5], s->f_band};
  FmParams P;
  fm2d_fill(P, s->f_nsrc, s->f_nrc, np, nx, ny, s->gr.xmin, s->gr.ymin, s->gr.dx, s->gr.dy, &o);
  const double* geo = (const double*)s->f_geo.p;
  P.scx = geo; P.scz = geo + s->f_nsrc; P.rcx = geo + 2 * s->f_nsrc; P.rcz = geo + 2 * s->f_nsrc + s->f_nrc;
  P.srs = (const int32_t*)s->mf.raystat.p; P.srs_ms = 2LL * nrr; // dat%raystat(nrr, 2, np): [.,1,period]
  P.velv = (const double*)s->f_vel.p; P.vel_es = np; P.vel_ms = 1;  // like%vel(np, ny+2, nx+2) in place
  P.ttime = (double*)s->time.p; P.err = (int32_t*)s->f_err.p;
  const int cap = 8 * ((nx - 1) * o.gridx + 1 + (ny - 1) * o.gridy + 1);
  if (group) { // uar = 0: the rays as well (their lengths are like%srdist, CalGroupTime integrates like%gvel along them)
    if ((rc = ensure(s->f_rays, 8 * nt * ((size_t)cap * 2 + 1) + 4 * nt))) return rc;
    P.srsv = (const int32_t*)s->mf.raystat.p + nrr; P.srsv_ms = 2LL * nrr; // [.,2,period]
    P.ray_cap = cap;
    P.ray_pts = (double*)s->f_rays.p; P.ray_len = P.ray_pts + nt * (size_t)cap * 2; P.ray_npts = (int32_t*)(P.ray_len + nt);
    P.crazy = (int32_t*)s->f_err.p + nprob;
    CK(cudaMemsetAsync(P.ray_len, 0, 8 * nt + 4 * nt, st));
    CK(cudaMemsetAsync(P.crazy, 0, 4 * (size_t)nprob, st));
  }
  if ((rc = fm2d_launch(P, st))) return rc;
  s->time_nrays = nrr;
  std::vector<int32_t> herr((size_t)nprob * 2, 0);
  CK(cudaMemcpyAsync(herr.data(), s->f_err.p, 4 * (size_t)nprob * (group ? 2 : 1), cudaMemcpyDeviceToHost, st));
  const double* d_srdist = (const double*)s->time.p; // like%srdist = like%phaseTime (:327-333)
  if (group) {
    CK(cudaStreamSynchronize(st));
    for (int p = 0; p < nprob; ++p)
      if (herr[p]) return fail(MCT_E_INVALID_ARG, "session_likelihood_fm2d: period %d, source %d: condition %d", p / s->f_nsrc + 1, p % s->f_nsrc + 1, herr[p]);
    long long ncrazy = 0;
    for (int p = 0; p < nprob; ++p) ncrazy += herr[(size_t)nprob + p];
    if (ncrazy > 0) { // any(crazyray > 0): like%like = huge(like%like), return (likelihood_surf.F90:338-343)
      out[0] = 1.7976931348623157e308; out[1] = 0.0; out[2] = 0.0;
      return MCT_OK;
    }
    {
      ProfScope ps(2, st);
      group_times_slots_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>((const double*)s->gvel.p, ov, np, nx, ny, s->gr.xmin, s->gr.ymin, s->gr.dx,
                                                                            s->gr.dy, P.ray_pts, P.ray_npts, cap, nrr, (double*)s->time.p);
    }
    CK(cudaGetLastError());
    g.host_stats.n_launches += 1;
    d_srdist = P.ray_len; // like%srdist = phaseRays%length()
  }
  if (phase_time) CK(cudaMemcpyAsync(phase_time, s->time.p, 8 * nt, cudaMemcpyDeviceToHost, st));
  rc = misfit_run(s->mf, (const double*)s->time.p, snoise0, snoise1, out, sigma, st, d_srdist);
  for (int p = 0; p < nprob; ++p)
    if (herr[p]) return fail(herr[p] == 6 ? MCT_E_FM2D_STALE : MCT_E_INVALID_ARG, "session_likelihood_fm2d: period %d, source %d: %s", p / s->f_nsrc + 1, p % s->f_nsrc + 1,
                             herr[p] == 1 ? "source outside the model" : herr[p] == 2 ? "narrow band exceeds band*nx*ny" :
                             herr[p] == 3 ? "receiver outside the model" : "source in the model's last cell row/column (the reference returns the previous source's field there)");
  return rc;
}

} // extern "C"

namespace {
void release_fm2d_globals() {
  DevBuf* bufs[] = {&fm_veln, &fm_slow, &fm_scratch, &fm_err, &fm_geo, &fm_srs, &fm_vel, &fm_tt, &fm_rays};
  for (DevBuf* b : bufs) { if (b->p) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
}
} // namespace
