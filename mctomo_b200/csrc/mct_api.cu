// mct_api.cu -- the C ABI of include/mctomo_b200.h: context, device buffers, host-side kd-tree
// build, kernel launches and the host<->device staging of the host-pointer entry points.
//
// Host logic mirrored here (reference lines per function in the header):
//   kdtree_to_grid   src/mcmc_loc2.f90:2002-2082        -> mct_voronoi_to_grid[_dev]
//   surf_likelihood  src/likelihood_surf.F90:155-231    -> mct_surf_dispersion[_dev]
//   program modelling src/forward_modelling.f90:393-429 -> same, with its layer_eps / preset
//   surfmodes / surfmmodes surfmodes/surfmodes.f90:39-183 -> mct_surfmodes_batch
// There is no CPU implementation of any kernel in this library.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/mctomo_b200.h"
#include "k1_voronoi.cuh"
#include "k2_dispersion.cuh"
#include "k2_dedup.cuh"
#include "k5_grt.cuh"
#include "kdtree_build.h"

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct Ctx {
  bool init = false;
  int device = -1;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  int count_on = 1;
  char err[512] = {0};
  // resident nuclei set: nb trees packed back to back
  DevBuf nodes, rpts, ind, params, kmodels;
  HostKdTree tree;
  int nset = 0;                       // models in the resident set
  std::vector<long long> node_off, pt_off; // per model offsets into nodes / points
  std::vector<int> roots, ncell;
  // optional per-kernel timing (mct_set_profiling)
  int prof_on = 0;
  struct Ev { cudaEvent_t a, b; int kind; };
  std::vector<Ev> ev_pending, ev_pool;
  double prof_ms[4] = {0, 0, 0, 0};   // K1, K2, other kernels, number of timed launches
  // staged model (host-pointer entry points)
  DevBuf m_vp, m_vs, m_rho, m_sites;
  // layered columns, their processing order, sort scratch
  DevBuf lay, layr, nlay, status, perm, bins;
  DevBuf layc, layrc, nlayc; // compact inputs of the columns to solve (sorted-list order)
  int compact_inputs = 1;    // MCT_COMPACT=0: per-column layout also for large batches (A/B)
  DevBuf dd_table, dd_i32; // de-duplication: hash table; rep0|minrep|mult0|rep|mult|kstat|skey (7 x stride int32) + neff
  // balanced sharding of one chain (mct_comm.cuh): rank shard_r of shard_n solves every shard_n-th entry of the sorted list
  int shard_n = 1, shard_r = 0;
  DevBuf sh_perm, sh_p, sh_g, sh_i;
  float sort_proxy = 1.f;  // > 0: secondary sort key (search-length proxy, 8 levels of 1/sort_proxy km/s) inside a layer-count bin (MCT_SORT_PROXY)
  int dedup = 1;           // fold bit-identical layer stacks before K2 (mct_set_dedup / MCT_DEDUP)
  int last_ncol = 0, last_neff = 0, last_lanes = 0; // what the last dispersion launch did (mct_last_launch)
  char last_kernel[48] = {0};
  bool k1_smem_set = false; // dynamic shared-memory opt-in of k1_column_kernel done on this device
  cudaEvent_t stage_ev = nullptr; // last use of the pinned nuclei staging block (pin_small)
  int k1_mode = 0;          // 0 culled brute force per column (+ tree replay for ties), 1 tree walk for every node
  int k2_mode = 0;          // 0 auto, 1 always one thread per column, 2 always one warp per column
  int k2_coop_max = 0;      // auto mode: batches below this many columns take the lane-cooperative kernels
                            // (0 = 1.7 x the resident lanes of the GPU: 128 819 columns on a B200)
  int k2_coop_lanes = 0;    // lanes per column of the cooperative kernel: 0 = choose per launch (MCT_K2_COOP_LANES)
  int sort_stable = 1;      // stable counting sort of the columns (MCT_SORT_STABLE=0: atomic-cursor sort)
  int k2_variant = 7; // 7 production, 3 plain secular functions (A/B reference), 0 plain + unsorted: launch shape of K2 (see launch_k2); MCT_K2_VARIANT overrides (experiments)
  // outputs (host-pointer entry points)
  DevBuf o_pvel, o_gvel, o_ierr;
  // prelayered staging
  DevBuf pl_thick, pl_vp, pl_vs, pl_rho, pl_off;
  DevBuf flags;    // int32[4]: [0] model_invalid, [1] max status, [2] k1 error
  DevBuf bflags;   // int32[2*nb] for host-pointer batch calls
  DevBuf counters; // u64[16]: [0..2] represented work, [3] self-test scratch, [4..6] executed work, [8..9] generalized R/T kernel
  // low-velocity columns (k5_grt.cuh): mct_set_grt
  int grt_on = 0;
  double grt_par[6] = {1e-6, 1e-5, (double)1e-3f, (double)5e-3f, 1e-3, 1e-3}; // tolmin, tolmax, smin_min, smin_max, dcm, dc2 (likelihood_surf.F90:175-182, tol = 1e-6)
  DevBuf grt_list, grt_scratch;
  long long grt_last_cols = 0;
  DevBuf ray_pts, ray_off, ray_time; // mct_group_times_dev staging
  PinBuf pin_a, pin_b, pin_small;
  mct_stats host_stats = {0, 0, 0, 0, 0, 0, 0, 0};
};

Ctx g;
// One context per process.  Every entry point takes this lock for its whole duration (NEED_INIT), so calls from several
// host threads -- e.g. two chains driven by OpenMP threads of one rank -- are safe; they serialise on the host side
// (the *_dev entry points only enqueue, so device work of different streams still overlaps).
std::recursive_mutex g_mu;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g.err, sizeof g.err, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(MCT_E_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

#define NEED_INIT()                                                            \
  std::lock_guard<std::recursive_mutex> api_lock_(g_mu);                       \
  if (!g.init) return fail(MCT_E_NOINIT, "mct_init has not been called")

int ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return MCT_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return MCT_OK;
}
int ensure_pin(PinBuf& b, size_t bytes) {
  if (bytes <= b.cap) return MCT_OK;
  if (b.p) CK(cudaFreeHost(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMallocHost(&b.p, want));
  b.cap = want;
  return MCT_OK;
}
void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}
void release(PinBuf& b) {
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
}

inline cudaStream_t pick(void* stream) { return stream ? (cudaStream_t)stream : g.stream; }

bool grid_ok(const mct_grid* gr) {
  return gr && gr->nx >= 1 && gr->ny >= 1 && gr->nz >= 1 && gr->dx > 0 && gr->dy > 0 && gr->dz > 0;
}

// floor((b - min)/d) + 1, clamped: src/mcmc_loc2.f90:2034-2045
void box_window(const mct_grid* gr, const double box[6], int32_t w[6]) {
  w[0] = (int32_t)std::floor((box[0] - gr->xmin) / gr->dx) + 1;
  w[1] = (int32_t)std::floor((box[3] - gr->xmin) / gr->dx) + 1;
  w[2] = (int32_t)std::floor((box[1] - gr->ymin) / gr->dy) + 1;
  w[3] = (int32_t)std::floor((box[4] - gr->ymin) / gr->dy) + 1;
  w[4] = (int32_t)std::floor((box[2] - gr->zmin) / gr->dz) + 1;
  w[5] = (int32_t)std::floor((box[5] - gr->zmin) / gr->dz) + 1;
  if (w[0] < 1) w[0] = 1;
  if (w[2] < 1) w[2] = 1;
  if (w[4] < 1) w[4] = 1;
  if (w[1] > gr->nx) w[1] = gr->nx;
  if (w[3] > gr->ny) w[3] = gr->ny;
  if (w[5] > gr->nz) w[5] = gr->nz;
}

// Build the trees on the host and upload them together with the nuclei parameters.  Model b owns
// nuclei offsets[b] .. offsets[b+1]-1 of points/params.
int upload_nuclei(const double* points, const double* params, const long long* offsets, int nb, cudaStream_t st) {
  if (!points || !params || !offsets || nb < 1) return fail(MCT_E_INVALID_ARG, "nuclei: NULL pointer or empty batch");
  const long long o0 = offsets[0];
  const long long ntot = offsets[nb] - o0;
  if (ntot < nb) return fail(MCT_E_INVALID_ARG, "nuclei: a model has no cells");
  std::vector<KdNodeDev> all_nodes;
  std::vector<double> all_rpts((size_t)ntot * 3);
  std::vector<int32_t> all_ind((size_t)ntot);
  g.node_off.assign(nb, 0); g.pt_off.assign(nb, 0); g.roots.assign(nb, 0); g.ncell.assign(nb, 0);
  for (int b = 0; b < nb; ++b) {
    const long long n = offsets[b + 1] - offsets[b];
    if (n < 1) return fail(MCT_E_INVALID_ARG, "nuclei: model %d has no cells", b);
    KdBuilder(points + 3 * offsets[b], (int)n, g.tree).run();
    if (g.tree.degenerate)
      return fail(MCT_E_DEGENERATE_NUCLEI, "model %d: more than 13 nuclei coincide in all three coordinates: the reference kd-tree build does not terminate", b);
    g.node_off[b] = (long long)all_nodes.size();
    g.pt_off[b] = offsets[b] - o0;
    g.roots[b] = g.tree.root;
    g.ncell[b] = (int)n;
    all_nodes.insert(all_nodes.end(), g.tree.nodes.begin(), g.tree.nodes.end());
    memcpy(&all_rpts[3 * (size_t)g.pt_off[b]], g.tree.rpts.data(), sizeof(double) * 3 * (size_t)n);
    memcpy(&all_ind[(size_t)g.pt_off[b]], g.tree.ind.data(), sizeof(int32_t) * (size_t)n);
  }
  const size_t nb_nodes = all_nodes.size() * sizeof(KdNodeDev);
  const size_t nb_pts = 3 * sizeof(double) * (size_t)ntot;
  const size_t nb_ind = sizeof(int32_t) * (size_t)ntot;
  int rc;
  if ((rc = ensure(g.nodes, nb_nodes))) return rc;
  if ((rc = ensure(g.rpts, nb_pts))) return rc;
  if ((rc = ensure(g.ind, nb_ind))) return rc;
  if ((rc = ensure(g.params, nb_pts))) return rc;
  // per-model descriptors for the batched nearest-nucleus launch
  std::vector<K1Model> km((size_t)nb);
  for (int b = 0; b < nb; ++b) km[b] = K1Model{g.node_off[b], g.pt_off[b], g.roots[b], g.ncell[b]};
  const size_t nb_km = sizeof(K1Model) * (size_t)nb;
  if ((rc = ensure(g.kmodels, nb_km))) return rc;
  // one pinned staging block so the five small copies are true async DMA
  const size_t tot = nb_nodes + 2 * nb_pts + nb_ind + nb_km;
  if ((rc = ensure_pin(g.pin_small, tot))) return rc;
  // the staging block may still be in flight from the previous call -- possibly on another stream: wait for the
  // event recorded after its last use, not for the caller's stream
  if (g.stage_ev) CK(cudaEventSynchronize(g.stage_ev));
  char* h = (char*)g.pin_small.p;
  memcpy(h, all_nodes.data(), nb_nodes);
  memcpy(h + nb_nodes, all_rpts.data(), nb_pts);
  memcpy(h + nb_nodes + nb_pts, params + 3 * o0, nb_pts);
  memcpy(h + nb_nodes + 2 * nb_pts, km.data(), nb_km); // (8-byte aligned: before the int32 block)
  memcpy(h + nb_nodes + 2 * nb_pts + nb_km, all_ind.data(), nb_ind);
  CK(cudaMemcpyAsync(g.nodes.p, h, nb_nodes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.rpts.p, h + nb_nodes, nb_pts, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.params.p, h + nb_nodes + nb_pts, nb_pts, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.kmodels.p, h + nb_nodes + 2 * nb_pts, nb_km, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.ind.p, h + nb_nodes + 2 * nb_pts + nb_km, nb_ind, cudaMemcpyHostToDevice, st));
  if (!g.stage_ev) CK(cudaEventCreateWithFlags(&g.stage_ev, cudaEventDisableTiming));
  CK(cudaEventRecord(g.stage_ev, st));
  g.nset = nb;
  return MCT_OK;
}
int upload_nuclei(const double* points, const double* params, int ncells, cudaStream_t st) {
  const long long off[2] = {0, ncells};
  if (ncells < 1) return fail(MCT_E_INVALID_ARG, "nuclei: ncells < 1");
  return upload_nuclei(points, params, off, 1, st);
}

// ---- optional per-kernel timing ------------------------------------------------------------------
struct ProfScope {
  Ctx::Ev ev{};
  bool on;
  cudaStream_t st;
  ProfScope(int kind, cudaStream_t s) : on(g.prof_on != 0), st(s) {
    if (!on) return;
    if (!g.ev_pool.empty()) { ev = g.ev_pool.back(); g.ev_pool.pop_back(); }
    else { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); }
    ev.kind = kind;
    cudaEventRecord(ev.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(ev.b, st);
    g.ev_pending.push_back(ev);
  }
};

int shard_exchange(int neff, int nout, bool with_group, double* d_pvel, double* d_gvel, int32_t* d_ierr, cudaStream_t st); // mct_comm.cuh

int grid_blocks(long long work_items, int threads, int per_sm) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = (long long)g.sm_count * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Launch K1 on device arrays with the given array geometry.
int launch_k1(const mct_grid* gr, const int32_t w[6], const double* pm, double* d_vp, double* d_vs, double* d_rho,
              int32_t* d_sites, int ia0, int ja0, int ka0, int ny_a, int nz_a, cudaStream_t st, int model = 0,
              int nbatch = 0, long long model_stride = 0) {
  // nbatch > 0: all models of the resident set in ONE launch (culled brute-force kernel only); model b's
  // arrays start at b*model_stride.
  const bool batched = nbatch > 0 && g.k1_mode != 1;
  K1Params P;
  P.models = batched ? (const K1Model*)g.kmodels.p : nullptr;
  P.model_stride = model_stride;
  P.nodes = (const KdNodeDev*)g.nodes.p + (batched ? 0 : g.node_off[model]);
  P.rpts = (const double*)g.rpts.p + (batched ? 0 : 3 * g.pt_off[model]);
  P.ind = (const int32_t*)g.ind.p + (batched ? 0 : g.pt_off[model]);
  P.params = (const double*)g.params.p + (batched ? 0 : 3 * g.pt_off[model]);
  P.root = g.roots[model];
  P.ix0 = w[0]; P.iy0 = w[2]; P.iz0 = w[4];
  P.wx = w[1] - w[0] + 1; P.wy = w[3] - w[2] + 1; P.wz = w[5] - w[4] + 1;
  if (P.wx <= 0 || P.wy <= 0 || P.wz <= 0) return MCT_OK; // empty window: the Fortran loops do nothing
  P.xmin = gr->xmin; P.ymin = gr->ymin; P.zmin = gr->zmin;
  P.dx = gr->dx; P.dy = gr->dy; P.dz = gr->dz;
  P.ia0 = ia0; P.ja0 = ja0; P.ka0 = ka0; P.ny_a = ny_a; P.nz_a = nz_a;
  P.vp = d_vp; P.vs = d_vs; P.rho = d_rho; P.sites = d_sites;
  P.use_pm = pm ? 1 : 0;
  P.pm_vp = pm ? pm[0] : 0.0;
  P.pm_vs = pm ? pm[1] : 0.0;
  P.pm_eps = (double)1e-8f; // real(kind=ii10), parameter :: eps = 1e-8 (mcmc_loc2.f90:51)
  P.err = (int32_t*)g.flags.p + 2;
  const long long total = (long long)P.wx * P.wy * P.wz;
  P.n = g.ncell[model];
  {
    ProfScope ps(0, st);
    if (g.k1_mode == 1) {
      k1_voronoi_kernel<<<grid_blocks(total, 256, 16), 256, 0, st>>>(P); // exact tree walk for every node
    } else if (g.k1_mode == 2 || P.wz > 500) { // round-1 shape (also for columns beyond the tile kernel's 16-bit item index): warp per column, every node scans the column's survivors
      if (!g.k1_smem_set) {
        CK(cudaFuncSetAttribute(k1_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1C_SMEM_BYTES));
        g.k1_smem_set = true;
      }
      const long long ntiles = (long long)((P.wx + K1C_TILE - 1) / K1C_TILE) * ((P.wy + K1C_TILE - 1) / K1C_TILE);
      const int nby = batched ? nbatch : 1;
      const int blocks = (int)std::min<long long>(ntiles, std::max<long long>(1, (long long)g.sm_count * 3 / nby));
      k1_column_kernel<<<dim3(blocks, nby), 32 * K1C_WARPS, K1C_SMEM_BYTES, st>>>(P);
    } else { // per-(tile, z-segment) candidate lists, flattened nodes, vector stores (k1_tile.cuh)
      // Tile and segment are sized against the nucleus spacing h = (grid volume / n)^(1/3): the tile's half diagonal up
      // to h/2, segments of h/2.  Larger boxes make the bound U_B loose (long lists), smaller ones repeat the passes over
      // the nuclei for too few nodes (measured: C2 with 1 x 4 tiles spent 90 % of its instructions in those passes).
      const double vol = std::max(gr->dx * (gr->nx - 1), gr->dx) * std::max(gr->dy * (gr->ny - 1), gr->dy) * std::max(gr->dz * (gr->nz - 1), gr->dz);
      const double h = gr->nz > 1 ? std::cbrt(vol / (double)std::max(1, P.n))
                                  : std::sqrt(std::max(gr->dx * (gr->nx - 1), gr->dx) * std::max(gr->dy * (gr->ny - 1), gr->dy) / (double)std::max(1, P.n)); // (the 2-D product)
      static const int shapes[5][2] = {{8, 32}, {4, 16}, {2, 8}, {1, 8}, {1, 4}};
      int ttx = 1, tty = 4;
      for (const auto& sh : shapes) {
        const double hd = 0.5 * std::sqrt((sh[0] * gr->dx) * (sh[0] * gr->dx) + (sh[1] * gr->dy) * (sh[1] * gr->dy));
        if (hd <= 0.5 * h) { ttx = sh[0]; tty = sh[1]; break; }
      }
      int seglen = (int)std::floor(0.5 * h / gr->dz);
      seglen = std::max(seglen, (P.wz + K1T_MAXSEG - 1) / K1T_MAXSEG);
      seglen = std::max(2, seglen + (seglen & 1)); // even: a 16-byte pair of nodes then lies inside one segment
      if (const char* v = getenv("MCT_K1_TILE")) { int a = 0, b = 0, c = 0; if (sscanf(v, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0) { ttx = a; tty = b; seglen = std::max(c, (P.wz + K1T_MAXSEG - 1) / K1T_MAXSEG); } } // experiments
      const int nby = batched ? nbatch : 1;
      const int npairs = (P.wz + 2) / 2;
      if (g.k1_mode == 3) { // round-2 first shape: float64 walk, items ordered column-major
        const long long ntiles = (long long)((P.wx + ttx - 1) / ttx) * ((P.wy + tty - 1) / tty);
        const long long want = std::max<long long>(1, ((long long)g.sm_count * 16 + nby - 1) / nby);
        const int blocks = (int)std::min<long long>(ntiles, want);
        const int nitems = ttx * tty * npairs;
        const int threads = nitems <= 640 ? 128 : 256;
        k1_tile_kernel<<<dim3(blocks, nby), threads, 0, st>>>(P, ttx, tty, seglen, (float)(1.25 * h), k1t_magic(npairs), k1t_magic(tty),
                                                              k1t_magic(seglen));
      } else { // warp per box, float32 walk, 4 (or 2) nodes per thread, persistent over (model, tile) (k1_box.cuh)
        const bool forced = getenv("MCT_K1_TILE") != nullptr;
        if (!forced) {
          if (P.wz > 2) seglen = std::max(4, (seglen + 3) / 4 * 4); // whole 4-node units per segment
          const int ups = std::max(1, seglen / (P.wz > 2 ? 4 : 2));
          while (ttx * tty * ups < 64) { // a box (tile x segment) should fill a warp (measured: two warps' worth of columns is best)
            if (2 * ttx <= tty) ttx *= 2;
            else tty *= 2;
          }
        }
        ttx = std::min(ttx, K1B_MAXC); tty = std::min(tty, K1B_MAXC);
        // nodes per thread: pairs (measured: quads -- half the index arithmetic per node -- were 20-45 % slower: too
        // few items per tile to keep the lanes busy); MCT_K1_NPT=4 selects quads where the segment length allows
        int npt = 2;
        if (const char* v = getenv("MCT_K1_NPT")) { if (atoi(v) == 4 && seglen % 4 == 0) npt = 4; } // experiments
        K1BGeom G;
        G.seglen = seglen; G.nseg = (P.wz + seglen - 1) / seglen;
        G.nunits = (P.wz + npt) / npt; G.ups = seglen / npt; G.ngroups = (G.nunits + G.ups - 1) / G.ups;
        while ((long long)G.ngroups * ttx * tty * G.ups >= 65536 && tty > 1) tty /= 2; // 16-bit item index (multiply-shift division)
        while ((long long)((P.wx + ttx - 1) / ttx) * ((P.wy + tty - 1) / tty) >= 65536 && (ttx < K1B_MAXC || tty < K1B_MAXC)) { // 16-bit tile index
          if (ttx < tty) ttx *= 2; else tty *= 2;
        }
        G.ttx = ttx; G.tty = tty;
        G.tiles_x = (P.wx + ttx - 1) / ttx; G.tiles_y = (P.wy + tty - 1) / tty;
        G.tiles_per_model = G.tiles_x * G.tiles_y;
        const long long total_tiles = (long long)G.tiles_per_model * nby;
        if (total_tiles >= 65536 && nby > 1) { // k1t_div(t, tiles_per_model) needs t < 2^16: batches of such size go model by model
          for (int b = 0; b < nbatch; ++b) {
            int rc = launch_k1(gr, w, pm, d_vp + (long long)b * model_stride, d_vs + (long long)b * model_stride, d_rho + (long long)b * model_stride,
                               d_sites + (long long)b * model_stride, ia0, ja0, ka0, ny_a, nz_a, st, b);
            if (rc) return rc;
          }
          return MCT_OK;
        }
        G.total_tiles = (int)total_tiles;
        G.per_group = ttx * tty * G.ups;
        G.near_r = (float)(1.25 * h);
        G.dv_group = k1t_magic(G.per_group); G.dv_ups = k1t_magic(G.ups); G.dv_ty = k1t_magic(tty); G.dv_seg = k1t_magic(seglen);
        G.dv_tpm = k1t_magic(G.tiles_per_model); G.dv_tiles_y = k1t_magic(G.tiles_y);
        // one block per (model, tile) up to 48 per SM: tile costs differ, the hardware block scheduler evens them out
        // (measured: 6 persistent blocks per SM striding over the tiles were 14 % slower on C2 x 32)
        long long cap = (long long)g.sm_count * 48;
        if (const char* v = getenv("MCT_K1_BPS")) { if (atoi(v) > 0) cap = (long long)g.sm_count * atoi(v); } // experiments
        const long long rounds = (total_tiles + cap - 1) / cap;
        const int blocks = (int)((total_tiles + rounds - 1) / rounds);
        const int threads = (long long)G.ngroups * G.per_group <= 1024 ? 128 : 256;
        if (npt == 4) k1_box_kernel<4><<<blocks, threads, 0, st>>>(P, G);
        else k1_box_kernel<2><<<blocks, threads, 0, st>>>(P, G);
      }
    }
  }
  CK(cudaGetLastError());
  g.host_stats.n_nodes += total * (batched ? nbatch : 1);
  g.host_stats.n_launches += 1;
  return MCT_OK;
}

struct DispPlan {
  int ix0, ix1, iy0, iy1, wx, wy, ncol, stride, nm, nout;
  int nb = 1;  // models stacked along the slowest axis; ncol counts all of them
  int cpm = 0; // columns per model
};

int plan_disp(const mct_grid* gr, int ix0, int ix1, int iy0, int iy1, int np, const mct_disp_opts* opt, DispPlan& pl, int nb = 1) {
  pl.nb = nb;
  if (nb < 1) return fail(MCT_E_INVALID_ARG, "dispersion: empty batch");
  if (!grid_ok(gr) || !opt) return fail(MCT_E_INVALID_ARG, "dispersion: bad grid or NULL options");
  if (np < 1 || np > MCT_MAX_PERIODS) return fail(MCT_E_INVALID_ARG, "dispersion: np must be in 1..%d", MCT_MAX_PERIODS);
  if (ix0 < 1 || iy0 < 1 || ix1 > gr->nx || iy1 > gr->ny || ix1 < ix0 || iy1 < iy0)
    return fail(MCT_E_INVALID_ARG, "dispersion: window %d..%d x %d..%d outside the %d x %d grid", ix0, ix1, iy0, iy1, gr->nx, gr->ny);
  if (opt->raylov != 0 && opt->raylov != 1) return fail(MCT_E_INVALID_ARG, "dispersion: raylov must be 0 (Love) or 1 (Rayleigh)");
  if (opt->nmodes > 1000) return fail(MCT_E_INVALID_ARG, "dispersion: nmodes must not exceed 1000 (the mode counter of the search state is 16 bits)");
  if (!(gr->scaling != 0.0)) return fail(MCT_E_INVALID_ARG, "dispersion: grid scaling must be non-zero");
  pl.ix0 = ix0; pl.ix1 = ix1; pl.iy0 = iy0; pl.iy1 = iy1;
  pl.wx = ix1 - ix0 + 1; pl.wy = iy1 - iy0 + 1;
  pl.cpm = pl.wx * pl.wy;
  pl.ncol = pl.cpm * pl.nb;
  pl.stride = (pl.ncol + 31) & ~31;
  pl.nm = opt->nmodes <= 0 ? 1 : opt->nmodes;
  pl.nout = np * pl.nm;
  return MCT_OK;
}

int launch_k2(int ncol, int stride, const double* freqs, int np, const mct_disp_opts* opt, double* d_pvel, double* d_gvel,
              int32_t* d_ierr, const int32_t* d_skip, int cols_per_model, int max_layers, cudaStream_t st) {
  K2Params P;
  P.lay = (const float4*)g.lay.p;
  P.nlay = (const int32_t*)g.nlay.p;
  P.status = (const int32_t*)g.status.p;
  P.ncol = ncol;
  P.stride = stride;
  P.kmax = np;
  P.nmode = opt->nmodes <= 0 ? 1 : opt->nmodes;
  P.mmode = opt->nmodes <= 0 ? 0 : 1;
  P.ifunc = opt->raylov == 1 ? 2 : 1; // surfmodes.f90:82,94: iwave 2 Rayleigh, 1 Love
  P.igr = opt->phaseGroup;
  P.count = g.count_on;
  P.ddc0 = (float)opt->dphase; // ddc0 = dphase (surfdisp96.f:130)
  P.dc = fabs((double)P.ddc0);
  P.preset_unsolved = opt->preset;
  P.pvel = d_pvel;
  P.gvel = d_gvel;
  P.ierr = d_ierr;
  P.skip = d_skip;
  P.cols_per_model = cols_per_model > 0 ? cols_per_model : ncol;
  P.counters = (unsigned long long*)g.counters.p;
  for (int i = 0; i < MCT_MAX_PERIODS; ++i) P.t[i] = (i < np) ? 1 / freqs[i] : 0.0; // dble(1/freqs), surfmodes.f90:82
  // Large batches get COMPACT inputs (k2_dedup.cuh): the columns to solve are gathered into sorted-list order after the
  // sort, and the reciprocal table is then built for those only.  Otherwise: reciprocals for every column, in place.
  const bool want_compact = g.compact_inputs && ncol >= 8192 && g.k2_variant == 7;
  P.compact = 0;
  if (!want_compact) {
    int rc;
    if ((rc = ensure(g.layr, sizeof(double4) * (size_t)stride * (size_t)max_layers))) return rc;
    ProfScope ps(2, st);
    layer_recips_kernel<<<(ncol + 127) / 128, 128, 0, st>>>(P.lay, P.nlay, ncol, stride, P.ifunc, (double4*)g.layr.p);
    g.host_stats.n_launches += 1;
    P.layr = (const double4*)g.layr.p;
  }
  CK(cudaGetLastError());
  // Fold bit-identical layer stacks (k2_dedup.cuh): K2 then solves one representative per group.
  const int32_t* sort_key = P.nlay;
  const int32_t* d_rep = nullptr;
  int neff = ncol; // columns K2 has to visit; known on the host only for large batches (see below)
  P.mult = nullptr;
  if (g.dedup && ncol > 1) {
    int rc;
    unsigned nslots = 64;
    while (nslots < 2u * (unsigned)ncol) nslots <<= 1;
    if ((rc = ensure(g.dd_table, sizeof(unsigned long long) * (size_t)nslots))) return rc;
    if ((rc = ensure(g.dd_i32, sizeof(int32_t) * (7 * (size_t)stride + 8)))) return rc;
    int32_t* base = (int32_t*)g.dd_i32.p;
    int32_t *rep0 = base, *minrep = base + stride, *mult0 = base + 2 * (size_t)stride, *rep = base + 3 * (size_t)stride,
            *mult = base + 4 * (size_t)stride, *kstat = base + 5 * (size_t)stride, *skey = base + 6 * (size_t)stride,
            *d_neff = base + 7 * (size_t)stride;
    ProfScope ps(2, st);
    CK(cudaMemsetAsync(g.dd_table.p, 0, sizeof(unsigned long long) * (size_t)nslots, st));
    CK(cudaMemsetAsync(mult0, 0, sizeof(int32_t) * (size_t)stride, st));
    CK(cudaMemsetAsync(d_neff, 0, sizeof(int32_t), st));
    fill_i32_kernel<<<grid_blocks(ncol, 256, 8), 256, 0, st>>>(minrep, ncol, 0x7fffffff);
    dedup_insert_kernel<<<(ncol + 127) / 128, 128, 0, st>>>(P.lay, P.nlay, P.status, ncol, stride, P.cols_per_model,
                                                           (unsigned long long*)g.dd_table.p, nslots, rep0);
    dedup_group_kernel<<<(ncol + 255) / 256, 256, 0, st>>>(rep0, ncol, minrep, mult0);
    dedup_finish_kernel<<<(ncol + 255) / 256, 256, 0, st>>>(rep0, minrep, mult0, P.status, P.nlay, ncol, rep, mult, kstat, skey, d_neff,
                                                           P.lay, stride, g.sort_proxy);
    CK(cudaGetLastError());
    g.host_stats.n_launches += 4;
    P.status = kstat;
    P.mult = mult;
    sort_key = skey;
    d_rep = rep;
    // Large batches: the launch shape below depends on how many columns are really solved, and K2 then runs for tens
    // of milliseconds -- one 4-byte read-back (a stream synchronisation) is noise.  Proposal-sized calls stay fully
    // asynchronous: their shape is chosen from ncol, and the duplicates' blocks exit at once.
    if (ncol >= 8192 || g.shard_n > 1) {
      int32_t h = ncol;
      CK(cudaMemcpyAsync(&h, d_neff, sizeof h, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      neff = h < 1 ? 1 : h;
    }
  }
  // Order the columns by layer count (descending) so every warp runs one layer-loop trip count; duplicates (key 0) last.
  const int variant = g.k2_variant;
  P.perm = nullptr;
  if (variant != 0 || d_rep) {
    int rc;
    if ((rc = ensure(g.perm, sizeof(int32_t) * (size_t)stride))) return rc;
    const int nb256 = (ncol + 255) / 256;
    if ((rc = ensure(g.bins, sizeof(int32_t) * 256 * ((size_t)nb256 + 1)))) return rc;
    ProfScope ps(2, st);
    if (g.sort_stable) { // deterministic, keeps spatial neighbours together inside a bin
      int32_t* bin_base = (int32_t*)g.bins.p;
      int32_t* blk = bin_base + 256;
      ssort_hist_kernel<<<nb256, SORT_BLOCK, 0, st>>>(sort_key, ncol, blk);
      ssort_offsets_kernel<<<1, 256, 0, st>>>(blk, nb256, bin_base);
      ssort_scatter_kernel<<<nb256, SORT_BLOCK, 0, st>>>(sort_key, ncol, blk, bin_base, (int32_t*)g.perm.p);
    } else {
      CK(cudaMemsetAsync(g.bins.p, 0, sizeof(int32_t) * 256, st));
      sort_hist_kernel<<<nb256, 256, 0, st>>>(sort_key, ncol, (int32_t*)g.bins.p);
      sort_scan_kernel<<<1, 32, 0, st>>>((int32_t*)g.bins.p);
      sort_scatter_kernel<<<nb256, 256, 0, st>>>(sort_key, ncol, (int32_t*)g.bins.p, (int32_t*)g.perm.p);
    }
    g.host_stats.n_launches += 3;
    P.perm = (const int32_t*)g.perm.p;
  }
  CK(cudaGetLastError());
  const int ncol_all = ncol;
  ncol = neff;    // the first neff entries of perm are the columns to visit (all of them when neff was not read back)
  if (g.shard_n > 1) { // balanced sharding: this rank's share of the sorted list (longest first: every rank gets the same mix)
    if (!P.perm) return fail(MCT_E_INVALID_ARG, "sharded dispersion needs the column sort (MCT_K2_VARIANT=0 disables it)");
    // chunks of 32 list entries are dealt round-robin; the last chunk may be short
    const int nchunks = (neff + SHARD_CHUNK - 1) / SHARD_CHUNK;
    const int mychunks = nchunks > g.shard_r ? (nchunks - g.shard_r + g.shard_n - 1) / g.shard_n : 0;
    int nmine = mychunks * SHARD_CHUNK;
    if (mychunks > 0 && shard_global(nmine - 1, g.shard_r, g.shard_n) >= neff) // my last chunk is the list's (short) last one
      nmine -= nchunks * SHARD_CHUNK - neff;
    int rc;
    if ((rc = ensure(g.sh_perm, sizeof(int32_t) * (size_t)std::max(nmine, 1)))) return rc;
    if (nmine > 0) shard_pick_kernel<<<(nmine + 255) / 256, 256, 0, st>>>((const int32_t*)g.perm.p, neff, g.shard_r, g.shard_n, nmine, (int32_t*)g.sh_perm.p);
    CK(cudaGetLastError());
    g.host_stats.n_launches += 1;
    P.perm = (const int32_t*)g.sh_perm.p;
    ncol = nmine;
  }
  P.ncol = ncol;
  if (want_compact && ncol > 0 && P.perm) {
    const int stride_c = (ncol + 31) & ~31;
    int rc;
    if ((rc = ensure(g.layc, sizeof(float4) * (size_t)stride_c * (size_t)max_layers)) ||
        (rc = ensure(g.layrc, sizeof(double4) * (size_t)stride_c * (size_t)max_layers)) ||
        (rc = ensure(g.nlayc, sizeof(int32_t) * 2 * (size_t)stride_c)))
      return rc;
    int32_t* nlayc = (int32_t*)g.nlayc.p;
    int32_t* statc = nlayc + stride_c;
    ProfScope ps(2, st);
    compact_layers_kernel<<<(ncol + 127) / 128, 128, 0, st>>>(P.lay, P.nlay, P.status, P.perm, ncol, stride, stride_c, (float4*)g.layc.p,
                                                             nlayc, statc);
    layer_recips_kernel<<<(ncol + 127) / 128, 128, 0, st>>>((const float4*)g.layc.p, nlayc, ncol, stride_c, P.ifunc, (double4*)g.layrc.p);
    CK(cudaGetLastError());
    g.host_stats.n_launches += 2;
    P.lay = (const float4*)g.layc.p;
    P.layr = (const double4*)g.layrc.p;
    P.nlay = nlayc;
    P.status = statc;
    P.stride = stride_c;
    P.compact = 1;
  } else if (want_compact) { // nothing to gather from (no sort): the plain path after all
    int rc;
    if ((rc = ensure(g.layr, sizeof(double4) * (size_t)stride * (size_t)max_layers))) return rc;
    ProfScope ps(2, st);
    layer_recips_kernel<<<(ncol_all + 127) / 128, 128, 0, st>>>(P.lay, P.nlay, ncol_all, stride, P.ifunc, (double4*)g.layr.p);
    g.host_stats.n_launches += 1;
    P.layr = (const double4*)g.layr.p;
  }
  const char* kname = "";
  int lanes = 1;
  if (ncol > 0) {
    ProfScope ps(1, st);
    const int nw = (ncol + 31) / 32;
    // Batches that cannot fill the GPU with one thread per column get G lanes per column: the largest power of
    // two, up to a whole warp, that keeps the launch within about two waves of resident lanes.
    const long long capacity = (long long)g.sm_count * 16 * 32; // resident lanes of the dispersion kernels
    // Hand-over to one thread per column: two lanes per column stay ahead until ~1.7x the resident lanes (65 536
    // columns: 78 ms against 94 ms), and are level with it at 131 072 (tools/lanes_sweep.py X huge).
    const int coop_max = g.k2_coop_max > 0 ? g.k2_coop_max : (int)std::min<long long>(capacity * 17 / 10, 1 << 30);
    const bool coop = (g.k2_mode == 2) || (g.k2_mode == 0 && ncol < coop_max);
    if (coop) {
      int G = 32;
      if (g.k2_coop_lanes > 0) G = g.k2_coop_lanes;
      else while (G > 2 && (long long)ncol * G * 5 > capacity * 11) G /= 2; // largest G within 2.2x the resident lanes:
      // measured optimum (tools/lanes_sweep.py big): shorter serial chains beat a fully resident grid up to ~2 waves
      // The smallest batches get several warps per column (one block per column, its warps on different SM
      // sub-partitions): 4 up to a fifth of the GPU's resident warps, 2 up to a third (tools/lanes_sweep.py: 400 columns
      // 1.39 ms with 4 warps against 1.80 with one; 576: 1.54 with 2; 1024: one warp is best; 8 warps never pay off
      // since the warps of a block are all busy during nevill's bisections too).
      if (g.k2_coop_lanes == 0 && G == 32) {
        const long long slots = (long long)g.sm_count * 16; // resident warps
        if ((long long)ncol * 5 <= slots) G = 128;
        else if ((long long)ncol * 3 <= slots) G = 64;
      }
      lanes = G;
      const int nblk = G > 32 ? ncol : (ncol + (32 / G) - 1) / (32 / G);
      switch (G) {
        case 64: k2_coopw2_kernel<<<nblk, 64, 0, st>>>(P); kname = "k2_coopw2_kernel"; break;
        case 128: k2_coopw4_kernel<<<nblk, 128, 0, st>>>(P); kname = "k2_coopw4_kernel"; break;
        case 256: k2_coopw8_kernel<<<nblk, 256, 0, st>>>(P); kname = "k2_coopw8_kernel"; break;
        case 2: k2_coop2_kernel<<<nblk, 32, 0, st>>>(P); kname = "k2_coop2_kernel"; break;
        case 4: k2_coop4_kernel<<<nblk, 32, 0, st>>>(P); kname = "k2_coop4_kernel"; break;
        case 8: k2_coop8_kernel<<<nblk, 32, 0, st>>>(P); kname = "k2_coop8_kernel"; break;
        case 16: k2_coop16_kernel<<<nblk, 32, 0, st>>>(P); kname = "k2_coop16_kernel"; break;
        default: k2_coop_kernel<<<nblk, 32, 0, st>>>(P); kname = "k2_coop_kernel"; break;
      }
    }
    else if (variant == 3 || variant == 0) { k2_dispersion_plain<<<nw, 32, 0, st>>>(P); kname = "k2_dispersion_plain"; } // A/B reference (0: unsorted too)
    else {
      k2_dispersion_fast_r128<<<nw, 32, 0, st>>>(P);
      kname = "k2_dispersion_fast_r128";
    }
  }
  CK(cudaGetLastError());
  // what this launch did, for mct_last_launch (the bench's roofline block names the kernel from here, not from a
  // re-statement of the rule above)
  snprintf(g.last_kernel, sizeof g.last_kernel, "%s", kname);
  g.last_ncol = ncol_all;
  g.last_neff = (g.dedup && ncol_all >= 8192) ? neff : -1;
  g.last_lanes = lanes;
  if (g.shard_n > 1) { // exchange the representatives' results: pack, in-place all-gather, unpack (mct_comm.cuh)
    int rc = shard_exchange(neff, P.kmax * P.nmode, P.igr > 0, P.pvel, P.gvel, P.ierr, st);
    if (rc) return rc;
  }
  if (d_rep) {
    ProfScope ps(2, st);
    dedup_scatter_kernel<<<grid_blocks((long long)ncol_all * P.kmax * P.nmode, 256, 16), 256, 0, st>>>(
        d_rep, ncol_all, P.kmax * P.nmode, P.skip, P.cols_per_model, P.pvel, P.gvel, P.ierr);
    CK(cudaGetLastError());
    g.host_stats.n_launches += 1;
  }
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  return MCT_OK;
}


// Low-velocity columns (status 2) after the dispersion kernel: the generalized R/T search of surfmodes.f90:84-87,96-99
// (k5_grt.cuh).  lp: the layering parameters of the call (model-column form) or null with the pre-layered buffers set.
int launch_grt(int ncol, const LayParams* lp, const PreLayParams* pp, const double* freqs, int np, const mct_disp_opts* opt, double* d_pvel,
               double* d_gvel, int32_t* d_ierr, const int32_t* d_skip, int32_t* d_flags, int cols_per_model, cudaStream_t st) {
  g.grt_last_cols = 0;
  if (!g.grt_on || opt->nmodes > 0 || g.shard_n > 1 || ncol < 1) return MCT_OK; // surfmmodes has no such branch (surfmodes.f90:153,165)
  int rc;
  if ((rc = ensure(g.grt_list, sizeof(int32_t) * ((size_t)ncol + 2)))) return rc;
  int32_t* list = (int32_t*)g.grt_list.p;
  int32_t* d_count = list + ncol; // [0] columns listed, [1] the persistent kernel's work counter
  CK(cudaMemsetAsync(d_count, 0, 2 * sizeof(int32_t), st));
  grt_collect_kernel<<<(ncol + 255) / 256, 256, 0, st>>>((const int32_t*)g.status.p, ncol, list, d_count);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  int32_t count = 0;
  CK(cudaMemcpyAsync(&count, d_count, sizeof count, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (count <= 0) return MCT_OK;
  // persistent one-warp blocks, as many as the GPU holds at 218 registers per thread (8 per SM); each owns one scratch slot
  const int blocks = std::min<int>(count, g.sm_count * 8);
  if ((rc = ensure(g.grt_scratch, sizeof(double) * (size_t)GRT_SCRATCH * (size_t)blocks))) return rc;
  GrtParams P;
  memset(&P, 0, sizeof P);
  if (lp) {
    P.vp = lp->vp; P.vs = lp->vs; P.rho = lp->rho;
    P.ny = lp->ny; P.nz = lp->nz; P.ix0 = lp->ix0; P.iy0 = lp->iy0; P.wx = lp->wx; P.wy = lp->wy;
    P.model_stride = lp->model_stride;
    P.dz = lp->dz; P.waterDepth = lp->waterDepth; P.scaling = lp->scaling; P.layer_eps = lp->layer_eps; P.water_thresh = lp->water_thresh;
  } else {
    P.pl_thick = pp->thick; P.pl_vp = pp->vp; P.pl_vs = pp->vs; P.pl_rho = pp->rho; P.pl_off = pp->offsets;
  }
  P.modetype = opt->raylov; P.phaseGroup = opt->phaseGroup; P.np = np;
  P.dc = opt->dphase;
  P.tolmin = g.grt_par[0]; P.tolmax = g.grt_par[1]; P.smin_min = g.grt_par[2]; P.smin_max = g.grt_par[3]; P.dcm = g.grt_par[4]; P.dc2 = g.grt_par[5];
  for (int i = 0; i < MCT_MAX_PERIODS; ++i) P.freqs[i] = i < np ? freqs[i] : 0.0;
  P.list = list; P.nlist = count; P.next = d_count + 1;
  P.skip = d_skip; P.cols_per_model = cols_per_model > 0 ? cols_per_model : ncol;
  P.scratch = (double*)g.grt_scratch.p;
  P.pvel = d_pvel; P.gvel = d_gvel; P.ierr = d_ierr;
  P.counters = g.count_on ? (unsigned long long*)g.counters.p + 8 : nullptr;
  P.flags = d_flags;
  {
    ProfScope ps(2, st);
    grt_kernel<<<blocks, 32, 0, st>>>(P);
    g.host_stats.n_launches += 1;
  }
  CK(cudaGetLastError());
  g.grt_last_cols = count;
  return MCT_OK;
}

// check_model + layerize + K2 on device-resident whole-grid arrays (pl.nb models stacked along x).
// d_flags: int32[2*nb]: per model {model_invalid, max condition code}.
int disp_core(const double* d_vp, const double* d_vs, const double* d_rho, const mct_grid* gr, const DispPlan& pl,
              const double* freqs, int np, const mct_disp_opts* opt, bool do_check, const int cw[4] /* check_model region:
              ix0, iy0, wx, wy */, double* d_pvel, double* d_gvel, int32_t* d_ierr, int32_t* d_flags, cudaStream_t st) {
  int rc;
  const long long model_stride = (long long)gr->nx * gr->ny * gr->nz;
  if ((rc = ensure(g.lay, sizeof(float4) * (size_t)pl.stride * (size_t)(gr->nz + 1)))) return rc;
  if ((rc = ensure(g.nlay, sizeof(int32_t) * (size_t)pl.stride))) return rc;
  if ((rc = ensure(g.status, sizeof(int32_t) * (size_t)pl.stride))) return rc;
  CK(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int32_t) * (size_t)pl.nb, st));
  if (do_check) {
    ProfScope ps(2, st);
    check_model_kernel<<<grid_blocks((long long)cw[2] * cw[3] * pl.nb * 32, 256, 8), 256, 0, st>>>(
        d_vs, cw[0], cw[1], cw[2], cw[3], gr->ny, pl.nb, model_stride, gr->nz, d_flags);
    g.host_stats.n_launches += 1;
  }
  CK(cudaGetLastError());
  LayParams L;
  L.vp = d_vp; L.vs = d_vs; L.rho = d_rho;
  L.ny = gr->ny; L.nz = gr->nz;
  L.ix0 = pl.ix0; L.iy0 = pl.iy0; L.wx = pl.wx; L.wy = pl.wy;
  L.nmodels = pl.nb; L.model_stride = model_stride;
  L.dz = gr->dz; L.waterDepth = gr->waterDepth; L.scaling = gr->scaling;
  L.layer_eps = opt->layer_eps; L.water_thresh = opt->water_thresh;
  L.modetype = opt->raylov;
  L.lay = (float4*)g.lay.p; L.nlay = (int32_t*)g.nlay.p; L.status = (int32_t*)g.status.p;
  L.stride = pl.stride;
  L.flags = d_flags;
  L.grt_on = (g.grt_on && opt->nmodes <= 0 && g.shard_n <= 1) ? 1 : 0;
  {
    ProfScope ps(2, st);
    layerize_kernel<<<(pl.ncol + 127) / 128, 128, 0, st>>>(L);
  }
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  if ((rc = launch_k2(pl.ncol, pl.stride, freqs, np, opt, d_pvel, d_gvel, d_ierr, do_check ? d_flags : nullptr, pl.cpm, gr->nz + 1, st))) return rc;
  return launch_grt(pl.ncol, &L, nullptr, freqs, np, opt, d_pvel, d_gvel, d_ierr, do_check ? d_flags : nullptr, d_flags, pl.cpm, st);
}

// K1 per model + property maps + disp_core over the x-slab ixs0..ixs1 of every model of the resident set.
int forward_core(const mct_grid* gr, int nb, int derive_vp_rho, const DispPlan& pl, const double* freqs, int np,
                 const mct_disp_opts* opt, double* d_vp, double* d_vs, double* d_rho, int32_t* d_sites, double* d_pvel,
                 double* d_gvel, int32_t* d_ierr, int32_t* d_flags, cudaStream_t st) {
  if (nb != g.nset) return fail(MCT_E_INVALID_ARG, "forward: batch of %d models but %d nuclei sets are resident", nb, g.nset);
  const int32_t w[6] = {pl.ix0, pl.ix1, 1, gr->ny, 1, gr->nz};
  const size_t slab = (size_t)gr->ny * gr->nz;
  const size_t model_stride = slab * (size_t)gr->nx;
  const size_t xoff = (size_t)(pl.ix0 - 1) * slab;
  int rc;
  CK(cudaMemsetAsync((int32_t*)g.flags.p + 2, 0, sizeof(int32_t), st));
  if (g.k1_mode != 1) { // one launch for the whole batch (blockIdx.y = model)
    if ((rc = launch_k1(gr, w, nullptr, d_vp, d_vs, d_rho, d_sites, 1, 1, 1, gr->ny, gr->nz, st, 0, nb, (long long)model_stride))) return rc;
  } else {
    for (int b = 0; b < nb; ++b) {
      const size_t mo = (size_t)b * model_stride;
      if ((rc = launch_k1(gr, w, nullptr, d_vp + mo, d_vs + mo, d_rho + mo, d_sites + mo, 1, 1, 1, gr->ny, gr->nz, st, b))) return rc;
    }
  }
  if (derive_vp_rho) {
    ProfScope ps(2, st);
    if (pl.wx == gr->nx) {
      if ((rc = mct_vs2vp_rho_dev(d_vs, d_vp, d_rho, (int64_t)(model_stride * (size_t)nb), st))) return rc;
    } else {
      for (int b = 0; b < nb; ++b) {
        const size_t o = (size_t)b * model_stride + xoff;
        if ((rc = mct_vs2vp_rho_dev(d_vs + o, d_vp + o, d_rho + o, (int64_t)((size_t)pl.wx * slab), st))) return rc;
      }
    }
  }
  const int cw[4] = {pl.ix0, 1, pl.wx, gr->ny}; // check_model over this rank's slab (whole grid when unsharded)
  return disp_core(d_vp, d_vs, d_rho, gr, pl, freqs, np, opt, true, cw, d_pvel, d_gvel, d_ierr, d_flags, st);
}

void release_misfit_globals(); // k4_misfit.cuh
void release_fm2d_globals();   // k6_fm2d.cuh
void comm_release();           // mct_comm.cuh
int shard_exchange(int neff, int nout, bool with_group, double* d_pvel, double* d_gvel, int32_t* d_ierr, cudaStream_t st); // mct_comm.cuh

int flags_to_code(int maxst) {
  return maxst == 2 ? MCT_E_GRT_NEEDED : (maxst == 3 ? MCT_E_TOO_MANY_LAYERS : MCT_E_FLUID_BELOW_TOP);
}

// ---- FP64 peak probe: independent DFMA chains, 8 per thread ---------------------------------------
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, int fused) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  if (fused) {
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
      a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c); a2 = __dadd_rn(__dmul_rn(a2, m), c);
      a3 = __dadd_rn(__dmul_rn(a3, m), c); a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
      a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace

extern "C" {

const char* mct_last_error(void) { return g.err; }

int mct_init(int device) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (g.init) {
    if (device == g.device) return MCT_OK;
    return fail(MCT_E_INVALID_ARG, "already initialised on device %d", g.device);
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(MCT_E_CUDA, "no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(MCT_E_INVALID_ARG, "device %d out of range (0..%d)", device, n - 1);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  g.sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  g.device = device;
  int rc;
  if ((rc = ensure(g.flags, 4 * sizeof(int32_t)))) return rc;
  if ((rc = ensure(g.counters, 16 * sizeof(unsigned long long)))) return rc;
  CK(cudaMemset(g.flags.p, 0, 4 * sizeof(int32_t)));
  CK(cudaMemset(g.counters.p, 0, 16 * sizeof(unsigned long long)));
  g.host_stats = mct_stats{0, 0, 0, 0, 0, 0, 0, 0};
  if (const char* v = getenv("MCT_K2_VARIANT")) g.k2_variant = atoi(v);
  if (const char* v = getenv("MCT_SORT_STABLE")) g.sort_stable = atoi(v);
  if (const char* v = getenv("MCT_DEDUP")) g.dedup = atoi(v) ? 1 : 0;
  if (const char* v = getenv("MCT_SORT_PROXY")) g.sort_proxy = (float)atof(v);
  if (const char* v = getenv("MCT_COMPACT")) g.compact_inputs = atoi(v) ? 1 : 0;
  if (const char* v = getenv("MCT_K2_COOP_LANES")) { // experiments only; same validation as mct_set_k2_lanes
    const int l = atoi(v);
    if (l == 0 || (l >= 2 && l <= 256 && (l & (l - 1)) == 0)) g.k2_coop_lanes = l;
  }
  g.init = true;
  return MCT_OK;
}

int mct_shutdown(void) {
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (!g.init) return MCT_OK;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  DevBuf* bufs[] = {&g.nodes, &g.rpts, &g.ind, &g.params, &g.kmodels, &g.m_vp, &g.m_vs, &g.m_rho, &g.m_sites, &g.lay, &g.layr, &g.nlay, &g.status, &g.perm, &g.bins, &g.dd_table, &g.dd_i32, &g.layc, &g.layrc, &g.nlayc, &g.sh_perm, &g.sh_p, &g.sh_g, &g.sh_i,
                    &g.o_pvel, &g.o_gvel, &g.o_ierr, &g.bflags, &g.pl_thick, &g.pl_vp, &g.pl_vs, &g.pl_rho, &g.pl_off, &g.flags, &g.counters, &g.grt_list, &g.grt_scratch, &g.ray_pts, &g.ray_off, &g.ray_time};
  for (DevBuf* b : bufs) release(*b);
  release_misfit_globals();
  release_fm2d_globals();
  comm_release();
  release(g.pin_a);
  release(g.pin_b);
  release(g.pin_small);
  if (g.stage_ev) cudaEventDestroy(g.stage_ev);
  g.stage_ev = nullptr;
  g.k1_smem_set = false;
  cudaStreamDestroy(g.stream);
  g.stream = nullptr;
  g.init = false;
  return MCT_OK;
}

int mct_set_counters(int on) {
  g.count_on = on ? 1 : 0;
  return MCT_OK;
}

int mct_reset_stats(void) {
  NEED_INIT();
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaMemset(g.counters.p, 0, 16 * sizeof(unsigned long long)));
  g.host_stats = mct_stats{0, 0, 0, 0, 0, 0, 0, 0};
  return MCT_OK;
}

int mct_get_stats(mct_stats* out) {
  NEED_INIT();
  if (!out) return fail(MCT_E_INVALID_ARG, "NULL stats pointer");
  unsigned long long c[8];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(c, g.counters.p, sizeof c, cudaMemcpyDeviceToHost));
  *out = g.host_stats;
  out->n_dltar = (int64_t)c[0];
  out->n_layer_steps = (int64_t)c[1];
  out->n_columns = (int64_t)c[2];
  out->n_dltar_executed = (int64_t)c[4];
  out->n_layer_steps_executed = (int64_t)c[5];
  out->n_columns_solved = (int64_t)c[6];
  return MCT_OK;
}

int mct_box_window(const mct_grid* gr, const double box[6], int32_t w[6]) {
  if (!grid_ok(gr) || !box || !w) return fail(MCT_E_INVALID_ARG, "box_window: bad arguments");
  box_window(gr, box, w);
  return MCT_OK;
}

int mct_voronoi_to_grid_dev(const double* points, const double* params, int ncells, const mct_grid* gr, const double box[6],
                            const double* pm, double* d_vp, double* d_vs, double* d_rho, int32_t* d_sites_id, void* stream) {
  NEED_INIT();
  if (!grid_ok(gr) || !box || !d_vp || !d_vs || !d_rho || !d_sites_id) return fail(MCT_E_INVALID_ARG, "voronoi_to_grid: bad arguments");
  cudaStream_t st = pick(stream);
  int rc = upload_nuclei(points, params, ncells, st);
  if (rc) return rc;
  int32_t w[6];
  box_window(gr, box, w);
  CK(cudaMemsetAsync((int32_t*)g.flags.p + 2, 0, sizeof(int32_t), st)); // read back by mct_k1_status()
  return launch_k1(gr, w, pm, d_vp, d_vs, d_rho, d_sites_id, 1, 1, 1, gr->ny, gr->nz, st);
}

int mct_k1_status(void* stream) {
  NEED_INIT();
  cudaStream_t st = pick(stream);
  int32_t k1err = 0;
  CK(cudaMemcpyAsync(&k1err, (int32_t*)g.flags.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (k1err) return fail(MCT_E_CUDA, "nearest-nucleus traversal stack overflow (tree deeper than %d)", K1_STACK);
  return MCT_OK;
}

int mct_voronoi_to_grid(const double* points, const double* params, int ncells, const mct_grid* gr, const double box[6],
                        const double* pm, double* vp, double* vs, double* rho, int32_t* sites_id) {
  NEED_INIT();
  if (!grid_ok(gr) || !box || !vp || !vs || !rho || !sites_id) return fail(MCT_E_INVALID_ARG, "voronoi_to_grid: bad arguments");
  cudaStream_t st = g.stream;
  int rc = upload_nuclei(points, params, ncells, st);
  if (rc) return rc;
  int32_t w[6];
  box_window(gr, box, w);
  const int wx = w[1] - w[0] + 1, wy = w[3] - w[2] + 1, wz = w[5] - w[4] + 1;
  if (wx <= 0 || wy <= 0 || wz <= 0) return MCT_OK;
  // The window is staged as a packed (wz,wy,wx) block: gather (pm mode only) -> H2D -> K1 -> D2H -> scatter.
  const size_t nn = (size_t)wx * wy * wz;
  if ((rc = ensure(g.m_vp, nn * 8))) return rc;
  if ((rc = ensure(g.m_vs, nn * 8))) return rc;
  if ((rc = ensure(g.m_rho, nn * 8))) return rc;
  if ((rc = ensure(g.m_sites, nn * 4))) return rc;
  if ((rc = ensure_pin(g.pin_a, nn * 28))) return rc;
  double* h_vp = (double*)g.pin_a.p;
  double* h_vs = h_vp + nn;
  double* h_rho = h_vs + nn;
  int32_t* h_sid = (int32_t*)(h_rho + nn);
  const size_t ny = gr->ny, nz = gr->nz;
  const bool full_z = (wz == gr->nz);
  auto for_rows = [&](auto&& fn) { // fn(src offset in the user arrays, dst offset in the packed block, run length)
    for (int i = 0; i < wx; ++i)
      for (int j = 0; j < wy; ++j) {
        const size_t uo = ((size_t)(w[0] - 1 + i) * ny + (size_t)(w[2] - 1 + j)) * nz + (size_t)(w[4] - 1);
        const size_t po = ((size_t)i * wy + j) * wz;
        fn(uo, po, (size_t)wz);
      }
  };
  (void)full_z;
  if (pm) {
    for_rows([&](size_t uo, size_t po, size_t n) {
      memcpy(h_vp + po, vp + uo, n * 8);
      memcpy(h_vs + po, vs + uo, n * 8);
      memcpy(h_rho + po, rho + uo, n * 8);
      memcpy(h_sid + po, sites_id + uo, n * 4);
    });
    CK(cudaMemcpyAsync(g.m_vp.p, h_vp, nn * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.m_vs.p, h_vs, nn * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.m_rho.p, h_rho, nn * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.m_sites.p, h_sid, nn * 4, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync((int32_t*)g.flags.p + 2, 0, sizeof(int32_t), st));
  rc = launch_k1(gr, w, pm, (double*)g.m_vp.p, (double*)g.m_vs.p, (double*)g.m_rho.p, (int32_t*)g.m_sites.p, w[0], w[2], w[4], wy, wz, st);
  if (rc) return rc;
  int32_t k1err = 0;
  CK(cudaMemcpyAsync(h_vp, g.m_vp.p, nn * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_vs, g.m_vs.p, nn * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_rho, g.m_rho.p, nn * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_sid, g.m_sites.p, nn * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&k1err, (int32_t*)g.flags.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (k1err) return fail(MCT_E_CUDA, "nearest-nucleus traversal stack overflow (tree deeper than %d)", K1_STACK);
  for_rows([&](size_t uo, size_t po, size_t n) {
    memcpy(vp + uo, h_vp + po, n * 8);
    memcpy(vs + uo, h_vs + po, n * 8);
    memcpy(rho + uo, h_rho + po, n * 8);
    memcpy(sites_id + uo, h_sid + po, n * 4);
  });
  return MCT_OK;
}

int mct_vs2vp_rho_dev(const double* d_vs, double* d_vp, double* d_rho, int64_t n, void* stream) {
  NEED_INIT();
  if (!d_vs || !d_vp || !d_rho || n < 0) return fail(MCT_E_INVALID_ARG, "vs2vp_rho: bad arguments");
  if (n == 0) return MCT_OK;
  vs2vp_rho_kernel<<<grid_blocks(n, 256, 8), 256, 0, pick(stream)>>>(d_vs, d_vp, d_rho, (long long)n);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  return MCT_OK;
}

int mct_vs2vp_rho(const double* vs, double* vp, double* rho, int64_t n) {
  NEED_INIT();
  if (!vs || !vp || !rho || n < 0) return fail(MCT_E_INVALID_ARG, "vs2vp_rho: bad arguments");
  if (n == 0) return MCT_OK;
  int rc;
  if ((rc = ensure(g.m_vs, (size_t)n * 8))) return rc;
  if ((rc = ensure(g.m_vp, (size_t)n * 8))) return rc;
  if ((rc = ensure(g.m_rho, (size_t)n * 8))) return rc;
  cudaStream_t st = g.stream;
  CK(cudaMemcpyAsync(g.m_vs.p, vs, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  if ((rc = mct_vs2vp_rho_dev((double*)g.m_vs.p, (double*)g.m_vp.p, (double*)g.m_rho.p, n, st))) return rc;
  CK(cudaMemcpyAsync(vp, g.m_vp.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(rho, g.m_rho.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MCT_OK;
}

int mct_vs2vp_rho_window(const double* vs, double* vp, double* rho, const mct_grid* gr, const int32_t w[6]) {
  NEED_INIT();
  if (!vs || !vp || !rho || !grid_ok(gr) || !w) return fail(MCT_E_INVALID_ARG, "vs2vp_rho_window: bad arguments");
  const int wx = w[1] - w[0] + 1, wy = w[3] - w[2] + 1, wz = w[5] - w[4] + 1;
  if (wx <= 0 || wy <= 0 || wz <= 0) return MCT_OK;
  if (w[0] < 1 || w[2] < 1 || w[4] < 1 || w[1] > gr->nx || w[3] > gr->ny || w[5] > gr->nz)
    return fail(MCT_E_INVALID_ARG, "vs2vp_rho_window: window outside the grid");
  // packed (wz,wy,wx) staging, like mct_voronoi_to_grid
  const size_t nn = (size_t)wx * wy * wz;
  int rc;
  if ((rc = ensure(g.m_vs, nn * 8))) return rc;
  if ((rc = ensure(g.m_vp, nn * 8))) return rc;
  if ((rc = ensure(g.m_rho, nn * 8))) return rc;
  if ((rc = ensure_pin(g.pin_a, nn * 24))) return rc;
  cudaStream_t st = g.stream;
  double* h_vs = (double*)g.pin_a.p;
  double* h_vp = h_vs + nn;
  double* h_rho = h_vp + nn;
  const size_t ny = gr->ny, nz = gr->nz;
  for (int i = 0; i < wx; ++i)
    for (int j = 0; j < wy; ++j)
      memcpy(h_vs + ((size_t)i * wy + j) * wz, vs + ((size_t)(w[0] - 1 + i) * ny + (size_t)(w[2] - 1 + j)) * nz + (size_t)(w[4] - 1),
             (size_t)wz * 8);
  CK(cudaMemcpyAsync(g.m_vs.p, h_vs, nn * 8, cudaMemcpyHostToDevice, st));
  if ((rc = mct_vs2vp_rho_dev((double*)g.m_vs.p, (double*)g.m_vp.p, (double*)g.m_rho.p, (int64_t)nn, st))) return rc;
  CK(cudaMemcpyAsync(h_vp, g.m_vp.p, nn * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_rho, g.m_rho.p, nn * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int i = 0; i < wx; ++i)
    for (int j = 0; j < wy; ++j) {
      const size_t uo = ((size_t)(w[0] - 1 + i) * ny + (size_t)(w[2] - 1 + j)) * nz + (size_t)(w[4] - 1);
      const size_t po = ((size_t)i * wy + j) * wz;
      memcpy(vp + uo, h_vp + po, (size_t)wz * 8);
      memcpy(rho + uo, h_rho + po, (size_t)wz * 8);
    }
  return MCT_OK;
}

int mct_surf_dispersion_dev(const double* d_vp, const double* d_vs, const double* d_rho, const mct_grid* gr, int ix0, int ix1,
                            int iy0, int iy1, const double* freqs, int np, const mct_disp_opts* opt, double* d_pvel,
                            double* d_gvel, int32_t* d_ierr, int32_t* d_flags, void* stream) {
  NEED_INIT();
  if (!d_vp || !d_vs || !d_rho || !freqs || !d_pvel || !d_gvel || !d_ierr) return fail(MCT_E_INVALID_ARG, "surf_dispersion: NULL pointer");
  DispPlan pl;
  int rc = plan_disp(gr, ix0, ix1, iy0, iy1, np, opt, pl);
  if (rc) return rc;
  const bool chk = d_flags != nullptr;
  int32_t* fl = d_flags ? d_flags : (int32_t*)g.flags.p;
  const int whole[4] = {1, 1, gr->nx, gr->ny}, win[4] = {pl.ix0, pl.iy0, pl.wx, pl.wy};
  return disp_core(d_vp, d_vs, d_rho, gr, pl, freqs, np, opt, chk, opt->check_scope == 1 ? win : whole, d_pvel, d_gvel, d_ierr, fl,
                   pick(stream));
}

int mct_surf_dispersion(const double* vp, const double* vs, const double* rho, const mct_grid* gr, int ix0, int ix1, int iy0,
                        int iy1, const double* freqs, int np, const mct_disp_opts* opt, double* pvel, double* gvel,
                        int32_t* ierr, int32_t* model_invalid) {
  NEED_INIT();
  if (!vp || !vs || !rho || !freqs || !pvel || !gvel || !ierr) return fail(MCT_E_INVALID_ARG, "surf_dispersion: NULL pointer");
  DispPlan pl;
  int rc = plan_disp(gr, ix0, ix1, iy0, iy1, np, opt, pl);
  if (rc) return rc;
  cudaStream_t st = g.stream;
  const size_t ncell = (size_t)gr->nx * gr->ny * gr->nz;
  if ((rc = ensure(g.m_vp, ncell * 8))) return rc;
  if ((rc = ensure(g.m_vs, ncell * 8))) return rc;
  if ((rc = ensure(g.m_rho, ncell * 8))) return rc;
  const size_t nbo = (size_t)pl.ncol * pl.nout * 8;
  if ((rc = ensure(g.o_pvel, nbo))) return rc;
  if ((rc = ensure(g.o_gvel, nbo))) return rc;
  if ((rc = ensure(g.o_ierr, (size_t)pl.ncol * 4))) return rc;
  // Only the columns the kernels will read are sent: the window's columns (for each x index of the window the
  // wy columns are one contiguous run of wy*nz values: a 2-D copy with the (ny*nz) slab as pitch) -- plus, when
  // check_model runs with the reference's whole-grid scope, all of vs.
  const size_t slab = (size_t)gr->ny * gr->nz;
  const size_t woff = (size_t)(pl.ix0 - 1) * slab + (size_t)(pl.iy0 - 1) * gr->nz;
  const size_t wrow = (size_t)pl.wy * gr->nz * 8;
  const bool whole_check = model_invalid != nullptr && opt->check_scope != 1;
  if (whole_check) CK(cudaMemcpyAsync(g.m_vs.p, vs, ncell * 8, cudaMemcpyHostToDevice, st));
  else CK(cudaMemcpy2DAsync((double*)g.m_vs.p + woff, slab * 8, vs + woff, slab * 8, wrow, (size_t)pl.wx, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpy2DAsync((double*)g.m_vp.p + woff, slab * 8, vp + woff, slab * 8, wrow, (size_t)pl.wx, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpy2DAsync((double*)g.m_rho.p + woff, slab * 8, rho + woff, slab * 8, wrow, (size_t)pl.wx, cudaMemcpyHostToDevice, st));
  int32_t* fl = (int32_t*)g.flags.p;
  const int whole[4] = {1, 1, gr->nx, gr->ny}, win[4] = {pl.ix0, pl.iy0, pl.wx, pl.wy};
  rc = disp_core((double*)g.m_vp.p, (double*)g.m_vs.p, (double*)g.m_rho.p, gr, pl, freqs, np, opt, model_invalid != nullptr,
                 whole_check ? whole : win, (double*)g.o_pvel.p, (double*)g.o_gvel.p, (int32_t*)g.o_ierr.p, fl, st);
  if (rc) return rc;
  int32_t hflags[2] = {0, 0};
  CK(cudaMemcpyAsync(hflags, fl, sizeof hflags, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (model_invalid) *model_invalid = hflags[0];
  if (model_invalid && hflags[0]) return MCT_OK; // reference returns before solving (likelihood_surf.F90:161-164)
  CK(cudaMemcpyAsync(pvel, g.o_pvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(gvel, g.o_gvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ierr, g.o_ierr.p, (size_t)pl.ncol * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hflags[1] >= 2) {
    const int code = flags_to_code(hflags[1]);
    return fail(code, "dispersion: at least one column reported condition %d (see ierr)", hflags[1]);
  }
  return MCT_OK;
}

int mct_surfmodes_batch(const double* thick, const double* vp, const double* vs, const double* rho, const int64_t* offsets,
                        int ncol, const double* freqs, int np, const mct_disp_opts* opt, double* phase, double* group,
                        int32_t* ierr) {
  NEED_INIT();
  if (!thick || !vp || !vs || !rho || !offsets || !freqs || !opt || !phase || !group || !ierr || ncol < 1)
    return fail(MCT_E_INVALID_ARG, "surfmodes_batch: bad arguments");
  if (np < 1 || np > MCT_MAX_PERIODS) return fail(MCT_E_INVALID_ARG, "surfmodes_batch: np must be in 1..%d", MCT_MAX_PERIODS);
  if (opt->raylov != 0 && opt->raylov != 1) return fail(MCT_E_INVALID_ARG, "surfmodes_batch: raylov must be 0 or 1");
  if (opt->nmodes > 1000) return fail(MCT_E_INVALID_ARG, "surfmodes_batch: nmodes must not exceed 1000");
  const int64_t ntot = offsets[ncol] - offsets[0];
  int maxl = 1;
  for (int c = 0; c < ncol; ++c) {
    const int64_t n = offsets[c + 1] - offsets[c];
    if (n < 1) return fail(MCT_E_INVALID_ARG, "surfmodes_batch: column %d has no layers", c);
    if (n > maxl) maxl = (int)(n > MCT_MAX_LAYERS ? MCT_MAX_LAYERS : n);
  }
  cudaStream_t st = g.stream;
  const int stride = (ncol + 31) & ~31;
  const int nm = opt->nmodes <= 0 ? 1 : opt->nmodes;
  const size_t nbo = (size_t)ncol * np * nm * 8;
  int rc;
  if ((rc = ensure(g.pl_thick, (size_t)ntot * 8))) return rc;
  if ((rc = ensure(g.pl_vp, (size_t)ntot * 8))) return rc;
  if ((rc = ensure(g.pl_vs, (size_t)ntot * 8))) return rc;
  if ((rc = ensure(g.pl_rho, (size_t)ntot * 8))) return rc;
  if ((rc = ensure(g.pl_off, (size_t)(ncol + 1) * 8))) return rc;
  if ((rc = ensure(g.lay, sizeof(float4) * (size_t)stride * (size_t)maxl))) return rc;
  if ((rc = ensure(g.nlay, 4 * (size_t)stride))) return rc;
  if ((rc = ensure(g.status, 4 * (size_t)stride))) return rc;
  if ((rc = ensure(g.o_pvel, nbo))) return rc;
  if ((rc = ensure(g.o_gvel, nbo))) return rc;
  if ((rc = ensure(g.o_ierr, (size_t)ncol * 4))) return rc;
  const int64_t o0 = offsets[0];
  CK(cudaMemcpyAsync(g.pl_thick.p, thick + o0, (size_t)ntot * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.pl_vp.p, vp + o0, (size_t)ntot * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.pl_vs.p, vs + o0, (size_t)ntot * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.pl_rho.p, rho + o0, (size_t)ntot * 8, cudaMemcpyHostToDevice, st));
  std::vector<long long> rel((size_t)ncol + 1);
  for (int c = 0; c <= ncol; ++c) rel[c] = (long long)(offsets[c] - o0);
  CK(cudaMemcpyAsync(g.pl_off.p, rel.data(), (size_t)(ncol + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st)); // rel is a stack-lifetime pageable buffer
  int32_t* fl = (int32_t*)g.flags.p;
  CK(cudaMemsetAsync(fl, 0, 2 * sizeof(int32_t), st));
  PreLayParams L;
  L.thick = (const double*)g.pl_thick.p; L.vp = (const double*)g.pl_vp.p; L.vs = (const double*)g.pl_vs.p; L.rho = (const double*)g.pl_rho.p;
  L.offsets = (const long long*)g.pl_off.p;
  L.ncol = ncol; L.modetype = opt->raylov; L.stride = stride;
  L.lay = (float4*)g.lay.p; L.nlay = (int32_t*)g.nlay.p; L.status = (int32_t*)g.status.p; L.flags = fl;
  L.grt_on = (g.grt_on && opt->nmodes <= 0) ? 1 : 0;
  prelayered_kernel<<<(ncol + 127) / 128, 128, 0, st>>>(L);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  if ((rc = launch_k2(ncol, stride, freqs, np, opt, (double*)g.o_pvel.p, (double*)g.o_gvel.p, (int32_t*)g.o_ierr.p, nullptr, ncol, maxl, st))) return rc;
  if ((rc = launch_grt(ncol, nullptr, &L, freqs, np, opt, (double*)g.o_pvel.p, (double*)g.o_gvel.p, (int32_t*)g.o_ierr.p, nullptr, fl, ncol, st))) return rc;
  int32_t hflags[2] = {0, 0};
  CK(cudaMemcpyAsync(phase, g.o_pvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(group, g.o_gvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ierr, g.o_ierr.p, (size_t)ncol * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hflags, fl, sizeof hflags, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hflags[1] >= 2) {
    const int code = flags_to_code(hflags[1]);
    return fail(code, "surfmodes_batch: at least one column reported condition %d (see ierr)", hflags[1]);
  }
  return MCT_OK;
}

int mct_set_nuclei_batch(const double* points, const double* params, const int64_t* offsets, int nb) {
  NEED_INIT();
  if (!offsets || nb < 1) return fail(MCT_E_INVALID_ARG, "set_nuclei_batch: bad arguments");
  std::vector<long long> off((size_t)nb + 1);
  for (int b = 0; b <= nb; ++b) off[b] = (long long)offsets[b];
  int rc = upload_nuclei(points, params, off.data(), nb, g.stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(g.stream));
  return MCT_OK;
}

int mct_forward_batch_dev(const mct_grid* gr, int nb, int derive_vp_rho, int ixs0, int ixs1, const double* freqs, int np,
                          const mct_disp_opts* opt, double* d_vp, double* d_vs, double* d_rho, int32_t* d_sites_id,
                          double* d_pvel, double* d_gvel, int32_t* d_ierr, int32_t* d_flags, void* stream) {
  NEED_INIT();
  if (!d_vp || !d_vs || !d_rho || !d_sites_id || !d_pvel || !d_gvel || !d_ierr || !d_flags || !freqs)
    return fail(MCT_E_INVALID_ARG, "forward_batch: NULL pointer");
  DispPlan pl;
  int rc = plan_disp(gr, ixs0, ixs1, 1, gr ? gr->ny : 0, np, opt, pl, nb);
  if (rc) return rc;
  return forward_core(gr, nb, derive_vp_rho, pl, freqs, np, opt, d_vp, d_vs, d_rho, d_sites_id, d_pvel, d_gvel, d_ierr,
                      d_flags, pick(stream));
}

int mct_forward_eval_dev(const double* points, const double* params, int ncells, const mct_grid* gr, int derive_vp_rho,
                         int ixs0, int ixs1, const double* freqs, int np, const mct_disp_opts* opt, double* d_vp,
                         double* d_vs, double* d_rho, int32_t* d_sites_id, double* d_pvel, double* d_gvel, int32_t* d_ierr,
                         int32_t* d_flags, void* stream) {
  NEED_INIT();
  int rc = upload_nuclei(points, params, ncells, pick(stream));
  if (rc) return rc;
  return mct_forward_batch_dev(gr, 1, derive_vp_rho, ixs0, ixs1, freqs, np, opt, d_vp, d_vs, d_rho, d_sites_id, d_pvel,
                               d_gvel, d_ierr, d_flags, stream);
}

int mct_forward_eval_batch(const double* points, const double* params, const int64_t* offsets, int nb, const mct_grid* gr,
                           int derive_vp_rho, const double* freqs, int np, const mct_disp_opts* opt, double* pvel,
                           double* gvel, int32_t* ierr, int32_t* model_invalid, double* vp, double* vs, double* rho,
                           int32_t* sites_id) {
  NEED_INIT();
  if (!pvel || !gvel || !ierr || !freqs || !offsets) return fail(MCT_E_INVALID_ARG, "forward_eval: NULL pointer");
  DispPlan pl;
  int rc = plan_disp(gr, 1, gr ? gr->nx : 0, 1, gr ? gr->ny : 0, np, opt, pl, nb);
  if (rc) return rc;
  cudaStream_t st = g.stream;
  std::vector<long long> off((size_t)nb + 1);
  for (int b = 0; b <= nb; ++b) off[b] = (long long)offsets[b];
  if ((rc = upload_nuclei(points, params, off.data(), nb, st))) return rc;
  const size_t ncell = (size_t)gr->nx * gr->ny * gr->nz * (size_t)nb;
  if ((rc = ensure(g.m_vp, ncell * 8))) return rc;
  if ((rc = ensure(g.m_vs, ncell * 8))) return rc;
  if ((rc = ensure(g.m_rho, ncell * 8))) return rc;
  if ((rc = ensure(g.m_sites, ncell * 4))) return rc;
  const size_t nbo = (size_t)pl.ncol * pl.nout * 8;
  if ((rc = ensure(g.o_pvel, nbo))) return rc;
  if ((rc = ensure(g.o_gvel, nbo))) return rc;
  if ((rc = ensure(g.o_ierr, (size_t)pl.ncol * 4))) return rc;
  if ((rc = ensure(g.bflags, (size_t)nb * 2 * sizeof(int32_t)))) return rc;
  int32_t* fl = (int32_t*)g.bflags.p;
  rc = forward_core(gr, nb, derive_vp_rho, pl, freqs, np, opt, (double*)g.m_vp.p, (double*)g.m_vs.p, (double*)g.m_rho.p,
                    (int32_t*)g.m_sites.p, (double*)g.o_pvel.p, (double*)g.o_gvel.p, (int32_t*)g.o_ierr.p, fl, st);
  if (rc) return rc;
  std::vector<int32_t> hflags((size_t)nb * 2);
  int32_t k1err = 0;
  // Outputs of rejected models are left untouched on the device side (K2 skips them); copy everything in one go --
  // the caller decides per model_invalid[b] what to read, like the Fortran which returns before filling pvel.
  CK(cudaMemcpyAsync(pvel, g.o_pvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(gvel, g.o_gvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ierr, g.o_ierr.p, (size_t)pl.ncol * 4, cudaMemcpyDeviceToHost, st));
  if (vp) CK(cudaMemcpyAsync(vp, g.m_vp.p, ncell * 8, cudaMemcpyDeviceToHost, st));
  if (vs) CK(cudaMemcpyAsync(vs, g.m_vs.p, ncell * 8, cudaMemcpyDeviceToHost, st));
  if (rho) CK(cudaMemcpyAsync(rho, g.m_rho.p, ncell * 8, cudaMemcpyDeviceToHost, st));
  if (sites_id) CK(cudaMemcpyAsync(sites_id, g.m_sites.p, ncell * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hflags.data(), fl, hflags.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&k1err, (int32_t*)g.flags.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (k1err) return fail(MCT_E_CUDA, "nearest-nucleus traversal stack overflow (tree deeper than %d)", K1_STACK);
  int maxst = 0;
  for (int b = 0; b < nb; ++b) {
    if (model_invalid) model_invalid[b] = hflags[2 * b];
    if (!hflags[2 * b] && hflags[2 * b + 1] > maxst) maxst = hflags[2 * b + 1];
  }
  if (maxst >= 2) return fail(flags_to_code(maxst), "forward_eval: at least one column reported condition %d (see ierr)", maxst);
  return MCT_OK;
}

int mct_forward_eval(const double* points, const double* params, int ncells, const mct_grid* gr, int derive_vp_rho,
                     const double* freqs, int np, const mct_disp_opts* opt, double* pvel, double* gvel, int32_t* ierr,
                     int32_t* model_invalid, double* vp, double* vs, double* rho, int32_t* sites_id) {
  const int64_t off[2] = {0, ncells};
  if (ncells < 1) return fail(MCT_E_INVALID_ARG, "forward_eval: ncells < 1");
  return mct_forward_eval_batch(points, params, off, 1, gr, derive_vp_rho, freqs, np, opt, pvel, gvel, ierr, model_invalid, vp,
                                vs, rho, sites_id);
}

int mct_set_profiling(int on) {
  g.prof_on = on ? 1 : 0;
  return MCT_OK;
}

int mct_kernel_times(double ms[4], int reset) {
  NEED_INIT();
  if (!ms) return fail(MCT_E_INVALID_ARG, "kernel_times: NULL pointer");
  CK(cudaDeviceSynchronize());
  for (auto& e : g.ev_pending) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e.a, e.b));
    g.prof_ms[e.kind < 2 ? e.kind : 2] += (double)t;
    g.prof_ms[3] += 1.0;
    g.ev_pool.push_back(e);
  }
  g.ev_pending.clear();
  for (int i = 0; i < 4; ++i) ms[i] = g.prof_ms[i];
  if (reset) for (int i = 0; i < 4; ++i) g.prof_ms[i] = 0.0;
  return MCT_OK;
}

int mct_fp64_peak_probe(double* tflops_fma, double* tflops_mul_add) {
  NEED_INIT();
  if (!tflops_fma || !tflops_mul_add) return fail(MCT_E_INVALID_ARG, "fp64_peak_probe: NULL pointer");
  const int blocks = g.sm_count * 8, threads = 256, iters = 20000;
  DevBuf out;
  int rc = ensure(out, sizeof(double) * (size_t)blocks * threads);
  if (rc) return rc;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double res[2] = {0, 0};
  for (int fused = 1; fused >= 0; --fused) {
    fp64_probe_kernel<<<blocks, threads, 0, g.stream>>>((double*)out.p, 2000, fused); // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(a, g.stream));
      fp64_probe_kernel<<<blocks, threads, 0, g.stream>>>((double*)out.p, iters, fused);
      CK(cudaEventRecord(b, g.stream));
      CK(cudaEventSynchronize(b));
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, a, b));
      if (t < best) best = t;
    }
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    res[fused ? 0 : 1] = flops / (best * 1e-3) / 1e12;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  release(out);
  *tflops_fma = res[0];
  *tflops_mul_add = res[1];
  return MCT_OK;
}

int mct_accumulate_stats_dev(const double* d_vs, const double* d_vp, double* d_aveS, double* d_stdS, double* d_aveP,
                             double* d_stdP, int64_t n, void* stream) {
  NEED_INIT();
  if (!d_vs || !d_vp || !d_aveS || !d_stdS || !d_aveP || !d_stdP || n < 0) return fail(MCT_E_INVALID_ARG, "accumulate_stats: bad arguments");
  if (n == 0) return MCT_OK;
  accumulate_stats_kernel<<<grid_blocks(n, 256, 8), 256, 0, pick(stream)>>>(d_vs, d_vp, d_aveS, d_stdS, d_aveP, d_stdP, (long long)n);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  return MCT_OK;
}

int mct_set_grt(int enable, const double* par6) {
  if (par6) {
    for (int i = 0; i < 6; ++i) if (!(par6[i] > 0)) return fail(MCT_E_INVALID_ARG, "set_grt: every parameter must be positive");
    for (int i = 0; i < 6; ++i) g.grt_par[i] = par6[i];
  }
  g.grt_on = enable ? 1 : 0;
  return MCT_OK;
}

int mct_grt_stats(int64_t out3[3]) {
  NEED_INIT();
  if (!out3) return fail(MCT_E_INVALID_ARG, "grt_stats: NULL pointer");
  unsigned long long c[2];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(c, (unsigned long long*)g.counters.p + 8, sizeof c, cudaMemcpyDeviceToHost));
  out3[0] = g.grt_last_cols; out3[1] = (int64_t)c[0]; out3[2] = (int64_t)c[1];
  return MCT_OK;
}

int mct_set_k1_mode(int mode) {
  if (mode < 0 || mode > 3) return fail(MCT_E_INVALID_ARG, "set_k1_mode: mode must be 0 (box lists, float32 walk), 1 (tree walk), 2 (column scan) or 3 (box lists, float64 walk)");
  g.k1_mode = mode;
  return MCT_OK;
}

int mct_set_k2_mode(int mode, int coop_max_columns) {
  if (mode < 0 || mode > 2) return fail(MCT_E_INVALID_ARG, "set_k2_mode: mode must be 0 (auto), 1 (thread per column) or 2 (warp per column)");
  g.k2_mode = mode;
  if (coop_max_columns >= 0) g.k2_coop_max = coop_max_columns;
  return MCT_OK;
}

int mct_set_dedup(int on) {
  g.dedup = on ? 1 : 0;
  return MCT_OK;
}

int mct_last_launch(mct_launch_info* out) {
  if (!out) return fail(MCT_E_INVALID_ARG, "last_launch: NULL pointer");
  memset(out, 0, sizeof *out);
  snprintf(out->kernel, sizeof out->kernel, "%s", g.last_kernel);
  out->columns = g.last_ncol;
  out->columns_solved = g.last_neff;
  out->lanes_per_column = g.last_lanes;
  out->sm_count = g.sm_count;
  return MCT_OK;
}

int mct_set_k2_lanes(int lanes_per_column) {
  const int l = lanes_per_column;
  if (l != 0 && (l < 2 || l > 256 || (l & (l - 1)) != 0))
    return fail(MCT_E_INVALID_ARG, "set_k2_lanes: lanes per column must be 0 (auto) or a power of two from 2 to 256");
  g.k2_coop_lanes = lanes_per_column;
  return MCT_OK;
}

int mct_selftest_division(int emax, int64_t* tested, int64_t* mismatches) {
  NEED_INIT();
  if (!tested || !mismatches || emax < 0 || emax > 1000) return fail(MCT_E_INVALID_ARG, "selftest_division: bad arguments");
  unsigned long long* d = (unsigned long long*)g.counters.p + 3;
  CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), g.stream));
  const int blocks = g.sm_count * 8, threads = 256, iters = 2048;
  div_selftest_kernel<<<blocks, threads, 0, g.stream>>>(0x1234567ull + (unsigned long long)emax, iters, emax, d);
  CK(cudaGetLastError());
  unsigned long long h = 0;
  CK(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  *tested = 3ll * blocks * threads * iters;
  *mismatches = (int64_t)h;
  return MCT_OK;
}

int mct_assemble_vel_dev(const double* d_pvel, int np, int nx, int ny, int ix0, int ix1, int iy0, int iy1, double* d_vel,
                         void* stream) {
  NEED_INIT();
  if (!d_pvel || !d_vel || np < 1 || ix0 < 1 || iy0 < 1 || ix1 > nx || iy1 > ny || ix1 < ix0 || iy1 < iy0)
    return fail(MCT_E_INVALID_ARG, "assemble_vel: bad arguments");
  cudaStream_t st = pick(stream);
  const int wx = ix1 - ix0 + 1, wy = iy1 - iy0 + 1;
  assemble_scatter_kernel<<<grid_blocks((long long)wx * wy * np, 256, 8), 256, 0, st>>>(d_pvel, np, ny, ix0, iy0, wx, wy, d_vel);
  CK(cudaGetLastError());
  const int ex0 = ix0 == 1, ex1 = ix1 == nx, ey0 = iy0 == 1, ey1 = iy1 == ny;
  if (ex0 || ex1) {
    assemble_edges_kernel<<<grid_blocks((long long)np * (ny + 2), 256, 8), 256, 0, st>>>(d_vel, np, nx, ny, ex0, ex1, ey0, ey1, 0);
    CK(cudaGetLastError());
    g.host_stats.n_launches += 1;
  }
  if (ey0 || ey1) {
    assemble_edges_kernel<<<grid_blocks((long long)np * (nx + 2), 256, 8), 256, 0, st>>>(d_vel, np, nx, ny, ex0, ex1, ey0, ey1, 1);
    CK(cudaGetLastError());
    g.host_stats.n_launches += 1;
  }
  g.host_stats.n_launches += 1;
  return MCT_OK;
}

// ---- the 2-D product and point location -----------------------------------------------------------------
// 2-D nuclei are embedded with z = 0: kdtree2's build never cuts a zero-extent dimension and a zero third term changes
// no sum, so tree, traversal and tie order are those of kdtree2 run with dim = 2 (pinned on kdtree2.o: fixture grid2d).
static void embed_2d(const double* p2, int n, std::vector<double>& p3) {
  p3.resize(3 * (size_t)n);
  for (int i = 0; i < n; ++i) { p3[3 * (size_t)i] = p2[2 * (size_t)i]; p3[3 * (size_t)i + 1] = p2[2 * (size_t)i + 1]; p3[3 * (size_t)i + 2] = 0.0; }
}

int mct_voronoi_to_grid_2d(const double* points2, const double* params, int ncells, int nx, int ny, double xmin, double ymin,
                           double dx, double dy, double* vp, double* vs, double* rho, int32_t* sites_id) {
  NEED_INIT();
  if (!points2 || !params || ncells < 1 || nx < 1 || ny < 1 || !(dx > 0) || !(dy > 0)) return fail(MCT_E_INVALID_ARG, "voronoi_to_grid_2d: bad arguments");
  std::vector<double> p3;
  embed_2d(points2, ncells, p3);
  mct_grid g3 = {nx, ny, 1, xmin, ymin, 0.0, dx, dy, 1.0, 0.0, 1.0};
  const double box[6] = {xmin - dx, ymin - dy, -1.0, xmin + nx * dx, ymin + ny * dy, 1.0}; // every node (mcmc2d/mcmc.f90:1497-1500)
  return mct_voronoi_to_grid(p3.data(), params, ncells, &g3, box, nullptr, vp, vs, rho, sites_id);
}

static int points_query(const double* points, int dim, int ncells, const double* queries, long long nq, const int32_t* d_sites,
                        const mct_grid* gr, int32_t* out) {
  if (!points || (dim != 2 && dim != 3) || ncells < 1 || !queries || nq < 0 || !out) return fail(MCT_E_INVALID_ARG, "nearest_nucleus: bad arguments");
  if (nq == 0) return MCT_OK;
  cudaStream_t st = g.stream;
  std::vector<double> p3, q3;
  const double* P3 = points;
  const double* Q3 = queries;
  if (dim == 2) {
    embed_2d(points, ncells, p3);
    q3.resize(3 * (size_t)nq);
    for (long long i = 0; i < nq; ++i) { q3[3 * i] = queries[2 * i]; q3[3 * i + 1] = queries[2 * i + 1]; q3[3 * i + 2] = 0.0; }
    P3 = p3.data(); Q3 = q3.data();
  }
  std::vector<double> dummy_par(3 * (size_t)ncells, 0.0);
  int rc = upload_nuclei(P3, dummy_par.data(), ncells, st);
  if (rc) return rc;
  if ((rc = ensure(g.ray_pts, (size_t)nq * 24)) || (rc = ensure(g.ray_off, (size_t)nq * 4))) return rc;
  CK(cudaMemcpyAsync(g.ray_pts.p, Q3, (size_t)nq * 24, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st)); // q3 / p3 are function-local
  K1Params P;
  memset(&P, 0, sizeof P);
  P.nodes = (const KdNodeDev*)g.nodes.p; P.rpts = (const double*)g.rpts.p; P.ind = (const int32_t*)g.ind.p; P.params = (const double*)g.params.p;
  P.root = g.roots[0]; P.n = ncells;
  if (gr) { P.xmin = gr->xmin; P.ymin = gr->ymin; P.zmin = gr->zmin; P.dx = gr->dx; P.dy = gr->dy; P.dz = gr->dz; }
  P.err = (int32_t*)g.flags.p + 2;
  CK(cudaMemsetAsync(P.err, 0, sizeof(int32_t), st));
  {
    ProfScope ps(0, st);
    k1_points_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(P, (const double*)g.ray_pts.p, nq, d_sites, gr ? gr->nx : 0, gr ? gr->ny : 0,
                                                                  gr ? gr->nz : 0, gr ? gr->scaling : 1.0, (int32_t*)g.ray_off.p);
  }
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  int32_t k1err = 0;
  CK(cudaMemcpyAsync(out, g.ray_off.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&k1err, P.err, sizeof k1err, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (k1err) return fail(MCT_E_CUDA, "nearest-nucleus traversal stack overflow (tree deeper than %d)", K1_STACK);
  return MCT_OK;
}

int mct_nearest_nucleus(const double* points, int dim, int ncells, const double* queries, int64_t nq, int32_t* idx) {
  NEED_INIT();
  return points_query(points, dim, ncells, queries, (long long)nq, nullptr, nullptr, idx);
}

int mct_sites_locate_dev(const double* points, int ncells, const int32_t* d_sites_id, const mct_grid* gr, const double* queries,
                         int64_t nq, int32_t* idx) {
  NEED_INIT();
  if (!d_sites_id || !grid_ok(gr) || gr->nx < 2 || gr->ny < 2 || gr->nz < 2 || !(gr->scaling != 0.0))
    return fail(MCT_E_INVALID_ARG, "sites_locate: bad arguments (the eight-node stencil needs nx, ny, nz >= 2)");
  return points_query(points, 3, ncells, queries, (long long)nq, d_sites_id, gr, idx);
}

int mct_sites_locate(const double* points, int ncells, const int32_t* sites_id, const mct_grid* gr, const double* queries, int64_t nq,
                     int32_t* idx) {
  NEED_INIT();
  if (!sites_id || !grid_ok(gr)) return fail(MCT_E_INVALID_ARG, "sites_locate: bad arguments");
  const size_t nn = (size_t)gr->nx * gr->ny * gr->nz;
  int rc = ensure(g.m_sites, nn * 4);
  if (rc) return rc;
  CK(cudaMemcpyAsync(g.m_sites.p, sites_id, nn * 4, cudaMemcpyHostToDevice, g.stream));
  return mct_sites_locate_dev(points, ncells, (const int32_t*)g.m_sites.p, gr, queries, nq, idx);
}

} // extern "C"

#include "mct_session.cuh" // mct_session_*: a chain's model resident in HBM between proposals
#include "k3_raytime.cuh"  // mct_group_times_dev: CalGroupTime on the device map
#include "mct_comm.cuh"    // mct_comm_*, mct_allgather_inplace, mct_forward_sharded_dev: NCCL data plane (config 5)
#include "k4_misfit.cuh"   // misfit sums, session likelihood / ray times / stat_rti accumulation
#include "k6_fm2d.cuh"     // mct_fm2d_times[_dev]: the 2-D fast-marching travel times of modrays
