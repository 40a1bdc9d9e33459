// mct_session.cuh -- a chain's model kept resident in HBM between proposals (included by mct_api.cu).
//
// The reference's sampler changes one Voronoi cell per iteration and then calls, in this order
// (src/mcmc_loc2.f90:199-228, 556-566; src/likelihood.f90:75-83; src/likelihood_surf.F90:155-231):
//   backup of the model  ->  kdtree_to_grid(RTI, grid, bnd_box, model[, pm])  ->  vs2vp_3d / vp2rho_3d over the grid
//   ->  check_model over the grid  ->  dispersion over the columns of the box + a one-column halo
//   ->  accept (keep) or reject (restore the backup).
// Through the host-pointer entry points each of those steps ships model windows over PCIe in both directions
// (~6 MB each way for a typical C1 proposal: more time than the kernels take).  A session keeps vp/vs/rho/sites_id
// and the dispersion maps of the CURRENT model on the device; a proposal sends the nuclei (48 B each) and the box,
// and brings back only the window's dispersion maps.  check_model keeps the reference's whole-grid scope at no
// cost, because the whole grid is already there.
#pragma once

// device copies of the likelihood's data (T_DATA%ttime, %raystat, like%srdist) and its work arrays: k4_misfit.cuh
struct MisfitBufs {
  DevBuf ttime, raystat, srdist, snoise, sigma, terms, out, time;
  int nrr = 0, np = 0, sigdep = 0, nrays_total = 0;
  bool have = false;
};

struct mct_session {
  mct_grid gr;
  mct_disp_opts opt;
  int np = 0, nout = 0, derive = 1;
  std::vector<double> freqs;
  DevBuf vp, vs, rho, sites;          // current model, (nz,ny,nx)
  DevBuf pvel, gvel, ierr;            // current maps, (nout,ny,nx) / (ny,nx)
  DevBuf b_vp, b_vs, b_rho, b_sites;  // packed backup of the last proposal's box
  DevBuf w_pvel, w_gvel, w_ierr;      // packed maps of the last proposal's window
  DevBuf flags;                       // int32[2]
  DevBuf r_pts, r_off;                // resident rays (straight-ray mode: set once), packed like mct_group_times_dev's
  int r_nrays = 0;
  DevBuf time;                        // ray times of the last group_times / likelihood call, (nrays, np)
  int time_nrays = 0;
  MisfitBufs mf;                      // observed data + misfit work arrays
  DevBuf acc;                         // stat_rti accumulators: aveS, stdS, aveP, stdP, (nz,ny,nx) each
  // curved-ray mode (settings%isStraight == 0, phase-velocity data): sources / receivers, the padded map like%vel, per-problem status
  DevBuf f_geo, f_vel, f_err, f_rays;
  int f_nsrc = 0, f_nrc = 0;
  int32_t f_opt_i[6] = {1, 1, 1, 4, 8, 1}; // gridx, gridy, sgref, sgdic, sgext, order
  double f_band = 0.5;
  bool f_have = false;
  long long nacc = 0;
  bool have_model = false, pending = false;
  bool maps_valid = false;            // the resident maps belong to a model check_model accepted
  int32_t pbox[6] = {0, 0, 0, 0, 0, 0}; // node window of the pending proposal
  int32_t pwin[4] = {0, 0, 0, 0};       // its column window (with halo)
  int pinvalid = 0, pcode = 0;
};

namespace {

// packed (wz,wy,wx) block <-> window of the (nz,ny,nx) arrays; dir 0: gather (array -> block), 1: scatter
__global__ void __launch_bounds__(256) box_copy_kernel(double* vp, double* vs, double* rho, int32_t* sites, double* b_vp, double* b_vs,
                                                       double* b_rho, int32_t* b_sites, int ix0, int iy0, int iz0, int wx, int wy,
                                                       int wz, int ny, int nz, int dir) {
  const long long n = (long long)wx * wy * wz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % wz);
    const long long r = t / wz;
    const int j = (int)(r % wy), i = (int)(r / wy);
    const size_t a = ((size_t)(ix0 - 1 + i) * ny + (size_t)(iy0 - 1 + j)) * nz + (size_t)(iz0 - 1 + k);
    if (dir == 0) { b_vp[t] = vp[a]; b_vs[t] = vs[a]; b_rho[t] = rho[a]; b_sites[t] = sites[a]; }
    else { vp[a] = b_vp[t]; vs[a] = b_vs[t]; rho[a] = b_rho[t]; sites[a] = b_sites[t]; }
  }
}

// vs2vp_3d / vp2rho_3d restricted to a node window (the rest of the grid already satisfies vp = 1.73 vs, ...)
__global__ void __launch_bounds__(256) vs2vp_rho_box_kernel(const double* __restrict__ vs, double* __restrict__ vp, double* __restrict__ rho,
                                                            int ix0, int iy0, int iz0, int wx, int wy, int wz, int ny, int nz) {
  const long long n = (long long)wx * wy * wz;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % wz);
    const long long r = t / wz;
    const int j = (int)(r % wy), i = (int)(r / wy);
    const size_t a = ((size_t)(ix0 - 1 + i) * ny + (size_t)(iy0 - 1 + j)) * nz + (size_t)(iz0 - 1 + k);
    const double p = vs[a] * (double)1.730f;        // src/utils.f90:102-112 (default-real literals, as in vs2vp_rho_kernel)
    vp[a] = p;
    rho[a] = (double)1.74f * mct_pow025(p);         // src/utils.f90:125-134
  }
}

// packed window maps (nout,wy,wx) -> the resident whole-grid maps
__global__ void __launch_bounds__(256) maps_commit_kernel(const double* __restrict__ w_pvel, const double* __restrict__ w_gvel,
                                                          const int32_t* __restrict__ w_ierr, double* pvel, double* gvel, int32_t* ierr,
                                                          int ix0, int iy0, int wx, int wy, int ny, int nout) {
  const long long n = (long long)wx * wy * nout;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % nout);
    const long long c = t / nout;
    const int j = (int)(c % wy), i = (int)(c / wy);
    const size_t col = (size_t)(ix0 - 1 + i) * ny + (size_t)(iy0 - 1 + j);
    pvel[col * nout + k] = w_pvel[t];
    gvel[col * nout + k] = w_gvel[t];
    if (k == 0) ierr[col] = w_ierr[c];
  }
}

__global__ void __launch_bounds__(256) fill_kernel(double* p, long long n, double v) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = v;
}

void session_release(mct_session* s) {
  DevBuf* bufs[] = {&s->vp, &s->vs, &s->rho, &s->sites, &s->pvel, &s->gvel, &s->ierr, &s->b_vp, &s->b_vs, &s->b_rho,
                    &s->b_sites, &s->w_pvel, &s->w_gvel, &s->w_ierr, &s->flags, &s->r_pts, &s->r_off, &s->time, &s->acc,
                    &s->mf.ttime, &s->mf.raystat, &s->mf.srdist, &s->mf.snoise, &s->mf.sigma, &s->mf.terms, &s->mf.out, &s->mf.time,
                    &s->f_geo, &s->f_vel, &s->f_err, &s->f_rays};
  for (DevBuf* b : bufs) release(*b);
}

} // namespace

extern "C" {

int mct_session_create(const mct_grid* gr, const double* freqs, int np, const mct_disp_opts* opt, int derive_vp_rho,
                       mct_session** out) {
  NEED_INIT();
  if (!out || !freqs) return fail(MCT_E_INVALID_ARG, "session_create: NULL pointer");
  DispPlan pl;
  int rc = plan_disp(gr, 1, gr ? gr->nx : 0, 1, gr ? gr->ny : 0, np, opt, pl); // validates grid, np, options
  if (rc) return rc;
  mct_session* s = new (std::nothrow) mct_session();
  if (!s) return fail(MCT_E_CUDA, "session_create: out of host memory");
  s->gr = *gr;
  s->opt = *opt;
  s->np = np;
  s->nout = pl.nout;
  s->derive = derive_vp_rho ? 1 : 0;
  s->freqs.assign(freqs, freqs + np);
  const size_t nn = (size_t)gr->nx * gr->ny * gr->nz, nc = (size_t)gr->nx * gr->ny;
  if ((rc = ensure(s->vp, nn * 8)) || (rc = ensure(s->vs, nn * 8)) || (rc = ensure(s->rho, nn * 8)) || (rc = ensure(s->sites, nn * 4)) ||
      (rc = ensure(s->pvel, nc * pl.nout * 8)) || (rc = ensure(s->gvel, nc * pl.nout * 8)) || (rc = ensure(s->ierr, nc * 4)) ||
      (rc = ensure(s->flags, 2 * sizeof(int32_t)))) {
    session_release(s);
    delete s;
    return rc;
  }
  // Until a valid model is set the maps hold the reference's presets (likelihood_surf.F90:186-191; 0 in the
  // multi-mode twin), never raw memory.
  fill_kernel<<<grid_blocks((long long)(nc * pl.nout), 256, 8), 256, 0, g.stream>>>((double*)s->pvel.p, (long long)(nc * pl.nout),
                                                                                   opt->nmodes <= 0 ? opt->preset : 0.0);
  fill_kernel<<<grid_blocks((long long)(nc * pl.nout), 256, 8), 256, 0, g.stream>>>((double*)s->gvel.p, (long long)(nc * pl.nout),
                                                                                   opt->nmodes <= 0 ? opt->preset : 0.0);
  cudaMemsetAsync(s->ierr.p, 0, nc * 4, g.stream);
  if (cudaGetLastError() != cudaSuccess) { session_release(s); delete s; return fail(MCT_E_CUDA, "session_create: initialising the maps failed"); }
  g.host_stats.n_launches += 2;
  *out = s;
  return MCT_OK;
}

int mct_session_destroy(mct_session* s) {
  std::lock_guard<std::recursive_mutex> api_lock_(g_mu);
  if (!s) return MCT_OK;
  if (g.init) cudaStreamSynchronize(g.stream);
  session_release(s);
  delete s;
  return MCT_OK;
}

// kdtree_to_grid over every node + property maps + check_model + dispersion of every column; becomes the current
// model.  pvel/gvel (nout,ny,nx), ierr (ny,nx) and model_invalid may be NULL (results stay on the device).
int mct_session_set_model(mct_session* s, const double* points, const double* params, int ncells, double* pvel, double* gvel,
                          int32_t* ierr, int32_t* model_invalid) {
  NEED_INIT();
  if (!s || !points || !params) return fail(MCT_E_INVALID_ARG, "session_set_model: NULL pointer");
  cudaStream_t st = g.stream;
  int rc = upload_nuclei(points, params, ncells, st);
  if (rc) return rc;
  const mct_grid* gr = &s->gr;
  DispPlan pl;
  if ((rc = plan_disp(gr, 1, gr->nx, 1, gr->ny, s->np, &s->opt, pl))) return rc;
  s->pending = false;
  s->have_model = false;
  rc = forward_core(gr, 1, s->derive, pl, s->freqs.data(), s->np, &s->opt, (double*)s->vp.p, (double*)s->vs.p, (double*)s->rho.p,
                    (int32_t*)s->sites.p, (double*)s->pvel.p, (double*)s->gvel.p, (int32_t*)s->ierr.p, (int32_t*)s->flags.p, st);
  if (rc) return rc;
  int32_t hf[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(hf, s->flags.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hf + 2, (int32_t*)g.flags.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  const size_t nc = (size_t)gr->nx * gr->ny;
  if (pvel) CK(cudaMemcpyAsync(pvel, s->pvel.p, nc * s->nout * 8, cudaMemcpyDeviceToHost, st));
  if (gvel) CK(cudaMemcpyAsync(gvel, s->gvel.p, nc * s->nout * 8, cudaMemcpyDeviceToHost, st));
  if (ierr) CK(cudaMemcpyAsync(ierr, s->ierr.p, nc * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hf[2]) return fail(MCT_E_CUDA, "nearest-nucleus traversal stack overflow (tree deeper than %d)", K1_STACK);
  if (model_invalid) *model_invalid = hf[0];
  s->have_model = true; // (an invalid model is still the current model: its maps are simply not solved)
  s->maps_valid = !hf[0] && hf[1] < 2;
  if (!hf[0] && hf[1] >= 2) return fail(flags_to_code(hf[1]), "dispersion: at least one column reported condition %d (see ierr)", hf[1]);
  return MCT_OK;
}

// One proposal.  points/params: the nuclei AFTER the move; box: the region whose nearest nucleus may have changed
// (what the sampler passes to kdtree_to_grid); pm: NULL, or the (vp,vs,rho) of the cell of a value-only move.
// Out: win = {ix0,ix1,iy0,iy1} (1-based, box columns + one-column halo, likelihood_surf.F90:155-169) and the window's
// maps, packed (nout,wy,wx) / (wy,wx) -- the caller provides room for the largest window it may get (nx*ny columns).
// Nothing is committed: follow with mct_session_accept or mct_session_reject.
int mct_session_propose(mct_session* s, const double* points, const double* params, int ncells, const double box[6],
                        const double* pm, int32_t win[4], double* pvel_win, double* gvel_win, int32_t* ierr_win,
                        int32_t* model_invalid) {
  NEED_INIT();
  if (!s || !points || !params || !box || !win || !pvel_win || !gvel_win || !ierr_win || !model_invalid)
    return fail(MCT_E_INVALID_ARG, "session_propose: NULL pointer");
  if (!s->have_model) return fail(MCT_E_INVALID_ARG, "session_propose: no current model (call mct_session_set_model first)");
  if (s->pending) return fail(MCT_E_INVALID_ARG, "session_propose: the previous proposal was neither accepted nor rejected");
  cudaStream_t st = g.stream;
  const mct_grid* gr = &s->gr;
  int rc = upload_nuclei(points, params, ncells, st);
  if (rc) return rc;
  int32_t w[6];
  box_window(gr, box, w);
  const int wx = w[1] - w[0] + 1, wy = w[3] - w[2] + 1, wz = w[5] - w[4] + 1;
  memcpy(s->pbox, w, sizeof w);
  *model_invalid = 0;
  if (wx <= 0 || wy <= 0 || wz <= 0) { // empty box: the model does not change (the Fortran loops do nothing)
    win[0] = win[2] = 1; win[1] = win[3] = 0;
    memcpy(s->pwin, win, 4 * sizeof(int32_t));
    s->pinvalid = 0;
    s->pcode = 0;
    s->pending = true;
    return MCT_OK;
  }
  const size_t nb = (size_t)wx * wy * wz;
  if ((rc = ensure(s->b_vp, nb * 8)) || (rc = ensure(s->b_vs, nb * 8)) || (rc = ensure(s->b_rho, nb * 8)) || (rc = ensure(s->b_sites, nb * 4)))
    return rc;
  double *vp = (double*)s->vp.p, *vs = (double*)s->vs.p, *rho = (double*)s->rho.p;
  int32_t* sites = (int32_t*)s->sites.p;
  {
    ProfScope ps(2, st);
    box_copy_kernel<<<grid_blocks((long long)nb, 256, 8), 256, 0, st>>>(vp, vs, rho, sites, (double*)s->b_vp.p, (double*)s->b_vs.p,
                                                                       (double*)s->b_rho.p, (int32_t*)s->b_sites.p, w[0], w[2], w[4], wx,
                                                                       wy, wz, gr->ny, gr->nz, 0);
  }
  s->pending = true; // from here on the resident model differs from the current one until accept/reject
  CK(cudaMemsetAsync((int32_t*)g.flags.p + 2, 0, sizeof(int32_t), st));
  if ((rc = launch_k1(gr, w, pm, vp, vs, rho, sites, 1, 1, 1, gr->ny, gr->nz, st))) return rc;
  if (s->derive) {
    ProfScope ps(2, st);
    vs2vp_rho_box_kernel<<<grid_blocks((long long)nb, 256, 8), 256, 0, st>>>(vs, vp, rho, w[0], w[2], w[4], wx, wy, wz, gr->ny, gr->nz);
  }
  CK(cudaGetLastError());
  g.host_stats.n_launches += 2;
  win[0] = std::max(w[0] - 1, 1); win[1] = std::min(w[1] + 1, gr->nx);
  win[2] = std::max(w[2] - 1, 1); win[3] = std::min(w[3] + 1, gr->ny);
  memcpy(s->pwin, win, 4 * sizeof(int32_t));
  DispPlan pl;
  if ((rc = plan_disp(gr, win[0], win[1], win[2], win[3], s->np, &s->opt, pl))) return rc;
  const size_t nbo = (size_t)pl.ncol * pl.nout * 8;
  if ((rc = ensure(s->w_pvel, nbo)) || (rc = ensure(s->w_gvel, nbo)) || (rc = ensure(s->w_ierr, (size_t)pl.ncol * 4))) return rc;
  const int whole[4] = {1, 1, gr->nx, gr->ny}, wn[4] = {pl.ix0, pl.iy0, pl.wx, pl.wy};
  rc = disp_core(vp, vs, rho, gr, pl, s->freqs.data(), s->np, &s->opt, true, s->opt.check_scope == 1 ? wn : whole, (double*)s->w_pvel.p,
                 (double*)s->w_gvel.p, (int32_t*)s->w_ierr.p, (int32_t*)s->flags.p, st);
  if (rc) return rc;
  int32_t hf[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(hf, s->flags.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hf + 2, (int32_t*)g.flags.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  // the maps are small (a window): copy them optimistically with the flags, one synchronisation per proposal
  CK(cudaMemcpyAsync(pvel_win, s->w_pvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(gvel_win, s->w_gvel.p, nbo, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ierr_win, s->w_ierr.p, (size_t)pl.ncol * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hf[2]) return fail(MCT_E_CUDA, "nearest-nucleus traversal stack overflow (tree deeper than %d)", K1_STACK);
  s->pinvalid = hf[0];
  s->pcode = hf[0] ? 0 : hf[1];
  *model_invalid = hf[0];
  if (!hf[0] && hf[1] >= 2) return fail(flags_to_code(hf[1]), "dispersion: at least one column reported condition %d (see ierr)", hf[1]);
  return MCT_OK;
}

// Keep the proposed model: its window maps are merged into the resident maps (unless check_model rejected it, in
// which case no column was solved and the maps of the window are left as they were, like the reference's `like`).
int mct_session_accept(mct_session* s) {
  NEED_INIT();
  if (!s || !s->pending) return fail(MCT_E_INVALID_ARG, "session_accept: no pending proposal");
  s->pending = false;
  const int wx = s->pwin[1] - s->pwin[0] + 1, wy = s->pwin[3] - s->pwin[2] + 1;
  if (wx <= 0 || wy <= 0 || s->pinvalid) return MCT_OK;
  cudaStream_t st = g.stream;
  maps_commit_kernel<<<grid_blocks((long long)wx * wy * s->nout, 256, 8), 256, 0, st>>>(
      (const double*)s->w_pvel.p, (const double*)s->w_gvel.p, (const int32_t*)s->w_ierr.p, (double*)s->pvel.p, (double*)s->gvel.p,
      (int32_t*)s->ierr.p, s->pwin[0], s->pwin[2], wx, wy, s->gr.ny, s->nout);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  return MCT_OK;
}

// Drop the proposed model: the box is restored from its backup (src/mcmc_loc2.f90:556-566).
int mct_session_reject(mct_session* s) {
  NEED_INIT();
  if (!s || !s->pending) return fail(MCT_E_INVALID_ARG, "session_reject: no pending proposal");
  s->pending = false;
  const int32_t* w = s->pbox;
  const int wx = w[1] - w[0] + 1, wy = w[3] - w[2] + 1, wz = w[5] - w[4] + 1;
  if (wx <= 0 || wy <= 0 || wz <= 0) return MCT_OK;
  cudaStream_t st = g.stream;
  box_copy_kernel<<<grid_blocks((long long)wx * wy * wz, 256, 8), 256, 0, st>>>(
      (double*)s->vp.p, (double*)s->vs.p, (double*)s->rho.p, (int32_t*)s->sites.p, (double*)s->b_vp.p, (double*)s->b_vs.p,
      (double*)s->b_rho.p, (int32_t*)s->b_sites.p, w[0], w[2], w[4], wx, wy, wz, s->gr.ny, s->gr.nz, 1);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  return MCT_OK;
}

// Read the resident arrays back (checkpoints, tests).  Any pointer may be NULL.
int mct_session_get_model(mct_session* s, double* vp, double* vs, double* rho, int32_t* sites_id) {
  NEED_INIT();
  if (!s) return fail(MCT_E_INVALID_ARG, "session_get_model: NULL session");
  const size_t nn = (size_t)s->gr.nx * s->gr.ny * s->gr.nz;
  cudaStream_t st = g.stream;
  if (vp) CK(cudaMemcpyAsync(vp, s->vp.p, nn * 8, cudaMemcpyDeviceToHost, st));
  if (vs) CK(cudaMemcpyAsync(vs, s->vs.p, nn * 8, cudaMemcpyDeviceToHost, st));
  if (rho) CK(cudaMemcpyAsync(rho, s->rho.p, nn * 8, cudaMemcpyDeviceToHost, st));
  if (sites_id) CK(cudaMemcpyAsync(sites_id, s->sites.p, nn * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MCT_OK;
}
int mct_session_get_maps(mct_session* s, double* pvel, double* gvel, int32_t* ierr) {
  NEED_INIT();
  if (!s) return fail(MCT_E_INVALID_ARG, "session_get_maps: NULL session");
  const size_t nc = (size_t)s->gr.nx * s->gr.ny;
  cudaStream_t st = g.stream;
  if (pvel) CK(cudaMemcpyAsync(pvel, s->pvel.p, nc * s->nout * 8, cudaMemcpyDeviceToHost, st));
  if (gvel) CK(cudaMemcpyAsync(gvel, s->gvel.p, nc * s->nout * 8, cudaMemcpyDeviceToHost, st));
  if (ierr) CK(cudaMemcpyAsync(ierr, s->ierr.p, nc * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MCT_OK;
}

} // extern "C"
