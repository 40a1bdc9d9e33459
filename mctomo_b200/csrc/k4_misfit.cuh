// k4_misfit.cuh -- the tail of surf_likelihood on the device (reference src/likelihood_surf.F90:227-231, 350,
// 356-404): noise level per ray, Gaussian misfit sums, sum(log(sigma)) -- and the session entry points that chain
// CalGroupTime (k3_raytime.cuh) and the misfit on the RESIDENT maps, so that in the reference's straight-ray mode
// (settings%isStraight == 1, likelihood_surf.F90:233-243) a whole surface-wave likelihood is: nuclei in, three
// doubles out.  Also the posterior accumulation of stat_rti (src/mcmc_loc2.f90:1966-1978) on the resident model.
//
// Bit parity: the reference adds the per-ray terms one by one in (period, source, receiver) order; floating-point
// addition is not associative, so the terms are computed in parallel (misfit_terms_kernel) and then added by ONE
// thread in exactly that order (misfit_reduce_kernel).  A few thousand dependent DADDs: microseconds.
#pragma once

#define MCT_E_ZERO_NOISE_CODE 6

// One thread per (ray, period): sigma and the three squared-residual terms; invalid rays get sigma = 1, terms 0.
// ttime (nrr,3,np): [.,0,.] observed time, [.,1,.] its noise level (sigdep == 0); raystat (nrr,2,np): [.,0,.] == 1
// for a ray that carries data.  flag: set to 1 when a valid ray has sigma < EPS (the reference raises an error).
__global__ void __launch_bounds__(256) misfit_terms_kernel(const double* __restrict__ time, int nrr, int np, int sigdep,
                                                           const double* __restrict__ ttime, const int32_t* __restrict__ raystat,
                                                           const double* __restrict__ snoise, /* [0..np) snoise0, [np..2np) snoise1 */
                                                           const double* __restrict__ srdist, double* __restrict__ sigma,
                                                           double* __restrict__ terms /* (3, nrr*np) */, int32_t* flag) {
  const long long n = (long long)nrr * np;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = (int)(t / nrr);
  const int r = (int)(t - (long long)i * nrr);
  const bool valid = raystat[(size_t)r + (size_t)nrr * 2 * (size_t)i] == 1;
  double sg;
  if (sigdep != 0) sg = valid ? snoise[i] * srdist[t] + snoise[np + i] : 1.0; // :358-370
  else sg = ttime[(size_t)r + (size_t)nrr * (1 + 3 * (size_t)i)];              // like%sigma = dat%ttime(:,2,:)  :372
  double tl = 0.0, tm = 0.0, tu = 0.0;
  if (valid) {
    if (sg < (double)1.0E-10f) *flag = 1; // EPS, :37,:387
    const double d = time[t] - ttime[(size_t)r + (size_t)nrr * 3 * (size_t)i];
    const double d2 = d * d;
    const double s2 = sg * sg;
    tl = d2 / (2 * s2); // :391-392
    tm = d2 / s2;       // :393-394
    tu = d2;            // :395-396
  } else {
    sg = 1.0; // :398
  }
  sigma[t] = sg;
  terms[t] = tl;
  terms[n + t] = tm;
  terms[2 * n + t] = tu;
}

// The sequential sums, in the reference's order.  out[0..2] = like, misfit, unweighted_misfit.
__global__ void misfit_reduce_kernel(const double* __restrict__ sigma, const double* __restrict__ terms, const int32_t* __restrict__ raystat,
                                     int nrr, int np, int nrays_total, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long n = (long long)nrr * np;
  double like = 0.0, mis = 0.0, unw = 0.0, slog = 0.0;
  for (long long t = 0; t < n; ++t) {
    const int i = (int)(t / nrr);
    const int r = (int)(t - (long long)i * nrr);
    if (raystat[(size_t)r + (size_t)nrr * 2 * (size_t)i] == 1) { // rays without data add nothing (not even +0.0)
      like = like + terms[t];
      mis = mis + terms[n + t];
      unw = unw + terms[2 * n + t];
    }
    slog = slog + mct_log(sigma[t]); // sum(log(like%sigma)), array element order, :404
  }
  // like%like + sum(log(like%sigma)) + dat%nrays/2.0 * log(PI2): integer/default-real is default real; PI2 is the
  // default-real literal 6.283185 held in a real(ii10) parameter (:36)
  const double pi2 = (double)6.283185f;
  like = (like + slog) + (double)((float)nrays_total / 2.0f) * mct_log(pi2);
  out[0] = like;
  out[1] = mis;
  out[2] = unw;
}

namespace {

void misfit_release(MisfitBufs& m) {
  DevBuf* b[] = {&m.ttime, &m.raystat, &m.srdist, &m.snoise, &m.sigma, &m.terms, &m.out, &m.time};
  for (DevBuf* x : b) release(*x);
  m.have = false;
}

int misfit_set_data(MisfitBufs& m, int nrr, int np, int sigdep, int nrays_total, const double* ttime, const int32_t* raystat,
                    const double* srdist, cudaStream_t st) {
  if (nrr < 1 || np < 1 || !ttime || !raystat || (sigdep != 0 && !srdist)) return fail(MCT_E_INVALID_ARG, "misfit: bad data arguments");
  const size_t n = (size_t)nrr * np;
  int rc;
  if ((rc = ensure(m.ttime, n * 3 * 8)) || (rc = ensure(m.raystat, n * 2 * 4)) || (rc = ensure(m.srdist, n * 8)) ||
      (rc = ensure(m.snoise, (size_t)np * 2 * 8)) || (rc = ensure(m.sigma, n * 8)) || (rc = ensure(m.terms, n * 3 * 8)) ||
      (rc = ensure(m.out, 4 * 8)) || (rc = ensure(m.time, n * 8)))
    return rc;
  CK(cudaMemcpyAsync(m.ttime.p, ttime, n * 3 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m.raystat.p, raystat, n * 2 * 4, cudaMemcpyHostToDevice, st));
  if (srdist) CK(cudaMemcpyAsync(m.srdist.p, srdist, n * 8, cudaMemcpyHostToDevice, st));
  else CK(cudaMemsetAsync(m.srdist.p, 0, n * 8, st));
  CK(cudaStreamSynchronize(st)); // the caller's arrays may be pageable and short-lived
  m.nrr = nrr; m.np = np; m.sigdep = sigdep; m.nrays_total = nrays_total; m.have = true;
  return MCT_OK;
}

// d_time (nrr,np) on the device -> out[3] (host), optionally sigma (host).  One synchronisation.
int misfit_run(MisfitBufs& m, const double* d_time, const double* snoise0, const double* snoise1, double out[3], double* sigma,
               cudaStream_t st, const double* d_srdist = nullptr /* like%srdist on the device instead of the resident copy */) {
  if (!m.have) return fail(MCT_E_INVALID_ARG, "misfit: no data set");
  if (m.sigdep != 0 && (!snoise0 || !snoise1)) return fail(MCT_E_INVALID_ARG, "misfit: sigdep /= 0 needs snoise0 and snoise1");
  const long long n = (long long)m.nrr * m.np;
  double hs[2 * MCT_MAX_PERIODS];
  if (m.np > MCT_MAX_PERIODS) return fail(MCT_E_INVALID_ARG, "misfit: np must not exceed %d", MCT_MAX_PERIODS);
  for (int i = 0; i < m.np; ++i) { hs[i] = snoise0 ? snoise0[i] : 0.0; hs[m.np + i] = snoise1 ? snoise1[i] : 0.0; }
  CK(cudaMemcpyAsync(m.snoise.p, hs, (size_t)m.np * 2 * 8, cudaMemcpyHostToDevice, st));
  int32_t* flag = (int32_t*)((double*)m.out.p + 3);
  CK(cudaMemsetAsync(flag, 0, 8, st));
  {
    ProfScope ps(2, st);
    misfit_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_time, m.nrr, m.np, m.sigdep, (const double*)m.ttime.p,
                                                                     (const int32_t*)m.raystat.p, (const double*)m.snoise.p,
                                                                     d_srdist ? d_srdist : (const double*)m.srdist.p, (double*)m.sigma.p, (double*)m.terms.p, flag);
    misfit_reduce_kernel<<<1, 32, 0, st>>>((const double*)m.sigma.p, (const double*)m.terms.p, (const int32_t*)m.raystat.p, m.nrr, m.np,
                                           m.nrays_total, (double*)m.out.p);
  }
  CK(cudaGetLastError());
  g.host_stats.n_launches += 2;
  double ho[4];
  CK(cudaMemcpyAsync(ho, m.out.p, sizeof ho, cudaMemcpyDeviceToHost, st));
  if (sigma) CK(cudaMemcpyAsync(sigma, m.sigma.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  out[0] = ho[0]; out[1] = ho[1]; out[2] = ho[2];
  int32_t f;
  memcpy(&f, &ho[3], sizeof f);
  if (f) return fail(MCT_E_ZERO_NOISE_CODE, "misfit: the noise level of a ray that carries data is below 1e-10 (the reference raises 'The noise level is 0!')");
  return MCT_OK;
}

MisfitBufs g_misfit; // the non-session entry point's buffers
void release_misfit_globals() { misfit_release(g_misfit); }

} // namespace

extern "C" {

int mct_surf_misfit(const double* time, int nrr, int np, int sigdep, int nrays_total, const double* ttime, const int32_t* raystat,
                    const double* snoise0, const double* snoise1, const double* srdist, double out[3], double* sigma) {
  NEED_INIT();
  if (!time || !out) return fail(MCT_E_INVALID_ARG, "surf_misfit: NULL pointer");
  cudaStream_t st = g.stream;
  int rc = misfit_set_data(g_misfit, nrr, np, sigdep, nrays_total, ttime, raystat, srdist, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(g_misfit.time.p, time, (size_t)nrr * np * 8, cudaMemcpyHostToDevice, st));
  return misfit_run(g_misfit, (const double*)g_misfit.time.p, snoise0, snoise1, out, sigma, st);
}

// ---- session: rays, data, likelihood, posterior accumulation -------------------------------------------------

// like%gvel = gvel (phaseGroup == 1) or pvel (otherwise), likelihood_surf.F90:226-230
static const double* session_time_map(const mct_session* s) { return (const double*)(s->opt.phaseGroup == 1 ? s->gvel.p : s->pvel.p); }
static int session_overlay(const mct_session* s, int pending, VelOverlay& ov) {
  ov = VelOverlay{nullptr, 0, 0, 0, 0};
  if (!s->have_model || !s->maps_valid) return fail(MCT_E_INVALID_ARG, "session: the current model has no dispersion maps (none set, or check_model rejected it)");
  if (!pending) {
    if (s->pending) return fail(MCT_E_INVALID_ARG, "session: a proposal is pending; accept or reject it first, or ask for the pending model");
    return MCT_OK;
  }
  if (!s->pending) return fail(MCT_E_INVALID_ARG, "session: no pending proposal");
  if (s->pinvalid || s->pcode >= 2) return fail(MCT_E_INVALID_ARG, "session: the pending proposal has no dispersion maps (rejected by check_model or unsupported columns)");
  const int wx = s->pwin[1] - s->pwin[0] + 1, wy = s->pwin[3] - s->pwin[2] + 1;
  if (wx > 0 && wy > 0) ov = VelOverlay{(const double*)(s->opt.phaseGroup == 1 ? s->w_gvel.p : s->w_pvel.p), s->pwin[0], s->pwin[2], wx, wy};
  return MCT_OK;
}

int mct_session_set_rays(mct_session* s, const double* ray_points, const int64_t* ray_offsets, int nrays) {
  NEED_INIT();
  if (!s || !ray_points || !ray_offsets || nrays < 1) return fail(MCT_E_INVALID_ARG, "session_set_rays: bad arguments");
  if (s->nout != s->np) return fail(MCT_E_INVALID_ARG, "session_set_rays: needs a single-mode session");
  int rc = upload_rays(ray_points, ray_offsets, (long long)s->np * nrays, s->r_pts, s->r_off, g.stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(g.stream));
  s->r_nrays = nrays;
  return MCT_OK;
}

int mct_session_set_data(mct_session* s, int nrr, int sigdep, int nrays_total, const double* ttime, const int32_t* raystat,
                         const double* srdist) {
  NEED_INIT();
  if (!s) return fail(MCT_E_INVALID_ARG, "session_set_data: NULL session");
  return misfit_set_data(s->mf, nrr, s->np, sigdep, nrays_total, ttime, raystat, srdist, g.stream);
}

static int session_times(mct_session* s, int pending, const double* ray_points, const int64_t* ray_offsets, int nrays, double** d_time) {
  if (s->nout != s->np) return fail(MCT_E_INVALID_ARG, "session: ray times need a single-mode session (the reference's likelihood uses the fundamental mode)");
  if (s->gr.nx < 2 || s->gr.ny < 2) return fail(MCT_E_INVALID_ARG, "group_times: the bilinear stencil needs nx, ny >= 2");
  VelOverlay ov;
  int rc = session_overlay(s, pending, ov);
  if (rc) return rc;
  cudaStream_t st = g.stream;
  const double* d_pts;
  const long long* d_off;
  if (ray_points) {
    if (!ray_offsets || nrays < 1) return fail(MCT_E_INVALID_ARG, "session: bad ray arguments");
    if ((rc = upload_rays(ray_points, ray_offsets, (long long)s->np * nrays, g.ray_pts, g.ray_off, st))) return rc;
    d_pts = (const double*)g.ray_pts.p; d_off = (const long long*)g.ray_off.p;
  } else {
    if (s->r_nrays < 1) return fail(MCT_E_INVALID_ARG, "session: no resident rays (mct_session_set_rays) and none passed");
    nrays = s->r_nrays;
    d_pts = (const double*)s->r_pts.p; d_off = (const long long*)s->r_off.p;
  }
  if ((rc = ensure(s->time, (size_t)s->np * nrays * 8))) return rc;
  if ((rc = launch_group_times(session_time_map(s), ov, s->np, &s->gr, d_pts, d_off, nrays, (double*)s->time.p, st))) return rc;
  s->time_nrays = nrays;
  *d_time = (double*)s->time.p;
  return MCT_OK;
}

// CalGroupTime on the session's resident map (the accepted model); rays from the host, or the resident ones when
// ray_points == NULL.  The map is like%gvel: the group map when phaseGroup == 1, else the phase map.
int mct_session_group_times(mct_session* s, const double* ray_points, const int64_t* ray_offsets, int nrays, double* time) {
  NEED_INIT();
  if (!s || !time) return fail(MCT_E_INVALID_ARG, "session_group_times: NULL pointer");
  double* d_time = nullptr;
  int rc = session_times(s, 0, ray_points, ray_offsets, nrays, &d_time);
  if (rc) return rc;
  CK(cudaMemcpyAsync(time, d_time, (size_t)s->np * s->time_nrays * 8, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return MCT_OK;
}
// The same for the PENDING proposal: its window maps overlaid on the resident maps (what surf_likelihood sees when
// the sampler evaluates a proposal, mcmc_loc2.f90:228).
int mct_session_group_times_pending(mct_session* s, const double* ray_points, const int64_t* ray_offsets, int nrays, double* time) {
  NEED_INIT();
  if (!s || !time) return fail(MCT_E_INVALID_ARG, "session_group_times_pending: NULL pointer");
  double* d_time = nullptr;
  int rc = session_times(s, 1, ray_points, ray_offsets, nrays, &d_time);
  if (rc) return rc;
  CK(cudaMemcpyAsync(time, d_time, (size_t)s->np * s->time_nrays * 8, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return MCT_OK;
}

// surf_likelihood's tail for the current (pending = 0) or pending (1) model: ray times through like%gvel, sigma,
// the three sums.  Rays as above; data from mct_session_set_data.  phase_time / sigma: optional host outputs (nrr,np).
int mct_session_likelihood(mct_session* s, int pending, const double* ray_points, const int64_t* ray_offsets, int nrays,
                           const double* snoise0, const double* snoise1, double out[3], double* phase_time, double* sigma) {
  NEED_INIT();
  if (!s || !out) return fail(MCT_E_INVALID_ARG, "session_likelihood: NULL pointer");
  if (!s->mf.have) return fail(MCT_E_INVALID_ARG, "session_likelihood: no data (call mct_session_set_data first)");
  double* d_time = nullptr;
  int rc = session_times(s, pending, ray_points, ray_offsets, nrays, &d_time);
  if (rc) return rc;
  if (s->time_nrays != s->mf.nrr) return fail(MCT_E_INVALID_ARG, "session_likelihood: %d rays per period but the data hold %d source-receiver pairs", s->time_nrays, s->mf.nrr);
  if (phase_time) CK(cudaMemcpyAsync(phase_time, d_time, (size_t)s->np * s->time_nrays * 8, cudaMemcpyDeviceToHost, g.stream));
  return misfit_run(s->mf, d_time, snoise0, snoise1, out, sigma, g.stream);
}

// stat_rti (src/mcmc_loc2.f90:1966-1978; same four sums as src/sample.f90:483-486) on the resident CURRENT model.
int mct_session_stat_accumulate(mct_session* s) {
  NEED_INIT();
  if (!s || !s->have_model) return fail(MCT_E_INVALID_ARG, "session_stat_accumulate: no current model");
  if (s->pending) return fail(MCT_E_INVALID_ARG, "session_stat_accumulate: a proposal is pending");
  const size_t nn = (size_t)s->gr.nx * s->gr.ny * s->gr.nz;
  int rc;
  if (!s->acc.p) {
    if ((rc = ensure(s->acc, nn * 4 * 8))) return rc;
    CK(cudaMemsetAsync(s->acc.p, 0, nn * 4 * 8, g.stream));
    s->nacc = 0;
  }
  double* a = (double*)s->acc.p;
  rc = mct_accumulate_stats_dev((const double*)s->vs.p, (const double*)s->vp.p, a, a + nn, a + 2 * nn, a + 3 * nn, (int64_t)nn, g.stream);
  if (rc) return rc;
  s->nacc += 1;
  return MCT_OK;
}
int mct_session_stat_get(mct_session* s, double* aveS, double* stdS, double* aveP, double* stdP, int64_t* nsamples) {
  NEED_INIT();
  if (!s) return fail(MCT_E_INVALID_ARG, "session_stat_get: NULL session");
  const size_t nn = (size_t)s->gr.nx * s->gr.ny * s->gr.nz;
  if (nsamples) *nsamples = s->nacc;
  double* outs[4] = {aveS, stdS, aveP, stdP};
  for (int k = 0; k < 4; ++k) {
    if (!outs[k]) continue;
    if (s->acc.p) CK(cudaMemcpyAsync(outs[k], (double*)s->acc.p + k * nn, nn * 8, cudaMemcpyDeviceToHost, g.stream));
    else memset(outs[k], 0, nn * 8);
  }
  CK(cudaStreamSynchronize(g.stream));
  return MCT_OK;
}
int mct_session_stat_reset(mct_session* s) {
  NEED_INIT();
  if (!s) return fail(MCT_E_INVALID_ARG, "session_stat_reset: NULL session");
  if (s->acc.p) CK(cudaMemsetAsync(s->acc.p, 0, (size_t)s->gr.nx * s->gr.ny * s->gr.nz * 4 * 8, g.stream));
  s->nacc = 0;
  return MCT_OK;
}

} // extern "C"
