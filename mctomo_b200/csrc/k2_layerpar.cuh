// k2_layerpar.cuh -- the Rayleigh secular function with its LAYERS spread over the lanes of a warp.
//
// Used by the warp-per-column kernel (k2_coop.cuh) for the evaluations it cannot parallelise over trial
// velocities: nevill's refinement and the first evaluation of every getsol (reference surfdisp96.f:903-1020, 774),
// which are strictly sequential in c.  Inside ONE evaluation, though, two thirds of the work of a layer step --
// the eigenfunction products of `var` and Dunkin's matrix `dnka` (surfdisp96.f:1220-1337, 1370-1414) -- depend on
// the layer and on c but not on the running vector e.  So lane j computes the 15 distinct matrix entries of layer
// (mmax-2-j) while the other lanes do theirs; the recursion e <- normalise(e * ca) then walks the layers in order,
// fetching each matrix from its lane with shuffles.  Per evaluation the dependent chain shrinks from
// L x (whole layer step) to one matrix + L x (5x5 product + normalisation): ~4-5x shorter for 10 layers.
// Same IEEE operations in the same order as layer_step_fast / dltar4_dev; any step whose operands leave the range of
// the fast division makes the whole evaluation fall back to the ordinary one (returns false).
#pragma once

struct CaMat { double c11, c12, c13, c14, c15, c21, c22, c23, c24, c33, c41, c42, c43, c51, c53; };

// First half of layer_step_fast: everything that does not involve the running vector.  false = declined.
__device__ __forceinline__ bool compute_ca(const float4 L, const double4 Rc, double wvno, double wvno2, double omega, double y_om,
                                           CaMat& M, RangeTrack& R) {
  const double a = (double)L.y, b = (double)L.z, dpth = (double)L.x, rho = (double)L.w;
  const double rho2 = rho * rho;
  const double y_a = Rc.x, y_b = Rc.y, y_rho = Rc.z, y_rho2 = Rc.w;
  // (a, b, rho, rho2 are range-checked once per layer by layer_recips_kernel: an out-of-range constant poisons
  //  y_a with NaN, which reaches ra below and sends the step to the exact path)
  const double xka = mct_div_r(omega, a, y_a);
  const double xkb = mct_div_r(omega, b, y_b);
  const double t = mct_div_r(b, omega, y_om);
  const double gammk = 2.0 * t * t;
  const double gam = gammk * wvno2;
  const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
  const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
  const double p = ra * dpth;
  const double q = rb * dpth;
  const double y_ra = mct_rcp(ra), y_rb = mct_rcp(rb);
  R.add(ra); R.add(rb); // zero (wvno == xk: the reference's equality branch) falls out of range -> exact path
  double cosp, w, x, cosq, y, z, a0;
  const bool posc = wvno < xka, sosc = wvno < xkb;
  if (posc && !sosc) return false;
  if (!sosc) {
    // P and S both evanescent (:1284-1291, :1306-1313): three independent exponentials
    const bool np_ = p < 16.0, nq_ = q < 16.0;
    const double exa = p + q;
    const bool na_ = exa < 60.0;
    double facp = 0.0, facq = 0.0;
    a0 = 0.0;
    if (!(p >= 16.0 && q >= 16.0 && exa >= 60.0)) { // (chained compares; `np_ | nq_ | na_` made the compiler build min(p, q) with NaN handling)
      const double fp = mct_exp_core(np_ ? -2.0 * p : -1.0);
      const double fq = mct_exp_core(nq_ ? -2.0 * q : -1.0);
      const double fa = mct_exp_core(na_ ? -exa : -1.0);
      facp = np_ ? fp : 0.0;
      facq = nq_ ? fq : 0.0;
      a0 = na_ ? fa : 0.0;
    }
    cosp = (1.0 + facp) * 0.5;
    const double sinp = (1.0 - facp) * 0.5;
    cosq = (1.0 + facq) * 0.5;
    const double sinq = (1.0 - facq) * 0.5;
    R.add(sinp); R.add(sinq);
    w = mct_div_r(sinp, ra, y_ra);
    x = ra * sinp;
    y = mct_div_r(sinq, rb, y_rb);
    z = rb * sinq;
  } else if (!posc) {
    // P evanescent, S oscillatory (:1284-1291, :1297-1301)
    const bool np_ = p < 16.0;
    const bool na_ = p < 60.0; // exa = pex + sex = p + 0
    const double fp = mct_exp_core(np_ ? -2.0 * p : -1.0);
    const double fa = mct_exp_core(na_ ? -(p + 0.0) : -1.0);
    double sinq;
    mct_sincos(q, &sinq, &cosq);
    const double facp = np_ ? fp : 0.0;
    a0 = na_ ? fa : 0.0;
    cosp = (1.0 + facp) * 0.5;
    const double sinp = (1.0 - facp) * 0.5;
    R.add(sinp); R.add(sinq);
    w = mct_div_r(sinp, ra, y_ra);
    x = ra * sinp;
    y = mct_div_r(sinq, rb, y_rb);
    z = -rb * sinq;
  } else {
    // both oscillatory (:1275-1279, :1297-1301); exa = 0 -> a0 = exp(-0) = 1 exactly
    double sinp, sinq;
    mct_sincos(p, &sinp, &cosp);
    mct_sincos(q, &sinq, &cosq);
    a0 = 1.0;
    R.add(sinp); R.add(sinq);
    w = mct_div_r(sinp, ra, y_ra);
    x = -ra * sinp;
    y = mct_div_r(sinq, rb, y_rb);
    z = -rb * sinq;
  }
  const double cpcq = cosp * cosq, cpy = cosp * y, cpz = cosp * z, cqw = cosq * w, cqx = cosq * x;
  const double xy = x * y, xz = x * z, wy = w * y, wz = w * z;
  const double gamm1 = gam - 1.0;
  const double twgm1 = gam + gamm1;
  const double gmgmk = gam * gammk;
  const double gmgm1 = gam * gamm1;
  const double gm1sq = gamm1 * gamm1;
  const double a0pq = a0 - cpcq;
  const double ca11 = cpcq - 2.0 * gmgm1 * a0pq - gmgmk * xz - wvno2 * gm1sq * wy;
  const double n12 = wvno2 * cpy - cqx;
  const double n13 = -(twgm1 * a0pq + gammk * xz + wvno2 * gamm1 * wy);
  const double n14 = cpz - wvno2 * cqw;
  const double n15 = -(2.0 * wvno2 * a0pq + xz + wvno2 * wvno2 * wy);
  R.add(n12); R.add(n13); R.add(n14); R.add(n15);
  const double ca12 = mct_div_r(n12, rho, y_rho);
  const double ca13 = mct_div_r(n13, rho, y_rho);
  const double ca14 = mct_div_r(n14, rho, y_rho);
  const double ca15 = mct_div_r(n15, rho2, y_rho2);
  const double ca21 = (gmgmk * cpz - gm1sq * cqw) * rho;
  const double ca22 = cpcq;
  const double ca23 = gammk * cpz - gamm1 * cqw;
  const double ca24 = -wz;
  const double ca41 = (gm1sq * cpy - gmgmk * cqx) * rho;
  const double ca42 = -xy;
  const double ca43 = gamm1 * cpy - gammk * cqx;
  const double ca51 = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2;
  const double ca53 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho;
  M.c11 = ca11; M.c12 = ca12; M.c13 = ca13; M.c14 = ca14; M.c15 = ca15;
  M.c21 = ca21; M.c22 = ca22; M.c23 = ca23; M.c24 = ca24; M.c33 = a0 + 2.0 * (cpcq - ca11);
  M.c41 = ca41; M.c42 = ca42; M.c43 = ca43; M.c51 = ca51; M.c53 = ca53;
  return true;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Second half: e <- normalise(e * ca) for the matrix held by lane `src`.  Uniform across the warp.
__device__ __forceinline__ bool apply_ca_from(const CaMat& Mine, int src, double wvno2, EVec& E) {
  const double ca11 = shfl_d(Mine.c11, src), ca12 = shfl_d(Mine.c12, src), ca13 = shfl_d(Mine.c13, src);
  const double ca14 = shfl_d(Mine.c14, src), ca15 = shfl_d(Mine.c15, src), ca21 = shfl_d(Mine.c21, src);
  const double ca22 = shfl_d(Mine.c22, src), ca23 = shfl_d(Mine.c23, src), ca24 = shfl_d(Mine.c24, src);
  const double ca33 = shfl_d(Mine.c33, src), ca41 = shfl_d(Mine.c41, src), ca42 = shfl_d(Mine.c42, src);
  const double ca43 = shfl_d(Mine.c43, src), ca51 = shfl_d(Mine.c51, src), ca53 = shfl_d(Mine.c53, src);
  const double tt = -2.0 * wvno2;
  const double ca31 = tt * ca53, ca32 = tt * ca43, ca34 = tt * ca23, ca35 = tt * ca13;
  const double e1 = E.e1, e2 = E.e2, e3 = E.e3, e4 = E.e4, e5 = E.e5;
  double ee1 = e1 * ca11; ee1 = ee1 + e2 * ca21; ee1 = ee1 + e3 * ca31; ee1 = ee1 + e4 * ca41; ee1 = ee1 + e5 * ca51;
  double ee2 = e1 * ca12; ee2 = ee2 + e2 * ca22; ee2 = ee2 + e3 * ca32; ee2 = ee2 + e4 * ca42; ee2 = ee2 + e5 * ca41;
  double ee3 = e1 * ca13; ee3 = ee3 + e2 * ca23; ee3 = ee3 + e3 * ca33; ee3 = ee3 + e4 * ca43; ee3 = ee3 + e5 * ca53;
  double ee4 = e1 * ca14; ee4 = ee4 + e2 * ca24; ee4 = ee4 + e3 * ca34; ee4 = ee4 + e4 * ca22; ee4 = ee4 + e5 * ca21;
  double ee5 = e1 * ca15; ee5 = ee5 + e2 * ca14; ee5 = ee5 + e3 * ca35; ee5 = ee5 + e4 * ca12; ee5 = ee5 + e5 * ca11;
  double t1 = dmax_nn(dmax_nn(dmax_nn(fabs(ee1), fabs(ee2)), dmax_nn(fabs(ee3), fabs(ee4))), fabs(ee5));
  if (t1 < 1.e-40) t1 = 1.0;
  const double y_t1 = mct_rcp(t1);
  // (starting the reciprocal for all five candidates before the maximum is known shortens the dependent chain by
  //  three compare/select levels but adds 20 FP64 instructions; measured: no gain)
  RangeTrack R;
  /* t1 is one of the |ee| (or 1.0) */ R.add(ee1); R.add(ee2); R.add(ee3); R.add(ee4); R.add(ee5);
  if (!R.ok()) return false;
  E.e1 = mct_div_r(ee1, t1, y_t1);
  E.e2 = mct_div_r(ee2, t1, y_t1);
  E.e3 = mct_div_r(ee3, t1, y_t1);
  E.e4 = mct_div_r(ee4, t1, y_t1);
  E.e5 = mct_div_r(ee5, t1, y_t1);
  return true;
}

// dltar4 for ONE trial wavenumber, all 32 lanes of the warp cooperating.  Every lane must call it with the same
// arguments; every lane gets the same result.  Returns false if the evaluation must be redone the ordinary way.
__device__ __noinline__ bool dltar4_layerpar_dev(const float4* __restrict__ lay, const double4* __restrict__ layr, int stride, int mmax,
                                                 int llw, double wvno, double omga, double& del) {
  const int lane = threadIdx.x & 31;
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  if (!mct_exp_ok(omega)) return false;
  const double wvno2 = wvno * wvno;
  const double y_om = mct_rcp(omega);
  EVec E;
  {
    const float4 L = __ldg(&lay[(size_t)(mmax - 1) * stride]);
    const double4 Rh = layr[(size_t)(mmax - 1) * stride];
    // the three divisions of the half-space start vector through the reciprocal table when it vouches for the
    // layer constants (a NaN first entry means it does not) -- same bits as `/`, a fraction of its latency
    const bool tab = (true /* omega checked above */) && Rh.x == Rh.x;
    const double xka = tab ? mct_div_r(omega, (double)L.y, Rh.x) : omega / (double)L.y;
    const double xkb = tab ? mct_div_r(omega, (double)L.z, Rh.y) : omega / (double)L.z;
    const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    const double t = tab ? mct_div_r((double)L.z, omega, y_om) : (double)L.z / omega;
    const double gammk = 2.0 * t * t;
    const double gam = gammk * wvno2;
    const double gamm1 = gam - 1.0;
    const double rho1 = (double)L.w;
    E.e1 = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
    E.e2 = -rho1 * ra;
    E.e3 = rho1 * (gamm1 - gammk * ra * rb);
    E.e4 = rho1 * rb;
    E.e5 = wvno2 - ra * rb;
  }
  const int nsteps = mmax - llw; // layers mmax-2 .. llw-1
  for (int base = 0; base < nsteps; base += 32) {
    const int j = base + lane;
    const bool have = j < nsteps;
    CaMat M;
    M.c11 = M.c12 = M.c13 = M.c14 = M.c15 = M.c21 = M.c22 = M.c23 = M.c24 = M.c33 = M.c41 = M.c42 = M.c43 = M.c51 = M.c53 = 0.0;
    bool okA = true;
    if (have) {
      const int m = mmax - 2 - j;
      RangeTrack R;
      okA = compute_ca(__ldg(&lay[(size_t)m * stride]), layr[(size_t)m * stride], wvno, wvno2, omega, y_om, M, R) && R.ok();
    }
    if (!__all_sync(0xffffffffu, okA)) return false;
    const int cnt = min(32, nsteps - base);
    for (int jj = 0; jj < cnt; ++jj)
      if (!apply_ca_from(M, jj, wvno2, E)) return false;
  }
  if (llw != 1) {
    const float4 Lw = __ldg(&lay[0]);
    const double xka = omega / (double)Lw.y;
    const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    const double dpth = (double)Lw.x;
    const double rho1 = (double)Lw.w;
    const double p = ra * dpth;
    double cosp, w, x, pex;
    eig_pair(p, ra, wvno, xka, dpth, cosp, w, x, pex);
    const double w0 = -rho1 * w;
    del = cosp * E.e1 + w0 * E.e2;
  } else {
    del = E.e1;
  }
  return true;
}
