// k2_rayleigh_fast.cuh -- the production form of the Rayleigh secular function (dltar4,
// reference surfmodes/surfdisp96.f:1119-1217 with var :1220-1337, dnka :1370-1414, normc :1341-1366).
//
// Same IEEE operations, same order, same results as dltar4_dev in k2_dispersion.cuh (which stays
// as the plainly written form and is the fallback for any layer step this file declines); what
// changes is the SHAPE of the instruction stream, driven by the ncu profiles under profiles/:
//   * dependent-DFMA latency, not FP64 issue rate, bounded the first version ("wait" stalls), so
//     a layer step is organised as one of three straight-line cases -- (P,S) both evanescent, P
//     evanescent + S oscillatory, both oscillatory -- in which the two or three transcendental
//     evaluations are independent instruction streams the scheduler can interleave;
//   * the 14 divisions of a layer step are done with refined reciprocals (mct_rcp / mct_div_r):
//     one reciprocal serves all quotients with the same denominator, and there is ONE range check per
//     layer step (integer min/max over the operands' exponent words) instead of a branch per
//     division;
//   * a step whose operands leave the range in which the fast division is provably exact (zeros,
//     equalities wvno == omega/alpha, absurd magnitudes) is redone by layer_step_exact().
#pragma once

struct EVec { double e1, e2, e3, e4, e5; };

// max of two doubles that are known not to be NaN when the result is used (a NaN operand fails the step's range check):
// one DSETP and two selects, where fmax() costs eight instructions for its NaN semantics
__device__ __forceinline__ double dmax_nn(double a, double b) { return a > b ? a : b; }

// exponent-word tracker: all operands of fast divisions must be normal with |x| in [2^-400, 2^400]
struct RangeTrack {
  unsigned lo, hi;
  __device__ __forceinline__ RangeTrack() : lo(0xffffffffu), hi(0u) {}
  __device__ __forceinline__ void add(double v) {
    const unsigned h = ((unsigned)__double2hiint(v)) << 1; // drop the sign, keep exponent + top mantissa bits
    lo = min(lo, h);
    hi = max(hi, h);
  }
  __device__ __forceinline__ bool ok() const {
    return lo >= ((1023u - 400u) << 21) && hi < ((1023u + 400u) << 21);
  }
};

// ---- the plainly written layer step (exact path) ------------------------------------------------------
// (E goes in and out BY VALUE: taking its address for a non-inlined call would pin the five running
// values to local memory for every step of the hot loop.)
__device__ __noinline__ EVec layer_step_exact(const float4 L, double wvno, double wvno2, double omega, const EVec E) {
  const double xka = omega / (double)L.y;
  const double xkb = omega / (double)L.z;
  const double t = (double)L.z / omega;
  const double gammk = 2.0 * t * t;
  const double gam = gammk * wvno2;
  const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
  const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
  const double dpth = (double)L.x;
  const double rho = (double)L.w;
  const double p = ra * dpth;
  const double q = rb * dpth;
  double cosp, w, x, pex, cosq, y, z, sex;
  eig_pair(p, ra, wvno, xka, dpth, cosp, w, x, pex);
  eig_pair(q, rb, wvno, xkb, dpth, cosq, y, z, sex);
  const double exa = pex + sex;
  double a0 = 0.0;
  if (exa < 60.0) a0 = mct_exp(-exa);
  const double cpcq = cosp * cosq, cpy = cosp * y, cpz = cosp * z, cqw = cosq * w, cqx = cosq * x;
  const double xy = x * y, xz = x * z, wy = w * y, wz = w * z;
  const double gamm1 = gam - 1.0;
  const double twgm1 = gam + gamm1;
  const double gmgmk = gam * gammk;
  const double gmgm1 = gam * gamm1;
  const double gm1sq = gamm1 * gamm1;
  const double rho2 = rho * rho;
  const double a0pq = a0 - cpcq;
  const double ca11 = cpcq - 2.0 * gmgm1 * a0pq - gmgmk * xz - wvno2 * gm1sq * wy;
  const double ca12 = (wvno2 * cpy - cqx) / rho;
  const double ca13 = -(twgm1 * a0pq + gammk * xz + wvno2 * gamm1 * wy) / rho;
  const double ca14 = (cpz - wvno2 * cqw) / rho;
  const double ca15 = -(2.0 * wvno2 * a0pq + xz + wvno2 * wvno2 * wy) / rho2;
  const double ca21 = (gmgmk * cpz - gm1sq * cqw) * rho;
  const double ca22 = cpcq;
  const double ca23 = gammk * cpz - gamm1 * cqw;
  const double ca24 = -wz;
  const double ca41 = (gm1sq * cpy - gmgmk * cqx) * rho;
  const double ca42 = -xy;
  const double ca43 = gamm1 * cpy - gammk * cqx;
  const double ca51 = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2;
  const double ca53 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho;
  const double tt = -2.0 * wvno2;
  const double ca31 = tt * ca53, ca32 = tt * ca43, ca33 = a0 + 2.0 * (cpcq - ca11), ca34 = tt * ca23, ca35 = tt * ca13;
  const double e1 = E.e1, e2 = E.e2, e3 = E.e3, e4 = E.e4, e5 = E.e5;
  double ee1 = 0.0 + e1 * ca11; ee1 = ee1 + e2 * ca21; ee1 = ee1 + e3 * ca31; ee1 = ee1 + e4 * ca41; ee1 = ee1 + e5 * ca51;
  double ee2 = 0.0 + e1 * ca12; ee2 = ee2 + e2 * ca22; ee2 = ee2 + e3 * ca32; ee2 = ee2 + e4 * ca42; ee2 = ee2 + e5 * ca41;
  double ee3 = 0.0 + e1 * ca13; ee3 = ee3 + e2 * ca23; ee3 = ee3 + e3 * ca33; ee3 = ee3 + e4 * ca43; ee3 = ee3 + e5 * ca53;
  double ee4 = 0.0 + e1 * ca14; ee4 = ee4 + e2 * ca24; ee4 = ee4 + e3 * ca34; ee4 = ee4 + e4 * ca22; ee4 = ee4 + e5 * ca21;
  double ee5 = 0.0 + e1 * ca15; ee5 = ee5 + e2 * ca14; ee5 = ee5 + e3 * ca35; ee5 = ee5 + e4 * ca12; ee5 = ee5 + e5 * ca11;
  double t1 = 0.0;
  if (fabs(ee1) > t1) t1 = fabs(ee1);
  if (fabs(ee2) > t1) t1 = fabs(ee2);
  if (fabs(ee3) > t1) t1 = fabs(ee3);
  if (fabs(ee4) > t1) t1 = fabs(ee4);
  if (fabs(ee5) > t1) t1 = fabs(ee5);
  if (t1 < 1.e-40) t1 = 1.0;
  EVec O;
  O.e1 = ee1 / t1; O.e2 = ee2 / t1; O.e3 = ee3 / t1; O.e4 = ee4 / t1; O.e5 = ee5 / t1;
  return O;
}

// ---- the fast layer step ----------------------------------------------------------------------------
// Returns false (E untouched) when the step must be redone by layer_step_exact.
// Rc = {1/alpha, 1/beta, 1/rho, 1/rho^2}: the refined reciprocals of this layer's constants, computed once per
// forward evaluation by layer_recips_kernel with the same mct_rcp sequence (bit-identical to computing them here).
// L, Rc: this layer's records on entry; when `more`, the NEXT layer's records (from nl, nr) on exit -- each reloaded
// right after its last use, so that the prefetch lands in the registers it vacates (a separate "next" copy cost 12
// register moves per step).
__device__ __forceinline__ bool layer_step_fast(float4& L, double4& Rc, const float4* __restrict__ nl, const double4* __restrict__ nr, bool more,
                                                double wvno, double wvno2, double omega, double y_om, EVec& E) {
  RangeTrack R;
  const double a = (double)L.y, b = (double)L.z, dpth = (double)L.x, rho = (double)L.w;
  const double rho2 = rho * rho;
  const double y_a = Rc.x, y_b = Rc.y, y_rho = Rc.z, y_rho2 = Rc.w;
  if (more) L = __ldg(nl);
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
  if (posc && !sosc) return false; // only possible when alpha < beta: not worth a fast case
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
  if (more) Rc = *nr;
  const double ca21 = (gmgmk * cpz - gm1sq * cqw) * rho;
  const double ca22 = cpcq;
  const double ca23 = gammk * cpz - gamm1 * cqw;
  const double ca24 = -wz;
  const double ca41 = (gm1sq * cpy - gmgmk * cqx) * rho;
  const double ca42 = -xy;
  const double ca43 = gamm1 * cpy - gammk * cqx;
  const double ca51 = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2;
  const double ca53 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho;
  const double tt = -2.0 * wvno2;
  const double ca31 = tt * ca53, ca32 = tt * ca43, ca33 = a0 + 2.0 * (cpcq - ca11), ca34 = tt * ca23, ca35 = tt * ca13;
  // ee(i) = sum_j e(j)*ca(j,i), left to right from 0 (:1184-1190); ca(2,5)=ca(1,4), ca(4,4)=ca(2,2), ca(4,5)=ca(1,2),
  // ca(5,2)=ca(4,1), ca(5,4)=ca(2,1), ca(5,5)=ca(1,1)
  const double e1 = E.e1, e2 = E.e2, e3 = E.e3, e4 = E.e4, e5 = E.e5;
  // (the reference starts every sum from 0.0; 0.0 + x differs from x only for x = -0.0, and that can matter only
  //  if the whole sum stays zero -- in which case the range check below rejects the step: dropping the add is exact)
  double ee1 = e1 * ca11; ee1 = ee1 + e2 * ca21; ee1 = ee1 + e3 * ca31; ee1 = ee1 + e4 * ca41; ee1 = ee1 + e5 * ca51;
  double ee2 = e1 * ca12; ee2 = ee2 + e2 * ca22; ee2 = ee2 + e3 * ca32; ee2 = ee2 + e4 * ca42; ee2 = ee2 + e5 * ca41;
  double ee3 = e1 * ca13; ee3 = ee3 + e2 * ca23; ee3 = ee3 + e3 * ca33; ee3 = ee3 + e4 * ca43; ee3 = ee3 + e5 * ca53;
  double ee4 = e1 * ca14; ee4 = ee4 + e2 * ca24; ee4 = ee4 + e3 * ca34; ee4 = ee4 + e4 * ca22; ee4 = ee4 + e5 * ca21;
  double ee5 = e1 * ca15; ee5 = ee5 + e2 * ca14; ee5 = ee5 + e3 * ca35; ee5 = ee5 + e4 * ca12; ee5 = ee5 + e5 * ca11;
  // normc (:1350-1360): max |ee|, then five quotients by the same scale
  double t1 = dmax_nn(dmax_nn(dmax_nn(fabs(ee1), fabs(ee2)), dmax_nn(fabs(ee3), fabs(ee4))), fabs(ee5));
  if (t1 < 1.e-40) t1 = 1.0;
  const double y_t1 = mct_rcp(t1);
  /* t1 is one of the |ee| (or 1.0) */ R.add(ee1); R.add(ee2); R.add(ee3); R.add(ee4); R.add(ee5);
  if (!R.ok()) return false;
  E.e1 = mct_div_r(ee1, t1, y_t1);
  E.e2 = mct_div_r(ee2, t1, y_t1);
  E.e3 = mct_div_r(ee3, t1, y_t1);
  E.e4 = mct_div_r(ee4, t1, y_t1);
  E.e5 = mct_div_r(ee5, t1, y_t1);
  return true;
}

// ---- dltar4: half-space start vector, layer recursion bottom -> top, optional water layer ---------------
__device__ __noinline__ double dltar4_fast_dev(const float4* __restrict__ lay, const double4* __restrict__ layr, int stride,
                                               int mmax, int llw, double wvno, double omga) {
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  const double wvno2 = wvno * wvno;
  const double y_om = mct_rcp(omega);
  const bool om_ok = mct_exp_ok(omega);
  EVec E;
  {
    const float4 L = __ldg(&lay[(size_t)(mmax - 1) * stride]);
    const double4 Rh = layr[(size_t)(mmax - 1) * stride];
    // the three divisions of the half-space start vector through the reciprocal table when it vouches for the
    // layer constants (a NaN first entry means it does not) -- same bits as `/`, a fraction of its latency
    const bool tab = (om_ok) && Rh.x == Rh.x;
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
  // Hot loop: fast steps only.  A step that declines (rare: zero operands, wvno == omega/alpha, absurd
  // magnitudes) leaves the loop; the remaining layers are then finished by the plainly written step in a
  // separate cold loop, which keeps the non-inlined call and its register shuffling out of the hot body.
  int m = mmax - 2;
  if (om_ok) {
    float4 L = __ldg(&lay[(size_t)max(m, 0) * stride]);
    double4 Rc = layr[(size_t)max(m, 0) * stride];
    const float4* nl = lay + (size_t)max(m - 1, 0) * stride;   // next layer's records: running pointers, one subtraction
    const double4* nr = layr + (size_t)max(m - 1, 0) * stride; // per step instead of a 64-bit multiply-add each
#pragma unroll 1
    for (; m >= llw - 1; --m) { // the next layer's records are fetched inside the step
      if (!layer_step_fast(L, Rc, nl, nr, m > 0, wvno, wvno2, omega, y_om, E)) break;
      if (m > 1) { nl -= stride; nr -= stride; }
    }
  }
#pragma unroll 1
  for (; m >= llw - 1; --m) {
    E = layer_step_exact(__ldg(&lay[(size_t)m * stride]), wvno, wvno2, omega, E);
    // (after one exact step the fast path could resume; not worth the control flow for a rare event)
  }
  if (llw != 1) {
    // water layer on top (:1196-1212): var(p, znul, ra, znul, wvno, xka, znul, dpth, ...)
    const float4 Lw = __ldg(&lay[0]);
    const double xka = omega / (double)Lw.y;
    const double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    const double dpth = (double)Lw.x;
    const double rho1 = (double)Lw.w;
    const double p = ra * dpth;
    double cosp, w, x, pex;
    eig_pair(p, ra, wvno, xka, dpth, cosp, w, x, pex);
    const double w0 = -rho1 * w;
    return cosp * E.e1 + w0 * E.e2;
  }
  return E.e1;
}
