// k2_love_fast.cuh -- the production form of the Love secular function (dltar1, reference
// surfmodes/surfdisp96.f:1056-1115): same IEEE operations and results as dltar1_dev in k2_dispersion.cuh,
// organised like k2_rayleigh_fast.cuh -- refined reciprocals instead of five compiler divisions per layer
// step, two straight-line cases (S oscillatory / S evanescent), one operand-range check per step, the
// plainly written step as fallback.
#pragma once

struct E2 { double e1, e2; };

__device__ __noinline__ E2 love_step_exact(const float4 L, double wvno, double omega, const E2 E) {
  const double beta1 = (double)L.z;
  const double rho1 = (double)L.w;
  const double dm = (double)L.x;
  const double xmu = rho1 * beta1 * beta1;
  const double xkb = omega / beta1;
  const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
  const double q = dm * rb;
  double cosq, y, z, ex;
  eig_pair(q, rb, wvno, xkb, dm, cosq, y, z, ex);
  const double e10 = E.e1 * cosq + E.e2 * xmu * z;
  const double e20 = E.e1 * y / xmu + E.e2 * cosq;
  double xnor = fabs(e10);
  const double ynor = fabs(e20);
  if (ynor > xnor) xnor = ynor;
  if (xnor < 1.e-40) xnor = 1.0;
  E2 O;
  O.e1 = e10 / xnor;
  O.e2 = e20 / xnor;
  return O;
}

// Rc = {1/beta, 1/(rho beta^2), -, -} from layer_recips_kernel
// L, Rc: this layer's records on entry, the next layer's on exit when `more` (reloaded in place, see layer_step_fast)
__device__ __forceinline__ bool love_step_fast(float4& L, double2& Rc, const float4* __restrict__ nl, const double4* __restrict__ nr, bool more,
                                               double wvno, double omega, E2& E) {
  RangeTrack R;
  const double b = (double)L.z, rho = (double)L.w, dm = (double)L.x;
  const double xmu = rho * b * b;
  const double y_b = Rc.x, y_mu = Rc.y;
  if (more) {
    L = __ldg(nl);
    Rc = *reinterpret_cast<const double2*>(nr); // only the first two of the four table entries are used by Love
  }
  const double xkb = mct_div_r(omega, b, y_b);
  const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
  const double q = dm * rb;
  const double y_rb = mct_rcp(rb);
  R.add(rb); // b, xmu: checked by layer_recips_kernel (NaN-poisoned y_b reaches rb);  rb = 0 (wvno == xkb, the reference's equality branch) -> exact path
  double cosq, sinq, z;
  if (wvno < xkb) {
    mct_sincos(q, &sinq, &cosq);
    z = -rb * sinq;
  } else {
    const bool nq_ = q < 16.0;
    const double f = mct_exp_core(nq_ ? -2.0 * q : -1.0);
    const double fac = nq_ ? f : 0.0;
    cosq = (1.0 + fac) * 0.5;
    sinq = (1.0 - fac) * 0.5;
    z = rb * sinq;
  }
  R.add(sinq);
  const double y = mct_div_r(sinq, rb, y_rb);
  const double e10 = E.e1 * cosq + E.e2 * xmu * z;
  const double n20 = E.e1 * y;
  R.add(n20);
  const double e20 = mct_div_r(n20, xmu, y_mu) + E.e2 * cosq;
  double xnor = dmax_nn(fabs(e10), fabs(e20));
  if (xnor < 1.e-40) xnor = 1.0;
  const double y_n = mct_rcp(xnor);
  R.add(e10); R.add(e20); // xnor is one of them (or 1.0)
  if (!R.ok()) return false;
  E.e1 = mct_div_r(e10, xnor, y_n);
  E.e2 = mct_div_r(e20, xnor, y_n);
  return true;
}

__device__ __noinline__ double dltar1_fast_dev(const float4* __restrict__ lay, const double4* __restrict__ layr, int stride,
                                               int mmax, int llw, double wvno, double omega) {
  const bool om_ok = mct_exp_ok(omega);
  E2 E;
  {
    const float4 L = __ldg(&lay[(size_t)(mmax - 1) * stride]);
    const double beta1 = (double)L.z;
    const double rho1 = (double)L.w;
    const double4 Rh = layr[(size_t)(mmax - 1) * stride];
    const double xkb = (om_ok && Rh.x == Rh.x) ? mct_div_r(omega, beta1, Rh.x) : omega / beta1; // table-vouched: same bits as `/`
    const double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    E.e1 = rho1 * rb;
    E.e2 = 1.0 / (beta1 * beta1);
  }
  int m = mmax - 2;
  if (om_ok) {
    float4 L = __ldg(&lay[(size_t)max(m, 0) * stride]);
    double2 Rc = *reinterpret_cast<const double2*>(&layr[(size_t)max(m, 0) * stride]);
    const float4* nl = lay + (size_t)max(m - 1, 0) * stride;
    const double4* nr = layr + (size_t)max(m - 1, 0) * stride;
#pragma unroll 1
    for (; m >= llw - 1; --m) {
      if (!love_step_fast(L, Rc, nl, nr, m > 0, wvno, omega, E)) break;
      if (m > 1) { nl -= stride; nr -= stride; }
    }
  }
#pragma unroll 1
  for (; m >= llw - 1; --m) E = love_step_exact(__ldg(&lay[(size_t)m * stride]), wvno, omega, E); // cold
  return E.e1;
}
