// k3_raytime.cuh -- CalGroupTime / GetVelocity on the device (reference src/likelihood_surf.F90:454-521).
// once the dispersion maps stay in HBM (session, _dev entry points) the group-velocity map never has to
// travel to the host: the rays fm2d produced on the host (a few thousand points) come in, the travel times go out.
// One thread per (ray, period); the accumulation along a ray is sequential, in the reference's order, so the times
// are the reference's bit for bit (every operation is an IEEE add/mul/div/sqrt; the library is built without FMA
// contraction).
#pragma once

// Optional overlay: a packed (np, wy, wx) window (a session's pending proposal) that replaces the resident map inside
// columns ix0..ix0+wx-1, iy0..iy0+wy-1 (1-based); wx = 0: no overlay.
struct VelOverlay {
  const double* w;
  int ix0, iy0, wx, wy;
};

__device__ __forceinline__ double vel_at(const double* __restrict__ vel, const VelOverlay& ov, int np, int ip, int ny, int iy /*1-based*/,
                                         int ix /*1-based*/) {
  const int ox = ix - ov.ix0, oy = iy - ov.iy0;
  if ((unsigned)ox < (unsigned)ov.wx && (unsigned)oy < (unsigned)ov.wy)
    return ov.w[(size_t)ip + (size_t)np * ((size_t)oy + (size_t)ov.wy * (size_t)ox)];
  return vel[(size_t)ip + (size_t)np * ((size_t)(iy - 1) + (size_t)ny * (size_t)(ix - 1))];
}

__device__ __forceinline__ double get_velocity_dev(const double* __restrict__ vel, const VelOverlay& ov, int np, int ip, int nx, int ny,
                                                   double xmin, double ymin, double dx, double dy, double px, double py) {
  int ix = (int)floor((px - xmin) / dx) + 1;
  int iy = (int)floor((py - ymin) / dy) + 1;
  if (ix < 1) ix = 1;
  if (iy < 1) iy = 1;
  if (ix >= nx) ix = nx - 1;
  if (iy >= ny) iy = ny - 1;
  const double dsx = px - (xmin + (double)(ix - 1) * dx);
  const double dsy = py - (ymin + (double)(iy - 1) * dy);
  double qv = 0.0;
#pragma unroll
  for (int i = 1; i <= 2; ++i)
#pragma unroll
    for (int j = 1; j <= 2; ++j) {
      const double weight = (1.0 - fabs((double)(i - 1) * dx - dsx) / dx) * (1.0 - fabs((double)(j - 1) * dy - dsy) / dy);
      qv = qv + weight * vel_at(vel, ov, np, ip, ny, iy + j - 1, ix + i - 1);
    }
  return qv;
}

__global__ void __launch_bounds__(128) group_times_kernel(const double* __restrict__ vel, const VelOverlay ov, int np, int nx, int ny,
                                                          double xmin, double ymin, double dx, double dy, const double* __restrict__ pts,
                                                          const long long* __restrict__ off, int nrays, double* __restrict__ time) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)np * nrays) return;
  const int ip = (int)(t / nrays);
  const long long a = off[t], b = off[t + 1];
  double acc = 0.0;
  if (b - a >= 2) {
    double hx = pts[2 * a], hy = pts[2 * a + 1];
    double vhead = get_velocity_dev(vel, ov, np, ip, nx, ny, xmin, ymin, dx, dy, hx, hy);
    for (long long n = a + 1; n < b; ++n) {
      const double qx = pts[2 * n], qy = pts[2 * n + 1];
      const double ex = qx - hx, ey = qy - hy;
      double dist = ex * ex + ey * ey;
      dist = sqrt(dist);
      const double vtail = get_velocity_dev(vel, ov, np, ip, nx, ny, xmin, ymin, dx, dy, qx, qy);
      acc = acc + dist * 2.0 / (vhead + vtail);
      vhead = vtail;
      hx = qx; hy = qy;
    }
  }
  time[t] = acc;
}

// The same for rays held in fixed-size slots (the device ray tracer's output, k6_fm2d.cuh): ray t = points
// pts[t*cap .. t*cap + npts[t] - 1].
__global__ void __launch_bounds__(128) group_times_slots_kernel(const double* __restrict__ vel, const VelOverlay ov, int np, int nx, int ny,
                                                                double xmin, double ymin, double dx, double dy, const double* __restrict__ pts,
                                                                const int32_t* __restrict__ npts, int cap, int nrays, double* __restrict__ time) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)np * nrays) return;
  const int ip = (int)(t / nrays);
  const long long a = t * cap, b = a + npts[t];
  double acc = 0.0;
  if (b - a >= 2) {
    double hx = pts[2 * a], hy = pts[2 * a + 1];
    double vhead = get_velocity_dev(vel, ov, np, ip, nx, ny, xmin, ymin, dx, dy, hx, hy);
    for (long long n = a + 1; n < b; ++n) {
      const double qx = pts[2 * n], qy = pts[2 * n + 1];
      const double ex = qx - hx, ey = qy - hy;
      double dist = ex * ex + ey * ey;
      dist = sqrt(dist);
      const double vtail = get_velocity_dev(vel, ov, np, ip, nx, ny, xmin, ymin, dx, dy, qx, qy);
      acc = acc + dist * 2.0 / (vhead + vtail);
      vhead = vtail;
      hx = qx; hy = qy;
    }
  }
  time[t] = acc;
}

namespace {
// rays -> device (packed), validated
int upload_rays(const double* ray_points, const int64_t* ray_offsets, long long nt, DevBuf& d_pts, DevBuf& d_off, cudaStream_t st) {
  const long long npts = ray_offsets[nt];
  for (long long i = 0; i < nt; ++i)
    if (ray_offsets[i + 1] < ray_offsets[i] || ray_offsets[i] < 0) return fail(MCT_E_INVALID_ARG, "group_times: offsets must be non-decreasing");
  int rc;
  if ((rc = ensure(d_pts, (size_t)std::max<long long>(npts, 1) * 16)) || (rc = ensure(d_off, (size_t)(nt + 1) * 8))) return rc;
  if (npts > 0) CK(cudaMemcpyAsync(d_pts.p, ray_points, (size_t)npts * 16, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_off.p, ray_offsets, (size_t)(nt + 1) * 8, cudaMemcpyHostToDevice, st));
  return MCT_OK;
}
int launch_group_times(const double* d_vel, const VelOverlay& ov, int np, const mct_grid* gr, const double* d_pts, const long long* d_off,
                       int nrays, double* d_time, cudaStream_t st) {
  const long long nt = (long long)np * nrays;
  ProfScope ps(2, st);
  group_times_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(d_vel, ov, np, gr->nx, gr->ny, gr->xmin, gr->ymin, gr->dx, gr->dy, d_pts,
                                                                    d_off, nrays, d_time);
  CK(cudaGetLastError());
  g.host_stats.n_launches += 1;
  return MCT_OK;
}
} // namespace

extern "C" {

// d_vel: device (np,ny,nx) map (e.g. a session's or mct_surf_dispersion_dev's gvel output).  ray_points / ray_offsets:
// HOST, packed as ray r of period ip = points ray_offsets[ip*nrays + r] .. ray_offsets[ip*nrays + r + 1]-1 (x,y pairs).
// time: HOST out, (nrays, np) with the ray index fastest -- the reference's time(k,j,i) flattened over (k,j).
int mct_group_times_dev(const double* d_vel, int np, const mct_grid* gr, const double* ray_points, const int64_t* ray_offsets,
                        int nrays, double* time, void* stream) {
  NEED_INIT();
  if (!d_vel || !grid_ok(gr) || !ray_points || !ray_offsets || !time || np < 1 || nrays < 0)
    return fail(MCT_E_INVALID_ARG, "group_times: bad arguments");
  if (gr->nx < 2 || gr->ny < 2) return fail(MCT_E_INVALID_ARG, "group_times: the bilinear stencil needs nx, ny >= 2");
  const long long nt = (long long)np * nrays;
  if (nt == 0) return MCT_OK;
  cudaStream_t st = pick(stream);
  int rc;
  if ((rc = upload_rays(ray_points, ray_offsets, nt, g.ray_pts, g.ray_off, st)) || (rc = ensure(g.ray_time, (size_t)nt * 8))) return rc;
  const VelOverlay none = {nullptr, 0, 0, 0, 0};
  if ((rc = launch_group_times(d_vel, none, np, gr, (const double*)g.ray_pts.p, (const long long*)g.ray_off.p, nrays, (double*)g.ray_time.p, st)))
    return rc;
  CK(cudaMemcpyAsync(time, g.ray_time.p, (size_t)nt * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MCT_OK;
}

} // extern "C"
