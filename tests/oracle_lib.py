"""ctypes wrapper of oracle/liboracle.so -- the CPU restatement used as the parity checker.
TEST INFRASTRUCTURE ONLY: nothing under mctomo_b200/ may import this."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "kdtree2_ref")

LIBM, PORTABLE = 0, 1
REFERENCE = 2   # math_mode 2: the reference's own surfdisp96.f (oracle/_ref, see use_reference_solver)


class orc_grid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("xmin", C.c_double), ("ymin", C.c_double),
                ("zmin", C.c_double), ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("waterDepth", C.c_double), ("scaling", C.c_double)]


_L = None


def L():
    global _L
    if _L is None:
        p = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(p):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
        _L = C.CDLL(p)
        vp = C.c_void_p
        _L.orc_kdtree2_create.restype = vp
        _L.orc_kdtree2_create.argtypes = [vp, C.c_int]
        _L.orc_kdtree2_destroy.argtypes = [vp]
        _L.orc_kdtree2_nearest_batch.argtypes = [vp, vp, C.c_int64, vp, vp]
        _L.orc_kdtree_to_grid.argtypes = [vp, vp, C.c_int, C.POINTER(orc_grid), vp, vp, vp, vp, vp, vp]
        _L.orc_box_window.argtypes = [C.POINTER(orc_grid), vp, vp]
        _L.orc_surfmodes.argtypes = [vp] * 4 + [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                                vp, vp, vp, vp]
        _L.orc_vs2vp_rho.argtypes = [vp, vp, vp, C.c_int64, C.c_int]
        _L.orc_check_model.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        _L.orc_convert_column.argtypes = [vp, vp, vp, C.c_int] + [C.c_double] * 5 + [vp] * 4
        _L.orc_surf_dispersion.argtypes = [vp, vp, vp, C.POINTER(orc_grid)] + [C.c_int] * 4 + [vp, C.c_int, C.c_int,
                                          C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_int, C.c_int, vp, vp, vp, vp]
        _L.orc_assemble_vel.argtypes = [vp] + [C.c_int] * 7 + [vp]
        _L.orc_mct_exp.restype = C.c_double
        _L.orc_mct_exp.argtypes = [C.c_double]
        _L.orc_mct_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _L.orc_mct_pow025.restype = C.c_double
        _L.orc_mct_pow025.argtypes = [C.c_double]
        _L.orc_gtsolh.restype = C.c_float
        _L.orc_gtsolh.argtypes = [C.c_float, C.c_float]
        _L.orc_nlvls1.argtypes = [vp, vp, C.c_int, C.c_int]
        _L.orc_cal_group_time.argtypes = [vp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + [vp, vp, C.c_int, vp]
    return _L


def ogrid(g):
    return orc_grid(g.nx, g.ny, g.nz, g.xmin, g.ymin, g.zmin, g.dx, g.dy, g.dz, g.waterDepth, g.scaling)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def kd_nearest(points, queries):
    """Oracle kdtree2 1-NN: returns (idx 1-based int32, squared distance).  2-D points/queries are embedded with z = 0:
    kdtree2's build never cuts a zero-extent dimension and a zero third term changes no sum (pinned against kdtree2.o
    run with dim = 2: tests/test_oracle_kdtree.py, fixture grid2d)."""
    points, queries = f64(points), f64(queries)
    if points.shape[1] == 2:
        points = np.ascontiguousarray(np.column_stack([points, np.zeros(len(points))]))
        queries = np.ascontiguousarray(np.column_stack([queries, np.zeros(len(queries))]))
    T = L().orc_kdtree2_create(points.ctypes.data, len(points))
    if not T:
        raise ValueError("degenerate nuclei")
    idx = np.zeros(len(queries), np.int32)
    dis = np.zeros(len(queries))
    L().orc_kdtree2_nearest_batch(T, queries.ctypes.data, len(queries), idx.ctypes.data, dis.ctypes.data)
    L().orc_kdtree2_destroy(T)
    return idx, dis


def have_ref_binary():
    return os.path.exists(REF_BIN)


def ref_kd_nearest(points, queries):
    """The reference's own kdtree2.o (oracle/_ref/kdtree2_ref, built by oracle/build_ref.sh); 3-D or 2-D points."""
    points, queries = f64(points), f64(queries)
    dim = points.shape[1]
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in.bin"), "wb") as f:
            f.write(np.array([len(points), len(queries)], np.int64).tobytes())
            f.write(points.tobytes())
            f.write(queries.tobytes())
        subprocess.check_call([REF_BIN, os.path.join(d, "in.bin"), os.path.join(d, "out.bin"), str(dim)])
        b = open(os.path.join(d, "out.bin"), "rb").read()
    n = len(queries)
    return np.frombuffer(b[:4 * n], np.int32).copy(), np.frombuffer(b[4 * n:], np.float64).copy()


def kdtree_to_grid(points, params, grid, box, vp, vs, rho, sites_id, pm=None):
    points, params, box = f64(points), f64(params), f64(box)
    pmv = None if pm is None else f64(pm)
    og = ogrid(grid)
    rc = L().orc_kdtree_to_grid(points.ctypes.data, params.ctypes.data, len(points), C.byref(og), box.ctypes.data,
                                None if pmv is None else pmv.ctypes.data, vp.ctypes.data, vs.ctypes.data,
                                rho.ctypes.data, sites_id.ctypes.data)
    if rc:
        raise ValueError("degenerate nuclei")


def box_window(grid, box):
    w = np.zeros(6, np.int32)
    box = f64(box)
    og = ogrid(grid)
    L().orc_box_window(C.byref(og), box.ctypes.data, w.ctypes.data)
    return w


def surfmodes(thick, vp, vs, rho, freqs, modetype=1, phaseGroup=0, nmodes=0, dc=1e-3, math_mode=PORTABLE):
    thick, vp, vs, rho, freqs = f64(thick), f64(vp), f64(vs), f64(rho), f64(freqs)
    nm = max(nmodes, 1)
    ph = np.zeros(len(freqs) * nm)
    gr = np.zeros(len(freqs) * nm)
    ierr = C.c_int(0)
    cnt = np.zeros(2, np.int64)
    rc = L().orc_surfmodes(thick.ctypes.data, vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, len(thick),
                           freqs.ctypes.data, len(freqs), modetype, phaseGroup, nmodes, dc, math_mode, ph.ctypes.data,
                           gr.ctypes.data, C.byref(ierr), cnt.ctypes.data)
    return rc, ph, gr, ierr.value, cnt


def vs2vp_rho(vs, math_mode=PORTABLE):
    vs = f64(vs)
    vp = np.empty_like(vs)
    rho = np.empty_like(vs)
    L().orc_vs2vp_rho(vs.ctypes.data, vp.ctypes.data, rho.ctypes.data, vs.size, math_mode)
    return vp, rho


def check_model(vs, grid):
    vs = f64(vs)
    return L().orc_check_model(vs.ctypes.data, grid.nx, grid.ny, grid.nz)


def surf_dispersion(vp, vs, rho, grid, window, freqs, raylov=1, phaseGroup=0, nmodes=0, dphase=1e-3,
                    layer_eps=float(np.float32(1e-10)), water_thresh=float(np.float32(1e-10)), preset=100.0,
                    math_mode=PORTABLE, nthreads=None):
    vp, vs, rho, freqs = f64(vp), f64(vs), f64(rho), f64(freqs)
    ix0, ix1, iy0, iy1 = (int(v) for v in window)
    wx, wy = ix1 - ix0 + 1, iy1 - iy0 + 1
    nm = max(nmodes, 1)
    pvel = np.zeros((wx, wy, nm * len(freqs)))
    gvel = np.zeros_like(pvel)
    ierr = np.zeros((wx, wy), np.int32)
    cnt = np.zeros(2, np.int64)
    og = ogrid(grid)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    nun = L().orc_surf_dispersion(vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, C.byref(og), ix0, ix1, iy0, iy1,
                                  freqs.ctypes.data, len(freqs), raylov, phaseGroup, nmodes, dphase, layer_eps,
                                  water_thresh, preset, math_mode, nthreads, pvel.ctypes.data, gvel.ctypes.data,
                                  ierr.ctypes.data, cnt.ctypes.data)
    return pvel, gvel, ierr, cnt, nun


def convert_column(vp, vs, rho, dz, waterDepth=0.0, scaling=1.0, layer_eps=float(np.float32(1e-10)),
                   water_thresh=float(np.float32(1e-10))):
    vp, vs, rho = f64(vp), f64(vs), f64(rho)
    out = [np.zeros(256) for _ in range(4)]
    n = L().orc_convert_column(vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, len(vs), dz, waterDepth, scaling,
                               layer_eps, water_thresh, *[o.ctypes.data for o in out])
    return n, [o[:max(n, 0)] for o in out]  # thick, alpha, beta, rho


def assemble_vel(pvel, np_, nx, ny, window, vel):
    ix0, ix1, iy0, iy1 = (int(v) for v in window)
    pvel = f64(pvel)
    L().orc_assemble_vel(pvel.ctypes.data, np_, nx, ny, ix0, ix1, iy0, iy1, vel.ctypes.data)


def forward_eval(points, params, grid, freqs, derive_vp_rho=True, math_mode=PORTABLE, nthreads=None, **kw):
    """Oracle version of the fused forward evaluation: kdtree_to_grid(full box) -> vs2vp/rho -> check -> dispersion."""
    vp = np.zeros(grid.shape)
    vs = np.zeros(grid.shape)
    rho = np.zeros(grid.shape)
    sid = np.zeros(grid.shape, np.int32)
    kdtree_to_grid(points, params, grid, grid.cover_box(), vp, vs, rho, sid)  # every node, like mct_forward_eval
    if derive_vp_rho:
        vp, rho = vs2vp_rho(vs, math_mode)
    inval = check_model(vs, grid)
    res = dict(vp=vp, vs=vs, rho=rho, sites_id=sid, model_invalid=inval)
    if not inval:
        pv, gv, ie, cnt, nun = surf_dispersion(vp, vs, rho, grid, (1, grid.nx, 1, grid.ny), freqs,
                                               math_mode=math_mode, nthreads=nthreads, **kw)
        res.update(pvel=pv, gvel=gv, ierr=ie, counters=cnt)
    return res


def cal_group_time(vel, grid, ray_points, ray_offsets, nrays):
    """CalGroupTime (likelihood_surf.F90:454-494): vel is the (nx, ny, np) C-ordered view of the (np,ny,nx) map."""
    vel, pts = f64(vel), f64(ray_points)
    off = np.ascontiguousarray(ray_offsets, dtype=np.int64)
    np_ = vel.shape[2]
    t = np.zeros((np_, nrays))
    L().orc_cal_group_time(vel.ctypes.data, np_, grid.nx, grid.ny, grid.xmin, grid.ymin, grid.dx, grid.dy, pts.ctypes.data,
                           off.ctypes.data, nrays, t.ctypes.data)
    return t


def surf_misfit(time, ttime, raystat, sigdep=0, nrays_total=None, snoise0=None, snoise1=None, srdist=None, math_mode=PORTABLE):
    """The Gaussian misfit of likelihood_surf.F90:356-404.  time (np,nrr), ttime (np,3,nrr), raystat (np,2,nrr),
    srdist (np,nrr) in C order (== the Fortran (nrr,..,np) arrays)."""
    t, tt = f64(time), f64(ttime)
    rs = np.ascontiguousarray(raystat, dtype=np.int32)
    np_, nrr = t.shape
    if nrays_total is None:
        nrays_total = int((rs[:, 0, :] == 1).sum())
    out, sg = np.zeros(3), np.zeros((np_, nrr))
    n0 = None if snoise0 is None else f64(snoise0)
    n1 = None if snoise1 is None else f64(snoise1)
    sd = None if srdist is None else f64(srdist)
    fn = L().orc_surf_misfit
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
    p = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    rc = fn(t.ctypes.data, nrr, np_, sigdep, nrays_total, tt.ctypes.data, rs.ctypes.data, p(n0), p(n1), p(sd), math_mode,
            out.ctypes.data, sg.ctypes.data)
    return dict(like=out[0], misfit=out[1], unweighted_misfit=out[2], sigma=sg, rc=rc)


# ---- the reference's own surfdisp96.f, mechanically translated to C (oracle/f77toc.py, oracle/build_ref.sh) ---------
F2C_LIB = os.path.join(ORACLE_DIR, "_ref", "libsurfdisp96_f2c.so")
_F2C = None


def have_f2c():
    return os.path.exists(F2C_LIB)


def f2c_surfdisp(thick, vp, vs, rho, freqs, iwave, igr, nmodes=0, dphase=1e-3):
    """Calls the translated surfdisp96 (nmodes <= 0) / surfdisp_mmodes exactly as surfmodes.f90:81-83,93-95 / :155-180
    call the Fortran: the model narrowed to real*4, periods dble(1/freqs), iflsph = 0.  Returns (cp, cg, ierr) with
    cp, cg of shape (nm*np,), mode-major like surfmodes.f90:179-180."""
    global _F2C
    if _F2C is None:
        _F2C = C.CDLL(F2C_LIB)
    n = len(thick)
    a = [np.zeros(200, np.float32) for _ in range(4)]      # real*4 thkm(NLAY) ...: NLAY = 200
    for dst, src in zip(a, (thick, vp, vs, rho)):
        dst[:n] = np.asarray(src, np.float64).astype(np.float32)
    np_ = len(freqs)
    t = np.zeros(60)                                       # double precision t(NP), NP = 60
    t[:np_] = 1.0 / np.asarray(freqs, np.float64)
    nm = max(nmodes, 1)
    cp, cg = np.zeros(np_ * nm), np.zeros(np_ * nm)
    ci = lambda v: C.byref(C.c_int(v))  # noqa: E731
    ierr = C.c_int(0)
    dph = C.c_double(dphase)
    fn = _F2C.surfdisp96_ if nmodes <= 0 else _F2C.surfdisp_mmodes_
    fn(a[0].ctypes.data_as(C.c_void_p), a[1].ctypes.data_as(C.c_void_p), a[2].ctypes.data_as(C.c_void_p),
       a[3].ctypes.data_as(C.c_void_p), ci(n), ci(0), ci(iwave), ci(nm), ci(igr), ci(np_), t.ctypes.data_as(C.c_void_p),
       C.byref(dph), cp.ctypes.data_as(C.c_void_p), cg.ctypes.data_as(C.c_void_p), C.byref(ierr))
    return cp, cg, ierr.value


def use_reference_solver():
    """Hands the translated reference's surfdisp96_ / surfdisp_mmodes_ to liboracle: math_mode=REFERENCE then solves every
    column with the reference's own code (layering, kd-tree and loops stay the port's).  False when oracle/_ref is absent."""
    global _F2C
    if not have_f2c():
        return False
    if _F2C is None:
        _F2C = C.CDLL(F2C_LIB)
    fn = L().orc_set_reference_solver
    fn.argtypes = [C.c_void_p, C.c_void_p]
    fn(C.cast(_F2C.surfdisp96_, C.c_void_p), C.cast(_F2C.surfdisp_mmodes_, C.c_void_p))
    return True


def sites_locate(points, sites_id, grid, queries):
    """sites_locate (likelihood_body.F90:799-831) for a batch of 3-D points."""
    points, queries = f64(points), f64(queries)
    sid = np.ascontiguousarray(sites_id, dtype=np.int32)
    out = np.zeros(len(queries), np.int32)
    fn = L().orc_sites_locate
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(orc_grid), C.c_void_p, C.c_int64, C.c_void_p]
    og = ogrid(grid)
    if fn(points.ctypes.data, len(points), sid.ctypes.data, C.byref(og), queries.ctypes.data, len(queries), out.ctypes.data):
        raise ValueError("degenerate nuclei")
    return out


# GRT parameter sets: (tolmin, tolmax, smin_min, smin_max, dcm, dc2) as the two reference callers set T_MODES_PARA
GRT_PAR_LIKELIHOOD = (1e-6, 1e-5, float(np.float32(1e-3)), float(np.float32(5e-3)), 1e-3, 1e-3)  # likelihood_surf.F90:173-182 with tol = 1e-6
GRT_PAR_MODELLING = (float(np.float32(1e-6)), float(np.float32(1e-7)), float(np.float32(1e-3)), float(np.float32(5e-3)),
                     float(np.float32(1e-3)), float(np.float32(1e-3)))  # forward_modelling.f90:395-404


def grt_modes(thick, vp, vs, rho, freqs, modetype=1, phaseGroup=0, dc=1e-3, par=GRT_PAR_LIKELIHOOD, math_mode=PORTABLE, preset=100.0):
    """The generalized R/T branch of surfmodes (oracle/grt_ref.c).  Returns (ierr, phase, group, counters)."""
    lib = L()
    vpt = C.c_void_p
    lib.orc_grt_modes.argtypes = [vpt] * 4 + [C.c_int, vpt, C.c_int, C.c_int, C.c_int, C.c_double, vpt, C.c_int, vpt, vpt, vpt]
    thick, vp, vs, rho, freqs = f64(thick), f64(vp), f64(vs), f64(rho), f64(freqs)
    par = f64(np.array(par))
    ph = np.full(len(freqs), preset)
    gr = np.full(len(freqs), preset)
    cnt = np.zeros(2, np.int64)
    ierr = lib.orc_grt_modes(thick.ctypes.data, vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, len(thick), freqs.ctypes.data,
                             len(freqs), modetype, phaseGroup, dc, par.ctypes.data, math_mode, ph.ctypes.data, gr.ctypes.data,
                             cnt.ctypes.data)
    return ierr, ph, gr, cnt


def grt_secfun(thick, vp, vs, rho, freq, modetype, c, math_mode=PORTABLE):
    lib = L()
    vpt = C.c_void_p
    lib.orc_grt_secfun.argtypes = [vpt] * 4 + [C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, vpt, vpt]
    thick, vp, vs, rho = f64(thick), f64(vp), f64(vs), f64(rho)
    re, im = C.c_double(0), C.c_double(0)
    rc = lib.orc_grt_secfun(thick.ctypes.data, vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, len(thick), freq, modetype, c,
                            math_mode, C.byref(re), C.byref(im))
    return rc, re.value, im.value


def fm2d_unreached(nsrc):
    """Arms the oracle's per-source count of nodes the march leaves unreached (0 for a skipped source); returns the int32 array the
    next fm2d_times / fm2d_rays call ON THIS THREAD fills.  A non-zero entry marks a source whose field is, in the reference, the
    previous source's (tests exclude those sources from bit comparisons and check that the device flags them)."""
    buf = np.zeros(nsrc, np.int32)
    L().orc_fm2d_set_unreached.argtypes = [C.c_void_p]
    L().orc_fm2d_set_unreached(buf.ctypes.data)
    return buf


def fm2d_disarm():
    L().orc_fm2d_set_unreached.argtypes = [C.c_void_p]
    L().orc_fm2d_set_unreached(None)


def fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, gdx=1, gdz=1, asgr=1, sgdl=4, sgs=8, fom=1, snb=0.5, want_field=False):
    """modrays for one velocity map, travel times only (oracle/fm2d_ref.c).  src (nsrc,2), rcv (nrc,2) as (x, z);
    srs (nsrc, nrc) 0/1; vel (nvx+2, nvz+2) C-order = the Fortran like%vel(period,:,:) (nvz+2, nvx+2) with its edge.
    Returns (err, ttime (nsrc, nrc), field (nsrc, nnx, nnz) or None, counters)."""
    lib = L()
    vpt = C.c_void_p
    lib.orc_fm2d_times.argtypes = [C.c_int, vpt, vpt, C.c_int, vpt, vpt, vpt, C.c_int, C.c_int] + [C.c_double] * 4 + [vpt] + [C.c_int] * 6 + \
                                  [C.c_double, vpt, vpt, vpt]
    src, rcv, vel = f64(src), f64(rcv), f64(vel)
    nsrc, nrc = len(src), len(rcv)
    nvx, nvz = vel.shape[0] - 2, vel.shape[1] - 2
    scx, scz = f64(src[:, 0].copy()), f64(src[:, 1].copy())
    rcx, rcz = f64(rcv[:, 0].copy()), f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(srs, dtype=np.int32)
    tt = np.full((nsrc, nrc), -1.0)
    nnx, nnz = (nvx - 1) * gdx + 1, (nvz - 1) * gdz + 1
    field = np.zeros((nsrc, nnx, nnz)) if want_field else None
    cnt = np.zeros(2, np.int64)
    err = lib.orc_fm2d_times(nsrc, scx.ctypes.data, scz.ctypes.data, nrc, rcx.ctypes.data, rcz.ctypes.data, srs.ctypes.data, nvx, nvz,
                             gox, goz, dvx, dvz, vel.ctypes.data, gdx, gdz, asgr, sgdl, sgs, fom, snb, tt.ctypes.data,
                             field.ctypes.data if want_field else None, cnt.ctypes.data)
    return err, tt, field, cnt


def fm2d_rays(src, rcv, srs, vel, gox, goz, dvx, dvz, gdx=1, gdz=1, asgr=1, sgdl=4, sgs=8, fom=1, snb=0.5, cap=None, srsv=None):
    """modrays with uar = 0 (group-velocity data): travel times AND ray geometry.  Returns (err, ttime (nsrc,nrc), npts (nsrc*nrc),
    pts (nsrc*nrc, cap, 2), length (nsrc*nrc), crazy); slot = srsv - 1 (default: the pair's own index, receiver fastest)."""
    lib = L()
    vpt = C.c_void_p
    lib.orc_fm2d_rays.argtypes = [C.c_int, vpt, vpt, C.c_int, vpt, vpt, vpt, vpt, C.c_int, C.c_int] + [C.c_double] * 4 + [vpt] + [C.c_int] * 6 + \
                                 [C.c_double, vpt, C.c_int, vpt, vpt, vpt, vpt]
    src, rcv, vel = f64(src), f64(rcv), f64(vel)
    nsrc, nrc = len(src), len(rcv)
    nvx, nvz = vel.shape[0] - 2, vel.shape[1] - 2
    scx, scz = f64(src[:, 0].copy()), f64(src[:, 1].copy())
    rcx, rcz = f64(rcv[:, 0].copy()), f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(srs, dtype=np.int32)
    if srsv is None:
        srsv = np.arange(1, nsrc * nrc + 1, dtype=np.int32).reshape(nsrc, nrc)
    srsv = np.ascontiguousarray(srsv, dtype=np.int32)
    if cap is None:
        cap = 8 * (nvx * gdx + nvz * gdz)
    tt = np.full((nsrc, nrc), -1.0)
    npts = np.zeros(nsrc * nrc, np.int32)
    pts = np.zeros((nsrc * nrc, cap, 2))
    ln = np.zeros(nsrc * nrc)
    crazy = C.c_int(0)
    err = lib.orc_fm2d_rays(nsrc, scx.ctypes.data, scz.ctypes.data, nrc, rcx.ctypes.data, rcz.ctypes.data, srs.ctypes.data, srsv.ctypes.data,
                            nvx, nvz, gox, goz, dvx, dvz, vel.ctypes.data, gdx, gdz, asgr, sgdl, sgs, fom, snb, tt.ctypes.data, cap,
                            npts.ctypes.data, pts.ctypes.data, ln.ctypes.data, C.byref(crazy))
    return err, tt, npts, pts, ln, crazy.value


# ---- the reference's own fm2d/fm2d_ttime.f90, mechanically translated to C (oracle/f90toc.py, oracle/build_ref.sh) ----------
FM2D_F2C_LIB = os.path.join(ORACLE_DIR, "_ref", "libfm2d_ttime_f2c.so")
_fm2d_f2c = None


def have_fm2d_reference():
    return os.path.exists(FM2D_F2C_LIB)


def fm2d_travel(impl, veln, gox, goz, dnx, dnz, fom, scx, scz, urg=0, window=None, ttn=None, nsts=None):
    """One call of `travel` (fm2d_ttime.f90:27-136) on a propagation grid veln (nnx, nnz) C-order = the Fortran's (nnz, nnx).
    impl "port": oracle/fm2d_ref.c; "reference": the translated Fortran.  window = (vnl, vnr, vnt, vnb) for urg = 1;
    ttn / nsts carry the state from call to call (urg = 2 continues).  Returns (stopped/err, ttn, nsts, heap (ntr, 2) as (px, pz))."""
    global _fm2d_f2c
    if impl == "reference":
        if _fm2d_f2c is None:
            _fm2d_f2c = C.CDLL(FM2D_F2C_LIB)
        fn = _fm2d_f2c.ref_fm2d_travel
    else:
        fn = L().orc_fm2d_travel
    vpt = C.c_void_p
    fn.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_int, vpt, vpt, vpt] + [C.c_int] * 5 + [C.c_double] * 2 + [vpt, vpt]
    fn.restype = C.c_int
    veln = f64(veln)
    nnx, nnz = veln.shape
    ttn = np.zeros((nnx, nnz)) if ttn is None else f64(ttn).copy()
    nsts = np.full((nnx, nnz), -1, np.int32) if nsts is None else np.ascontiguousarray(nsts, dtype=np.int32).copy()
    heap = np.zeros((nnx * nnz + 2, 2), np.int32)
    ntr = C.c_int(0)
    vnl, vnr, vnt, vnb = window if window is not None else (1, nnx, 1, nnz)
    rc = fn(nnx, nnz, gox, goz, dnx, dnz, fom, veln.ctypes.data, ttn.ctypes.data, nsts.ctypes.data, urg, vnl, vnr, vnt, vnb, scx, scz,
            heap.ctypes.data, C.byref(ntr))
    return rc, ttn, nsts, heap[:ntr.value].copy()


def _fm2d_fn(impl, name):
    global _fm2d_f2c
    if impl == "reference":
        if _fm2d_f2c is None:
            _fm2d_f2c = C.CDLL(FM2D_F2C_LIB)
        return getattr(_fm2d_f2c, "ref_fm2d_" + name)
    return getattr(L(), "orc_fm2d_" + name)


def fm2d_gridder(impl, velv, gdx, gdz):
    """gridder (fm2dray_cartesian.f90:490-590): velv (nvx+2, nvz+2) C-order -> veln (nnx, nnz)"""
    fn = _fm2d_fn(impl, "gridder")
    velv = f64(velv)
    nvx, nvz = velv.shape[0] - 2, velv.shape[1] - 2
    out = np.zeros(((nvx - 1) * gdx + 1, (nvz - 1) * gdz + 1))
    fn.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2
    fn(nvx, nvz, gdx, gdz, velv.ctypes.data, out.ctypes.data)
    return out


def fm2d_bsplrefine(impl, velv, gdx, gdz, sgdl, window):
    """bsplrefine (fm2dray_cartesian.f90:598-668) on the window (vnl, vnr, vnt, vnb) of the propagation grid"""
    fn = _fm2d_fn(impl, "bsplrefine")
    velv = f64(velv)
    nvx, nvz = velv.shape[0] - 2, velv.shape[1] - 2
    vnl, vnr, vnt, vnb = window
    nnxr, nnzr = (vnr - vnl) * sgdl + 1, (vnb - vnt) * sgdl + 1
    out = np.zeros((nnxr, nnzr))
    fn.argtypes = [C.c_int] * 9 + [C.c_void_p] + [C.c_int] * 2 + [C.c_void_p]
    fn(nvx, nvz, gdx, gdz, sgdl, vnl, vnr, vnt, vnb, velv.ctypes.data, nnxr, nnzr, out.ctypes.data)
    return out


def fm2d_srtimes(impl, veln, ttn, gox, goz, dnx, dnz, scx, scz, rcv, srs):
    """srtimes (fm2dray_cartesian.f90:676-770) for one source: rcv (nrc, 2), srs (nrc) -> (stopped/err, ttime (nrc), -1 where untouched)"""
    fn = _fm2d_fn(impl, "srtimes")
    veln, ttn, rcv = f64(veln), f64(ttn), f64(rcv)
    nnx, nnz = veln.shape
    rcx, rcz = f64(rcv[:, 0].copy()), f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(srs, dtype=np.int32)
    tt = np.full(len(rcv), -1.0)
    vpt = C.c_void_p
    fn.argtypes = [C.c_int] * 2 + [C.c_double] * 4 + [vpt, vpt] + [C.c_double] * 2 + [C.c_int, vpt, vpt, vpt, vpt]
    fn.restype = C.c_int
    rc = fn(nnx, nnz, gox, goz, dnx, dnz, veln.ctypes.data, ttn.ctypes.data, scx, scz, len(rcv), rcx.ctypes.data, rcz.ctypes.data,
            srs.ctypes.data, tt.ctypes.data)
    return rc, tt


def fm2d_times_reference(src, rcv, srs, vel, gox, goz, dvx, dvz, gdx=1, gdz=1, asgr=1, sgdl=4, sgs=8, fom=1, snb=0.5):
    """modrays for travel times (uar = 1) through the TRANSLATED reference: gridder, the per-source body of modrays' loop,
    travel, bsplrefine and srtimes are the Fortran's own statements (oracle/f90toc.py); the allocations and the loop skeleton
    around them are oracle/ref_harness/fm2d_f90_harness.c.  Same arguments and results as fm2d_times(want_field=True)."""
    fn = _fm2d_fn("reference", "modrays_times")
    vpt = C.c_void_p
    fn.argtypes = [C.c_int, vpt, vpt, C.c_int, vpt, vpt, vpt, C.c_int, C.c_int] + [C.c_double] * 4 + [vpt] + [C.c_int] * 6 + [C.c_double, vpt, vpt]
    fn.restype = C.c_int
    src, rcv, vel = f64(src), f64(rcv), f64(vel)
    nsrc, nrc = len(src), len(rcv)
    nvx, nvz = vel.shape[0] - 2, vel.shape[1] - 2
    scx, scz = f64(src[:, 0].copy()), f64(src[:, 1].copy())
    rcx, rcz = f64(rcv[:, 0].copy()), f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(srs, dtype=np.int32)
    tt = np.full((nsrc, nrc), -1.0)
    nnx, nnz = (nvx - 1) * gdx + 1, (nvz - 1) * gdz + 1
    field = np.zeros((nsrc, nnx, nnz))
    err = fn(nsrc, scx.ctypes.data, scz.ctypes.data, nrc, rcx.ctypes.data, rcz.ctypes.data, srs.ctypes.data, nvx, nvz, gox, goz, dvx, dvz,
             vel.ctypes.data, gdx, gdz, asgr, sgdl, sgs, fom, snb, tt.ctypes.data, field.ctypes.data)
    return err, tt, field


def fm2d_rays_reference(src, rcv, srs, vel, gox, goz, dvx, dvz, gdx=1, gdz=1, asgr=1, sgdl=4, sgs=8, fom=1, snb=0.5, cap=None, srsv=None):
    """modrays with uar = 0 through the TRANSLATED reference (rpaths included): same arguments as fm2d_rays; returns
    (err, ttime, npts, pts, crazy)."""
    fn = _fm2d_fn("reference", "modrays")
    vpt = C.c_void_p
    fn.argtypes = [C.c_int, vpt, vpt, C.c_int, vpt, vpt, vpt, C.c_int, C.c_int] + [C.c_double] * 4 + [vpt] + [C.c_int] * 6 + \
                  [C.c_double, vpt, vpt, C.c_int, vpt, C.c_int, vpt, vpt, vpt]
    fn.restype = C.c_int
    src, rcv, vel = f64(src), f64(rcv), f64(vel)
    nsrc, nrc = len(src), len(rcv)
    nvx, nvz = vel.shape[0] - 2, vel.shape[1] - 2
    scx, scz = f64(src[:, 0].copy()), f64(src[:, 1].copy())
    rcx, rcz = f64(rcv[:, 0].copy()), f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(srs, dtype=np.int32)
    if srsv is None:
        srsv = np.arange(1, nsrc * nrc + 1, dtype=np.int32).reshape(nsrc, nrc)
    srsv = np.ascontiguousarray(srsv, dtype=np.int32)
    if cap is None:
        cap = 8 * (nvx * gdx + nvz * gdz)
    tt = np.full((nsrc, nrc), -1.0)
    npts = np.zeros(nsrc * nrc, np.int32)
    pts = np.zeros((nsrc * nrc, cap, 2))
    crazy = C.c_int(0)
    err = fn(nsrc, scx.ctypes.data, scz.ctypes.data, nrc, rcx.ctypes.data, rcz.ctypes.data, srs.ctypes.data, nvx, nvz, gox, goz, dvx, dvz,
             vel.ctypes.data, gdx, gdz, asgr, sgdl, sgs, fom, snb, tt.ctypes.data, None, 0, srsv.ctypes.data, cap, npts.ctypes.data,
             pts.ctypes.data, C.byref(crazy))
    return err, tt, npts, pts, crazy.value


# ---- the reference's own surfmodes/Love.f90, mechanically translated to C (oracle/f90toc_love.py, oracle/build_ref.sh) ------
LOVE_F2C_LIB = os.path.join(ORACLE_DIR, "_ref", "liblove_f2c.so")
_love_f2c = None


def have_love_reference():
    return os.path.exists(LOVE_F2C_LIB)


def _grt_state(thick, vp, vs, rho, freq, modetype, c):
    vpt = C.c_void_p
    st = L().orc_grt_state
    st.argtypes = [vpt] * 4 + [C.c_int, C.c_double, C.c_int, C.c_double] + [vpt] * 6
    thick, vp, vs, rho = f64(thick), f64(vp), f64(vs), f64(rho)
    n = len(thick)
    d, p, v, mu = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    ints = np.zeros(3, np.int32)
    w = C.c_double(0)
    rc = st(thick.ctypes.data, vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, n, freq, modetype, c, d.ctypes.data, p.ctypes.data,
            v.ctypes.data, mu.ctypes.data, ints.ctypes.data, C.byref(w))
    assert rc == 0
    return n, d, p, v, mu, ints, w.value


def grt_love_secfun_reference(thick, vp, vs, rho, freq, c):
    """SecFuns_L(1 + ifs, c, GRT, Imf) of the TRANSLATED Love.f90 on the T_GRT that the restatement's setup_grt + startl build
    for this column (the search logic and the set-up stay the restatement's).  Returns (value, imf)."""
    global _love_f2c
    if _love_f2c is None:
        _love_f2c = C.CDLL(LOVE_F2C_LIB)
    vpt = C.c_void_p
    n, d, p, v, mu, ints, w = _grt_state(thick, vp, vs, rho, freq, 0, c)
    fn = _love_f2c.ref_love_secfun
    fn.argtypes = [C.c_int, vpt, vpt, vpt, C.c_int, C.c_int, C.c_double, C.c_double, vpt, vpt]
    val, imf = C.c_double(0), C.c_double(0)
    fn(n, d.ctypes.data, v.ctypes.data, mu.ctypes.data, int(ints[0]), int(ints[1]), w, c, C.byref(val), C.byref(imf))
    return val.value, imf.value


RAYLEIGH_F2C_LIB = os.path.join(ORACLE_DIR, "_ref", "librayleigh_f2c.so")
_rayleigh_f2c = None


def have_rayleigh_reference():
    return os.path.exists(RAYLEIGH_F2C_LIB)


def grt_rayleigh_secfun_reference(thick, vp, vs, rho, freq, c):
    """startl + SecFunSurf(0, c, GRT, Imf) of the TRANSLATED Rayleigh.f90 (columns without water: ifs = 0) on the T_GRT that the
    restatement's setup_grt builds.  Returns (value, imf, ll chosen by the translated startl, ll of the restatement)."""
    global _rayleigh_f2c
    if _rayleigh_f2c is None:
        _rayleigh_f2c = C.CDLL(RAYLEIGH_F2C_LIB)
    vpt = C.c_void_p
    n, d, p, v, mu, ints, w = _grt_state(thick, vp, vs, rho, freq, 1, c)
    assert ints[0] == 0, "a water layer: SecFunSt, not SecFunSurf"
    fn = _rayleigh_f2c.ref_rayleigh_secfunsurf
    fn.argtypes = [C.c_int, vpt, vpt, vpt, vpt, C.c_int, C.c_double, C.c_double, vpt, vpt, vpt]
    val, imf, ll = C.c_double(0), C.c_double(0), C.c_int(0)
    fn(n, d.ctypes.data, p.ctypes.data, v.ctypes.data, mu.ctypes.data, int(ints[2]), w, c, C.byref(val), C.byref(imf), C.byref(ll))
    return val.value, imf.value, ll.value, int(ints[1])


def grt_bisecim(impl, thick, vp, vs, rho, freq, modetype, k1, k2, smin=1e-4, tol=1e-6):
    """One root refinement as the searches issue it (startl at k2, the secular function at both ends, bisecim).  impl "port": the
    restatement; "reference": bisecim of util.f90 driving SecFunSurf / SecFuns_L, all translated.  Returns (iq, root, f1, f2)."""
    vpt = C.c_void_p
    out = np.zeros(3)
    if impl == "port":
        fn = L().orc_grt_bisecim
        fn.argtypes = [vpt] * 4 + [C.c_int, C.c_double, C.c_int] + [C.c_double] * 4 + [vpt]
        a = [f64(x) for x in (thick, vp, vs, rho)]
        iq = fn(*[x.ctypes.data for x in a], len(a[0]), freq, modetype, k1, k2, smin, tol, out.ctypes.data)
        return iq, out[0], out[1], out[2]
    n, d, p, v, mu, ints, w = _grt_state(thick, vp, vs, rho, freq, modetype, k2)
    global _love_f2c, _rayleigh_f2c
    if modetype == 0:
        if _love_f2c is None:
            _love_f2c = C.CDLL(LOVE_F2C_LIB)
        fn = _love_f2c.ref_love_bisecim
        fn.argtypes = [C.c_int, vpt, vpt, vpt, C.c_int, C.c_int] + [C.c_double] * 5 + [vpt]
        iq = fn(n, d.ctypes.data, v.ctypes.data, mu.ctypes.data, int(ints[0]), int(ints[1]), w, k1, k2, smin, tol, out.ctypes.data)
    else:
        assert ints[0] == 0
        if _rayleigh_f2c is None:
            _rayleigh_f2c = C.CDLL(RAYLEIGH_F2C_LIB)
        fn = _rayleigh_f2c.ref_rayleigh_bisecim
        fn.argtypes = [C.c_int, vpt, vpt, vpt, vpt, C.c_int] + [C.c_double] * 5 + [vpt]
        iq = fn(n, d.ctypes.data, p.ctypes.data, v.ctypes.data, mu.ctypes.data, int(ints[2]), w, k1, k2, smin, tol, out.ctypes.data)
    return iq, out[0], out[1], out[2]


def grt_cinterval(impl, thick, vp, vs, rho, freq, modetype, tol=1e-5):
    """The trial phase velocities of one frequency (C_Interval / C_Interval_L).  impl "port": the restatement; "reference": the
    translated Fortran on the T_GRT the restatement's setup_grt builds.  Returns (ccc[:ncc], im1, overflow flag of the port)."""
    vpt = C.c_void_p
    a = [f64(x) for x in (thick, vp, vs, rho)]
    n = len(a[0])
    ccc, v, extra, counts = np.zeros(20008), np.zeros(2 * n + 8), np.zeros(3), np.zeros(4, np.int32)
    fn = L().orc_grt_cinterval
    fn.argtypes = [vpt] * 4 + [C.c_int, C.c_double, C.c_int, C.c_double] + [vpt] * 4
    rc = fn(*[x.ctypes.data for x in a], n, freq, modetype, tol, ccc.ctypes.data, v.ctypes.data, extra.ctypes.data, counts.ctypes.data)
    assert rc == 0
    if impl == "port":
        return ccc[:counts[0]].copy(), int(counts[1]), int(counts[3])
    _, d, p, vsl, mu, ints, w = _grt_state(thick, vp, vs, rho, freq, modetype, 3.0)
    global _love_f2c, _rayleigh_f2c
    if modetype == 0:
        if _love_f2c is None:
            _love_f2c = C.CDLL(LOVE_F2C_LIB)
        g = _love_f2c.ref_love_cinterval
    else:
        if _rayleigh_f2c is None:
            _rayleigh_f2c = C.CDLL(RAYLEIGH_F2C_LIB)
        g = _rayleigh_f2c.ref_rayleigh_cinterval
    g.argtypes = [C.c_int, vpt, vpt, vpt, vpt, C.c_int] + [C.c_double] * 5 + [vpt, vpt]
    out, cnt = np.zeros(20008), np.zeros(2, np.int32)
    g(n, d.ctypes.data, p.ctypes.data, vsl.ctypes.data, v.ctypes.data, int(counts[2]), extra[0], extra[1], extra[2], w, tol, out.ctypes.data,
      cnt.ctypes.data)
    return out[:cnt[0]].copy(), int(cnt[1]), int(counts[3])


def grt_setup(impl, thick, vp, vs, rho, modetype):
    """Everything setup_grt leaves in the T_GRT of a column.  impl "port": the restatement; "reference": the translated
    surfmodes.f90:320-450 on a T_GRT initialised as init_grt does.  Returns (rc, mu, v[:nv], lvls, ints, dbl)."""
    vpt = C.c_void_p
    a = [f64(x) for x in (thick, vp, vs, rho)]
    n = len(a[0])
    mu, v, lvls, ints, dbl = np.zeros(n), np.zeros(2 * n), np.zeros(n // 2 + 1, np.int32), np.zeros(8, np.int32), np.zeros(5)
    if impl == "port":
        fn = L().orc_grt_setup
        fn.argtypes = [vpt] * 4 + [C.c_int, C.c_int] + [vpt] * 5
        rc = fn(*[x.ctypes.data for x in a], n, modetype, mu.ctypes.data, v.ctypes.data, lvls.ctypes.data, ints.ctypes.data, dbl.ctypes.data)
    else:
        global _rayleigh_f2c
        if _rayleigh_f2c is None:
            _rayleigh_f2c = C.CDLL(RAYLEIGH_F2C_LIB)
        fn = _rayleigh_f2c.ref_setup_grt
        fn.argtypes = [C.c_int] + [vpt] * 4 + [C.c_int] + [vpt] * 5
        rc = fn(n, *[x.ctypes.data for x in a], modetype, mu.ctypes.data, v.ctypes.data, lvls.ctypes.data, ints.ctypes.data, dbl.ctypes.data)
    return rc, mu, v[:ints[7]].copy(), lvls, ints, dbl


def grt_love_modes_reference(thick, vp, vs, rho, freqs, dc=1e-3, par=GRT_PAR_LIKELIHOOD, group=False):
    """surfmodes for a Love column with a low-velocity layer, phase velocities, through the TRANSLATED reference (setup_grt,
    C_Interval_L, FundaMode, SecFuns_L, bisecim ...; the frequency loop and SearchLove's five calls are the driver's).
    Returns (ierr, phase)."""
    global _love_f2c
    if _love_f2c is None:
        _love_f2c = C.CDLL(LOVE_F2C_LIB)
    vpt = C.c_void_p
    fn = _love_f2c.ref_love_modes
    fn.argtypes = [C.c_int] + [vpt] * 4 + [C.c_int, vpt, C.c_double, vpt, vpt, vpt]
    a = [f64(x) for x in (thick, vp, vs, rho)]
    freqs, par = f64(freqs), f64(np.array(par))
    ph = np.full(len(freqs), 100.0)
    gr = np.full(len(freqs), 100.0)
    ierr = fn(len(a[0]), *[x.ctypes.data for x in a], len(freqs), freqs.ctypes.data, dc, par.ctypes.data, ph.ctypes.data,
              gr.ctypes.data if group else None)
    return (ierr, ph, gr) if group else (ierr, ph)


def grt_rayleigh_modes_reference(thick, vp, vs, rho, freqs, dc=1e-3, par=GRT_PAR_LIKELIHOOD, group=False):
    """surfmodes for a Rayleigh column with a low-velocity layer and no water, phase velocities, through the TRANSLATED reference
    (setup_grt, C_Interval, FundaMode with CR0_Finder, SecFunSurf, bisecim ...).  Returns (ierr, phase); ierr -2: no low-velocity
    layer, -3: a water layer."""
    global _rayleigh_f2c
    if _rayleigh_f2c is None:
        _rayleigh_f2c = C.CDLL(RAYLEIGH_F2C_LIB)
    vpt = C.c_void_p
    fn = _rayleigh_f2c.ref_rayleigh_modes
    fn.argtypes = [C.c_int] + [vpt] * 4 + [C.c_int, vpt, C.c_double, vpt, vpt, vpt]
    a = [f64(x) for x in (thick, vp, vs, rho)]
    freqs, par = f64(freqs), f64(np.array(par))
    ph = np.full(len(freqs), 100.0)
    gr = np.full(len(freqs), 100.0)
    ierr = fn(len(a[0]), *[x.ctypes.data for x in a], len(freqs), freqs.ctypes.data, dc, par.ctypes.data, ph.ctypes.data,
              gr.ctypes.data if group else None)
    return (ierr, ph, gr) if group else (ierr, ph)


def grt_stoneley_secfun_reference(thick, vp, vs, rho, freq, c):
    """startl + SecFunSt(ifs, c, GRT, Imf) of the TRANSLATED Rayleigh.f90 (a column with a water layer on top) on the T_GRT the
    restatement's setup_grt builds.  Returns (value, imf, ll of the translated startl, ll of the restatement)."""
    global _rayleigh_f2c
    if _rayleigh_f2c is None:
        _rayleigh_f2c = C.CDLL(RAYLEIGH_F2C_LIB)
    vpt = C.c_void_p
    n, d, p, v, mu, ints, w = _grt_state(thick, vp, vs, rho, freq, 1, c)
    assert ints[0] == 1, "no water layer: SecFunSurf, not SecFunSt"
    mu0 = grt_setup("port", thick, vp, vs, rho, 1)[5][0]
    r = f64(rho)
    fn = _rayleigh_f2c.ref_rayleigh_secfunst
    fn.argtypes = [C.c_int] + [vpt] * 5 + [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, vpt, vpt, vpt]
    val, imf, ll = C.c_double(0), C.c_double(0), C.c_int(0)
    fn(n, d.ctypes.data, p.ctypes.data, v.ctypes.data, r.ctypes.data, mu.ctypes.data, float(mu0), int(ints[0]), int(ints[2]), w, c,
       C.byref(val), C.byref(imf), C.byref(ll))
    return val.value, imf.value, ll.value, int(ints[1])


def live_seed(tag):
    """Seed of a live comparison against a translated reference: fixed per test by default (a CI run must not depend on the
    draw); MCT_TEST_SEED=random draws a fresh one (what the development soak used), MCT_TEST_SEED=<int> replays one."""
    import zlib
    v = os.environ.get("MCT_TEST_SEED", "")
    if v == "random":
        return int.from_bytes(os.urandom(4), "little")
    if v:
        return int(v) + zlib.crc32(tag.encode()) % 1000
    return zlib.crc32(tag.encode())
