"""ctypes binding of libmctomo_b200.so (include/mctomo_b200.h) plus a thin host-side mirror of
the two reference subroutines the library stands behind:

  kdtree_to_grid(RTI, grid, bnd_box, model, pm)      reference src/mcmc_loc2.f90:2002-2082
  surf_likelihood's dispersion block                  reference src/likelihood_surf.F90:155-231

The production host is Fortran (fortran/mctomo_b200_shim.f90); this module exists so the parity
tests and bench.py can drive the very same C entry points from Python.  There is no fallback:
if the shared library is missing or no CUDA device is usable, importing/initialising raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCT_LIB", os.path.join(_HERE, "libmctomo_b200.so"))  # MCT_LIB: A/B against another build

MCT_OK = 0
MCT_E_INVALID_ARG = 1
MCT_E_GRT_NEEDED = 2
MCT_E_TOO_MANY_LAYERS = 3
MCT_E_DEGENERATE_NUCLEI = 4
MCT_E_FLUID_BELOW_TOP = 5
MCT_E_ZERO_NOISE = 6
MCT_E_FM2D_STALE = 7
MCT_E_NOINIT = -1
MCT_E_CUDA = -2

# default-real Fortran literals widened to double, as the reference stores them
EPS_LIKELIHOOD = float(np.float32(1e-10))   # likelihood_surf.F90:37
EPS_MODELLING = float(np.float32(1e-5))     # forward_modelling.f90:27
DPHASE_DEFAULT = 1e-3                       # examples/example1/MCTomo.inp:48


class MctError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mctomo_b200 error {code}: {msg}")
        self.code = code


class mct_grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("xmin", C.c_double), ("ymin", C.c_double), ("zmin", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("waterDepth", C.c_double), ("scaling", C.c_double)]


class mct_disp_opts(C.Structure):
    _fields_ = [("raylov", C.c_int32), ("phaseGroup", C.c_int32), ("nmodes", C.c_int32), ("check_scope", C.c_int32),
                ("dphase", C.c_double), ("layer_eps", C.c_double), ("water_thresh", C.c_double),
                ("preset", C.c_double)]


class mct_stats(C.Structure):
    _fields_ = [("n_dltar", C.c_int64), ("n_layer_steps", C.c_int64), ("n_columns", C.c_int64),
                ("n_nodes", C.c_int64), ("n_launches", C.c_int64), ("n_dltar_executed", C.c_int64),
                ("n_layer_steps_executed", C.c_int64), ("n_columns_solved", C.c_int64)]


class mct_launch_info(C.Structure):
    _fields_ = [("kernel", C.c_char * 48), ("columns", C.c_int32), ("columns_solved", C.c_int32),
                ("lanes_per_column", C.c_int32), ("sm_count", C.c_int32)]


@dataclass
class Grid:
    """Mirror of T_GRID + grid_setup (reference src/settings.f90:20-28,100-123)."""
    nx: int
    ny: int
    nz: int
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    zmin: float
    zmax: float
    waterDepth: float = 0.0
    scaling: float = 1.0

    def __post_init__(self):
        # grid_setup: zmin/zmax are scaled, spacing = extent/(n-1)
        self.zmin = self.zmin * self.scaling
        self.zmax = self.zmax * self.scaling
        # (a single node along an axis -- e.g. ny = 1 for the 2-D product's depth profiles -- gets unit spacing)
        self.dx = (self.xmax - self.xmin) / (self.nx - 1) if self.nx > 1 else 1.0
        self.dy = (self.ymax - self.ymin) / (self.ny - 1) if self.ny > 1 else 1.0
        self.dz = (self.zmax - self.zmin) / (self.nz - 1) if self.nz > 1 else 1.0

    def c(self) -> mct_grid:
        return mct_grid(self.nx, self.ny, self.nz, self.xmin, self.ymin, self.zmin, self.dx, self.dy, self.dz,
                        self.waterDepth, self.scaling)

    @property
    def shape(self):
        """numpy shape of a (nz,ny,nx) Fortran array viewed C-contiguously: [ix][iy][iz]."""
        return (self.nx, self.ny, self.nz)

    def full_box(self):
        """The box the reference passes for "the whole grid": (xmin,ymin,zmin)-(xmax,ymax,zmax).  NOTE: the
        reference turns a box into indices with floor((b-min)/d)+1 (src/mcmc_loc2.f90:2034-2045); when
        (zmax-zmin)/dz rounds just below nz-1 (e.g. nz = 60 over 12 km) the LAST PLANE is not covered.  That
        is the reference's behaviour and is reproduced; use cover_box() to address every node."""
        return np.array([self.xmin, self.ymin, self.zmin, self.xmax, self.ymax, self.zmax], dtype=np.float64)

    def cover_box(self):
        """full_box() grown by half a cell on every side: its index window is exactly 1..n in each dimension."""
        h = 0.5 * np.array([self.dx, self.dy, self.dz])
        lo = np.array([self.xmin, self.ymin, self.zmin]) - h
        hi = np.array([self.xmax, self.ymax, self.zmax]) + h
        return np.concatenate([lo, hi]).astype(np.float64)


def disp_opts(raylov=1, phaseGroup=0, nmodes=0, dphase=DPHASE_DEFAULT, variant="likelihood", check_scope=0) -> mct_disp_opts:
    """variant 'likelihood' = likelihood_surf.F90 constants, 'modelling' = forward_modelling.f90 constants;
    check_scope 0 = check_model over the whole grid (reference), 1 = over the window's columns only."""
    if variant == "likelihood":
        return mct_disp_opts(raylov, phaseGroup, nmodes, check_scope, dphase, EPS_LIKELIHOOD, EPS_LIKELIHOOD, 100.0)
    if variant == "modelling":
        return mct_disp_opts(raylov, phaseGroup, nmodes, check_scope, dphase, EPS_MODELLING, 0.0, 1000.0)
    raise ValueError(variant)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    gp = C.POINTER(mct_grid)
    op = C.POINTER(mct_disp_opts)
    L.mct_last_error.restype = C.c_char_p
    L.mct_init.argtypes = [C.c_int]
    L.mct_get_stats.argtypes = [C.POINTER(mct_stats)]
    L.mct_set_counters.argtypes = [C.c_int]
    L.mct_box_window.argtypes = [gp, vp, vp]
    L.mct_voronoi_to_grid.argtypes = [vp, vp, C.c_int, gp, vp, vp, vp, vp, vp, vp]
    L.mct_voronoi_to_grid_dev.argtypes = [vp, vp, C.c_int, gp, vp, vp, vp, vp, vp, vp, vp]
    L.mct_vs2vp_rho.argtypes = [vp, vp, vp, C.c_int64]
    L.mct_vs2vp_rho_dev.argtypes = [vp, vp, vp, C.c_int64, vp]
    L.mct_surf_dispersion.argtypes = [vp, vp, vp, gp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, op, vp, vp, vp, vp]
    L.mct_surf_dispersion_dev.argtypes = [vp, vp, vp, gp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, op, vp, vp, vp, vp, vp]
    L.mct_surfmodes_batch.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, op, vp, vp, vp]
    L.mct_forward_eval.argtypes = [vp, vp, C.c_int, gp, C.c_int, vp, C.c_int, op, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mct_forward_eval_dev.argtypes = [vp, vp, C.c_int, gp, C.c_int, C.c_int, C.c_int, vp, C.c_int, op,
                                       vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mct_assemble_vel_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    _lib = L
    return L


def _check(rc: int, allow=()):
    if rc != MCT_OK and rc not in allow:
        raise MctError(rc, lib().mct_last_error().decode())
    return rc


def init(device: int = 0):
    _check(lib().mct_init(device))


def shutdown():
    _check(lib().mct_shutdown())


def stats() -> dict:
    s = mct_stats()
    _check(lib().mct_get_stats(C.byref(s)))
    return {k: getattr(s, k) for k, _ in mct_stats._fields_}


def reset_stats():
    _check(lib().mct_reset_stats())


def set_counters(on: bool):
    _check(lib().mct_set_counters(1 if on else 0))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


def box_window(grid: Grid, box) -> np.ndarray:
    w = np.zeros(6, np.int32)
    b = _f64(box)
    _check(lib().mct_box_window(C.byref(grid.c()), b.ctypes.data, w.ctypes.data))
    return w


def kdtree_to_grid(points, params, grid: Grid, bnd_box, vp, vs, rho, sites_id, pm=None):
    """Drop-in for kdtree_to_grid (reference src/mcmc_loc2.f90:2002-2082).

    points, params: (ncells,3) C-contiguous == Fortran (3,ncells); vp, vs, rho (float64) and sites_id
    (int32) have shape grid.shape == Fortran (nz,ny,nx) and are updated IN PLACE inside the window.
    """
    points = _f64(points)
    params = _f64(params)
    for a, dt in ((vp, np.float64), (vs, np.float64), (rho, np.float64), (sites_id, np.int32)):
        assert a.dtype == dt and a.flags.c_contiguous and a.shape == grid.shape
    box = _f64(bnd_box)
    pmv = None if pm is None else _f64(pm)
    _check(lib().mct_voronoi_to_grid(points.ctypes.data, params.ctypes.data, len(points), C.byref(grid.c()),
                                     box.ctypes.data, _ptr(pmv), vp.ctypes.data, vs.ctypes.data, rho.ctypes.data,
                                     sites_id.ctypes.data))


def vs2vp_rho(vs):
    """vs2vp_3d + vp2rho_3d (reference src/utils.f90:102-134)."""
    vs = _f64(vs)
    vp = np.empty_like(vs)
    rho = np.empty_like(vs)
    _check(lib().mct_vs2vp_rho(vs.ctypes.data, vp.ctypes.data, rho.ctypes.data, vs.size))
    return vp, rho


def vs2vp_rho_window(vs, vp, rho, grid: Grid, w):
    """vs2vp_3d + vp2rho_3d restricted to the index window w = (ix0,ix1,iy0,iy1,iz0,iz1); vp, rho updated in place."""
    L = lib()
    L.mct_vs2vp_rho_window.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(mct_grid), C.c_void_p]
    for a in (vs, vp, rho):
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == grid.shape
    wv = np.ascontiguousarray(w, dtype=np.int32)
    _check(L.mct_vs2vp_rho_window(vs.ctypes.data, vp.ctypes.data, rho.ctypes.data, C.byref(grid.c()), wv.ctypes.data))


def surf_dispersion(vp, vs, rho, grid: Grid, window, freqs, opts: mct_disp_opts, check=True):
    """The dispersion block of surf_likelihood (reference src/likelihood_surf.F90:155-231).

    window = (ix0, ix1, iy0, iy1), 1-based inclusive.  Returns (pvel, gvel, ierr, model_invalid, rc) with
    pvel/gvel of shape (wx, wy, nm*np) and ierr (wx, wy).
    """
    vp, vs, rho, freqs = _f64(vp), _f64(vs), _f64(rho), _f64(freqs)
    ix0, ix1, iy0, iy1 = (int(v) for v in window)
    wx, wy = ix1 - ix0 + 1, iy1 - iy0 + 1
    nm = max(opts.nmodes, 1)
    pvel = np.zeros((wx, wy, nm * len(freqs)))
    gvel = np.zeros_like(pvel)
    ierr = np.zeros((wx, wy), np.int32)
    inval = C.c_int32(0)
    rc = _check(lib().mct_surf_dispersion(vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, C.byref(grid.c()), ix0, ix1,
                                          iy0, iy1, freqs.ctypes.data, len(freqs), C.byref(opts), pvel.ctypes.data,
                                          gvel.ctypes.data, ierr.ctypes.data, C.byref(inval) if check else None),
                allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))
    return pvel, gvel, ierr, inval.value, rc


def surfmodes_batch(thick, vp, vs, rho, offsets, freqs, opts: mct_disp_opts):
    """surfmodes / surfmmodes over pre-layered columns (reference surfmodes/surfmodes.f90:39-183)."""
    thick, vp, vs, rho, freqs = _f64(thick), _f64(vp), _f64(vs), _f64(rho), _f64(freqs)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    ncol = len(offsets) - 1
    nm = max(opts.nmodes, 1)
    phase = np.zeros((ncol, nm * len(freqs)))
    group = np.zeros_like(phase)
    ierr = np.zeros(ncol, np.int32)
    rc = _check(lib().mct_surfmodes_batch(thick.ctypes.data, vp.ctypes.data, vs.ctypes.data, rho.ctypes.data,
                                          offsets.ctypes.data, ncol, freqs.ctypes.data, len(freqs), C.byref(opts),
                                          phase.ctypes.data, group.ctypes.data, ierr.ctypes.data),
                allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))
    return phase, group, ierr, rc


def forward_eval(points, params, grid: Grid, freqs, opts: mct_disp_opts, derive_vp_rho=True, want_model=False,
                 out=None):
    """kdtree_to_grid (full box) -> vs2vp/vp2rho -> check_model -> dispersion, model resident in HBM.

    `out` may hold preallocated (ideally pinned) arrays 'pvel','gvel','ierr' to avoid allocations.
    """
    points, params, freqs = _f64(points), _f64(params), _f64(freqs)
    nm = max(opts.nmodes, 1)
    if out is None:
        out = {}
    pvel = out.get("pvel")
    if pvel is None:
        pvel = np.zeros((grid.nx, grid.ny, nm * len(freqs)))
    gvel = out.get("gvel")
    if gvel is None:
        gvel = np.zeros_like(pvel)
    ierr = out.get("ierr")
    if ierr is None:
        ierr = np.zeros((grid.nx, grid.ny), np.int32)
    inval = C.c_int32(0)
    vp = vs = rho = sid = None
    if want_model:
        vp = np.zeros(grid.shape)
        vs = np.zeros(grid.shape)
        rho = np.zeros(grid.shape)
        sid = np.zeros(grid.shape, np.int32)
    rc = _check(lib().mct_forward_eval(points.ctypes.data, params.ctypes.data, len(points), C.byref(grid.c()),
                                       1 if derive_vp_rho else 0, freqs.ctypes.data, len(freqs), C.byref(opts),
                                       pvel.ctypes.data, gvel.ctypes.data, ierr.ctypes.data, C.byref(inval), _ptr(vp),
                                       _ptr(vs), _ptr(rho), _ptr(sid)),
                allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))
    res = {"pvel": pvel, "gvel": gvel, "ierr": ierr, "model_invalid": inval.value, "rc": rc}
    if want_model:
        res.update(vp=vp, vs=vs, rho=rho, sites_id=sid)
    return res


# ---- device-pointer forms (torch tensors give the pointers and the stream) -----------------------

def forward_eval_dev(points, params, grid: Grid, freqs, opts: mct_disp_opts, d_vp, d_vs, d_rho, d_sites, d_pvel,
                     d_gvel, d_ierr, d_flags, stream, derive_vp_rho=True, slab=None):
    """mct_forward_eval_dev: arguments are raw device pointers (ints) and a cudaStream_t handle (int)."""
    points, params, freqs = _f64(points), _f64(params), _f64(freqs)
    ixs0, ixs1 = (1, grid.nx) if slab is None else slab
    return _check(lib().mct_forward_eval_dev(points.ctypes.data, params.ctypes.data, len(points), C.byref(grid.c()),
                                             1 if derive_vp_rho else 0, ixs0, ixs1, freqs.ctypes.data, len(freqs),
                                             C.byref(opts), d_vp, d_vs, d_rho, d_sites, d_pvel, d_gvel, d_ierr, d_flags,
                                             stream))


def surf_dispersion_dev(d_vp, d_vs, d_rho, grid: Grid, window, freqs, opts, d_pvel, d_gvel, d_ierr, d_flags, stream):
    freqs = _f64(freqs)
    ix0, ix1, iy0, iy1 = (int(v) for v in window)
    return _check(lib().mct_surf_dispersion_dev(d_vp, d_vs, d_rho, C.byref(grid.c()), ix0, ix1, iy0, iy1,
                                                freqs.ctypes.data, len(freqs), C.byref(opts), d_pvel, d_gvel, d_ierr,
                                                d_flags, stream))


def voronoi_to_grid_dev(points, params, grid: Grid, box, d_vp, d_vs, d_rho, d_sites, stream, pm=None):
    points, params, box = _f64(points), _f64(params), _f64(box)
    pmv = None if pm is None else _f64(pm)
    return _check(lib().mct_voronoi_to_grid_dev(points.ctypes.data, params.ctypes.data, len(points), C.byref(grid.c()),
                                                box.ctypes.data, _ptr(pmv), d_vp, d_vs, d_rho, d_sites, stream))


def accumulate_stats_dev(d_vs, d_vp, d_aveS, d_stdS, d_aveP, d_stdP, n, stream):
    """aveS += vs, stdS += vs**2, aveP += vp, stdP += vp**2 on the device (reference src/sample.f90:483-486)."""
    L = lib()
    L.mct_accumulate_stats_dev.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
    return _check(L.mct_accumulate_stats_dev(d_vs, d_vp, d_aveS, d_stdS, d_aveP, d_stdP, n, stream))


def assemble_vel_dev(d_pvel, np_, nx, ny, window, d_vel, stream):
    ix0, ix1, iy0, iy1 = (int(v) for v in window)
    return _check(lib().mct_assemble_vel_dev(d_pvel, np_, nx, ny, ix0, ix1, iy0, iy1, d_vel, stream))


# ---- batches of independent models (chains sharing one GPU) ---------------------------------------

def _bind_batch():
    L = lib()
    if getattr(L, "_batch_bound", False):
        return L
    vp = C.c_void_p
    gp = C.POINTER(mct_grid)
    op = C.POINTER(mct_disp_opts)
    L.mct_set_nuclei_batch.argtypes = [vp, vp, vp, C.c_int]
    L.mct_forward_batch_dev.argtypes = [gp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, op] + [vp] * 9
    L.mct_forward_eval_batch.argtypes = [vp, vp, vp, C.c_int, gp, C.c_int, vp, C.c_int, op] + [vp] * 8
    L.mct_set_profiling.argtypes = [C.c_int]
    L.mct_kernel_times.argtypes = [vp, C.c_int]
    L.mct_fp64_peak_probe.argtypes = [vp, vp]
    L._batch_bound = True
    return L


def pack_models(models):
    """models: list of (points (n_b,3), params (n_b,3)) -> (points_all, params_all, offsets int64[nb+1])."""
    pts = np.ascontiguousarray(np.concatenate([np.asarray(m[0], dtype=np.float64) for m in models]))
    par = np.ascontiguousarray(np.concatenate([np.asarray(m[1], dtype=np.float64) for m in models]))
    off = np.zeros(len(models) + 1, np.int64)
    off[1:] = np.cumsum([len(m[0]) for m in models])
    return pts, par, off


def set_nuclei_batch(points, params, offsets):
    L = _bind_batch()
    points, params = _f64(points), _f64(params)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    _check(L.mct_set_nuclei_batch(points.ctypes.data, params.ctypes.data, offsets.ctypes.data, len(offsets) - 1))


def forward_batch_dev(grid: Grid, nb, freqs, opts, d_vp, d_vs, d_rho, d_sites, d_pvel, d_gvel, d_ierr, d_flags, stream,
                      derive_vp_rho=True, slab=None):
    L = _bind_batch()
    freqs = _f64(freqs)
    ixs0, ixs1 = (1, grid.nx) if slab is None else slab
    return _check(L.mct_forward_batch_dev(C.byref(grid.c()), nb, 1 if derive_vp_rho else 0, ixs0, ixs1, freqs.ctypes.data,
                                          len(freqs), C.byref(opts), d_vp, d_vs, d_rho, d_sites, d_pvel, d_gvel, d_ierr,
                                          d_flags, stream))


def forward_eval_batch(points, params, offsets, grid: Grid, freqs, opts, derive_vp_rho=True, out=None, want_model=False):
    """nb independent models in one call: nuclei (host) in, dispersion maps (host) out."""
    L = _bind_batch()
    points, params, freqs = _f64(points), _f64(params), _f64(freqs)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nb = len(offsets) - 1
    nm = max(opts.nmodes, 1)
    out = {} if out is None else out
    pvel = out.get("pvel")
    if pvel is None:
        pvel = np.zeros((nb, grid.nx, grid.ny, nm * len(freqs)))
    gvel = out.get("gvel")
    if gvel is None:
        gvel = np.zeros_like(pvel)
    ierr = out.get("ierr")
    if ierr is None:
        ierr = np.zeros((nb, grid.nx, grid.ny), np.int32)
    inval = np.zeros(nb, np.int32)
    vp = vs = rho = sid = None
    if want_model:
        vp, vs, rho = (np.zeros((nb,) + grid.shape) for _ in range(3))
        sid = np.zeros((nb,) + grid.shape, np.int32)
    rc = _check(L.mct_forward_eval_batch(points.ctypes.data, params.ctypes.data, offsets.ctypes.data, nb,
                                         C.byref(grid.c()), 1 if derive_vp_rho else 0, freqs.ctypes.data, len(freqs),
                                         C.byref(opts), pvel.ctypes.data, gvel.ctypes.data, ierr.ctypes.data,
                                         inval.ctypes.data, _ptr(vp), _ptr(vs), _ptr(rho), _ptr(sid)),
                allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))
    res = {"pvel": pvel, "gvel": gvel, "ierr": ierr, "model_invalid": inval, "rc": rc}
    if want_model:
        res.update(vp=vp, vs=vs, rho=rho, sites_id=sid)
    return res


def set_profiling(on: bool):
    _check(_bind_batch().mct_set_profiling(1 if on else 0))


def kernel_times(reset=True) -> dict:
    ms = np.zeros(4)
    _check(_bind_batch().mct_kernel_times(ms.ctypes.data, 1 if reset else 0))
    return {"k1_ms": ms[0], "k2_ms": ms[1], "other_ms": ms[2], "launches": int(ms[3])}


def fp64_peak_probe() -> dict:
    a = C.c_double(0)
    b = C.c_double(0)
    _check(_bind_batch().mct_fp64_peak_probe(C.byref(a), C.byref(b)))
    return {"dfma_tflops": a.value, "dmul_dadd_tflops": b.value}


class mct_fm2d_opts(C.Structure):
    _fields_ = [("gridx", C.c_int32), ("gridy", C.c_int32), ("sgref", C.c_int32), ("sgdic", C.c_int32), ("sgext", C.c_int32),
                ("order", C.c_int32), ("band", C.c_double)]


LAST_FM2D_RC = 0  # status of the last fm2d_times / fm2d_rays call (MCT_E_FM2D_STALE: a source in the model's last cell row/column)


def fm2d_opts(gridx=1, gridy=1, sgref=1, sgdic=4, sgext=8, order=1, band=0.5) -> mct_fm2d_opts:
    """examples/example1/MCTomo.inp:61-72 by default."""
    return mct_fm2d_opts(gridx, gridy, sgref, sgdic, sgext, order, band)


def fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, opts: mct_fm2d_opts, ttime=None, want_field=False):
    """modrays for phase-velocity data (travel times only): src (nsrc,2), rcv (nrc,2) as (x, z); srs (nmaps, nsrc, nrc) 0/1;
    vel (nmaps, nvx+2, nvz+2) C-order = the Fortran (nvz+2, nvx+2, nmaps).  Returns (ttime (nmaps, nsrc, nrc), field or None)."""
    L = lib()
    vp = C.c_void_p
    L.mct_fm2d_times.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + \
                                [C.POINTER(mct_fm2d_opts), vp, vp]
    src, rcv, vel = _f64(src), _f64(rcv), _f64(vel)
    nmaps, nsrc, nrc = vel.shape[0], len(src), len(rcv)
    nvx, nvz = vel.shape[1] - 2, vel.shape[2] - 2
    sx, sz = _f64(src[:, 0].copy()), _f64(src[:, 1].copy())
    rx, rz = _f64(rcv[:, 0].copy()), _f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(np.broadcast_to(srs, (nmaps, nsrc, nrc)), dtype=np.int32)
    tt = np.full((nmaps, nsrc, nrc), -1.0) if ttime is None else np.ascontiguousarray(ttime, dtype=np.float64)
    nnx, nnz = (nvx - 1) * opts.gridx + 1, (nvz - 1) * opts.gridy + 1
    field = np.zeros((nmaps, nsrc, nnx, nnz)) if want_field else None
    global LAST_FM2D_RC
    LAST_FM2D_RC = _check(L.mct_fm2d_times(sx.ctypes.data, sz.ctypes.data, nsrc, rx.ctypes.data, rz.ctypes.data, nrc, srs.ctypes.data, vel.ctypes.data,
                                           nmaps, nvx, nvz, gox, goz, dvx, dvz, C.byref(opts), tt.ctypes.data, field.ctypes.data if want_field else None),
                          allow=(MCT_E_FM2D_STALE,))
    return tt, field


def fm2d_rays(src, rcv, srs, vel, gox, goz, dvx, dvz, opts: mct_fm2d_opts, srsv=None, cap=None):
    """modrays with uar = 0: travel times and ray geometry.  Returns dict(ttime (nmaps,nsrc,nrc), npts (nmaps,nrr), pts (nmaps,nrr,cap,2),
    length (nmaps,nrr), crazy (nmaps))."""
    L = lib()
    vp = C.c_void_p
    L.mct_fm2d_rays.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + \
                               [C.POINTER(mct_fm2d_opts), vp, C.c_int, vp, vp, vp, vp]
    src, rcv, vel = _f64(src), _f64(rcv), _f64(vel)
    nmaps, nsrc, nrc = vel.shape[0], len(src), len(rcv)
    nrr = nsrc * nrc
    nvx, nvz = vel.shape[1] - 2, vel.shape[2] - 2
    sx, sz = _f64(src[:, 0].copy()), _f64(src[:, 1].copy())
    rx, rz = _f64(rcv[:, 0].copy()), _f64(rcv[:, 1].copy())
    srs = np.ascontiguousarray(np.broadcast_to(srs, (nmaps, nsrc, nrc)), dtype=np.int32)
    if srsv is None:
        srsv = np.arange(1, nrr + 1, dtype=np.int32).reshape(nsrc, nrc)
    srsv = np.ascontiguousarray(np.broadcast_to(srsv, (nmaps, nsrc, nrc)), dtype=np.int32)
    if cap is None:
        cap = 8 * (nvx * opts.gridx + nvz * opts.gridy)
    tt = np.full((nmaps, nsrc, nrc), -1.0)
    npts = np.zeros((nmaps, nrr), np.int32)
    pts = np.zeros((nmaps, nrr, cap, 2))
    ln = np.zeros((nmaps, nrr))
    crazy = np.zeros(nmaps, np.int32)
    global LAST_FM2D_RC
    LAST_FM2D_RC = _check(L.mct_fm2d_rays(sx.ctypes.data, sz.ctypes.data, nsrc, rx.ctypes.data, rz.ctypes.data, nrc, srs.ctypes.data, srsv.ctypes.data,
                                          vel.ctypes.data, nmaps, nvx, nvz, gox, goz, dvx, dvz, C.byref(opts), tt.ctypes.data, cap, npts.ctypes.data,
                                          pts.ctypes.data, ln.ctypes.data, crazy.ctypes.data), allow=(MCT_E_FM2D_STALE,))
    return dict(ttime=tt, npts=npts, pts=pts, length=ln, crazy=crazy)


def fm2d_times_dev(d_src_xz, nsrc, d_rcv_xz, nrc, d_srs, srs_map_stride, d_vel, vel_elem_stride, vel_map_stride, nmaps, nvx, nvz, gox, goz,
                   dvx, dvz, opts: mct_fm2d_opts, d_ttime, d_err, stream):
    """Device-pointer form (asynchronous on `stream`); see include/mctomo_b200.h."""
    L = lib()
    vp = C.c_void_p
    L.mct_fm2d_times_dev.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_longlong, vp, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int] + \
                                    [C.c_double] * 4 + [C.POINTER(mct_fm2d_opts), vp, vp, vp]
    return _check(L.mct_fm2d_times_dev(d_src_xz, nsrc, d_rcv_xz, nrc, d_srs, srs_map_stride, d_vel, vel_elem_stride, vel_map_stride, nmaps,
                                       nvx, nvz, gox, goz, dvx, dvz, C.byref(opts), d_ttime, d_err, stream))


def fm2d_stats():
    L = lib()
    L.mct_fm2d_stats.argtypes = [C.c_void_p]
    out = np.zeros(2, np.int64)
    _check(L.mct_fm2d_stats(out.ctypes.data))
    return {"accepted": int(out[0]), "updates": int(out[1])}


def set_grt(enable: bool, par6=None):
    """Solve low-velocity columns with the generalized R/T kernel (surfmodes.f90:84-87,96-99) instead of reporting ierr = 2."""
    L = lib()
    L.mct_set_grt.argtypes = [C.c_int, C.c_void_p]
    if par6 is None:
        _check(L.mct_set_grt(1 if enable else 0, None))
    else:
        p = np.ascontiguousarray(par6, dtype=np.float64)
        assert p.size == 6
        _check(L.mct_set_grt(1 if enable else 0, p.ctypes.data))


def grt_stats():
    L = lib()
    L.mct_grt_stats.argtypes = [C.c_void_p]
    out = np.zeros(3, np.int64)
    _check(L.mct_grt_stats(out.ctypes.data))
    return {"columns": int(out[0]), "secfun": int(out[1]), "interface_steps": int(out[2])}


def set_k1_mode(mode: int = 0):
    """0 culled brute force per column (default), 1 kdtree2 traversal for every node."""
    L = _bind_batch()
    L.mct_set_k1_mode.argtypes = [C.c_int]
    _check(L.mct_set_k1_mode(mode))


def set_k2_mode(mode: int = 0, coop_max_columns: int = -1):
    """0 auto (lane groups per column below coop_max_columns, else one thread per column), 1 always one thread per column,
    2 always lane groups (size: set_k2_lanes)."""
    L = _bind_batch()
    L.mct_set_k2_mode.argtypes = [C.c_int, C.c_int]
    _check(L.mct_set_k2_mode(mode, coop_max_columns))


def set_dedup(on: bool = True):
    _check(lib().mct_set_dedup(1 if on else 0))


def last_launch() -> dict:
    li = mct_launch_info()
    _check(lib().mct_last_launch(C.byref(li)))
    return dict(kernel=li.kernel.decode(), columns=li.columns, columns_solved=li.columns_solved,
                lanes_per_column=li.lanes_per_column, sm_count=li.sm_count)


def set_k2_lanes(lanes_per_column: int = 0):
    """Lanes per column of the cooperative dispersion kernel: 0 automatic, else a power of two from 2 to 256."""
    L = _bind_batch()
    L.mct_set_k2_lanes.argtypes = [C.c_int]
    _check(L.mct_set_k2_lanes(lanes_per_column))


def selftest_division(emax: int = 300):
    """(tested, mismatches) of the shared-reciprocal division against IEEE `/` on the device."""
    L = _bind_batch()
    L.mct_selftest_division.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    t = C.c_int64(0)
    m = C.c_int64(0)
    _check(L.mct_selftest_division(emax, C.byref(t), C.byref(m)))
    return t.value, m.value


def group_times_dev(d_vel, np_, grid: Grid, ray_points, ray_offsets, nrays, stream=None):
    """mct_group_times_dev: travel times (np, nrays) of packed host rays through a device (np,ny,nx) map."""
    L = lib()
    L.mct_group_times_dev.argtypes = [C.c_void_p, C.c_int, C.POINTER(mct_grid), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    pts = _f64(ray_points)
    off = np.ascontiguousarray(ray_offsets, dtype=np.int64)
    t = np.zeros((np_, nrays))
    _check(L.mct_group_times_dev(d_vel, np_, C.byref(grid.c()), pts.ctypes.data, off.ctypes.data, nrays, t.ctypes.data, stream))
    return t


class Session:
    """One chain's model resident in HBM between proposals (mct_session_*, include/mctomo_b200.h).

    set_model(points, params) -> full evaluation;  propose(points, params, box[, pm]) -> (window, pvel, gvel, ierr,
    model_invalid) for the box's columns + halo;  accept() / reject() commit or restore."""

    def __init__(self, grid: Grid, freqs, opts: mct_disp_opts, derive_vp_rho=True):
        L = lib()
        vp = C.c_void_p
        L.mct_session_create.argtypes = [C.POINTER(mct_grid), vp, C.c_int, C.POINTER(mct_disp_opts), C.c_int, C.POINTER(vp)]
        L.mct_session_destroy.argtypes = [vp]
        L.mct_session_set_model.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.mct_session_propose.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.mct_session_accept.argtypes = [vp]
        L.mct_session_reject.argtypes = [vp]
        L.mct_session_get_model.argtypes = [vp, vp, vp, vp, vp]
        L.mct_session_get_maps.argtypes = [vp, vp, vp, vp]
        self._L = L
        self.grid = grid
        self.freqs = _f64(freqs)
        self.opts = opts
        self.nout = len(self.freqs) * max(opts.nmodes, 1)
        h = vp()
        _check(L.mct_session_create(C.byref(grid.c()), self.freqs.ctypes.data, len(self.freqs), C.byref(opts),
                                    1 if derive_vp_rho else 0, C.byref(h)))
        self._h = h
        ncol = grid.nx * grid.ny
        self._pv = np.zeros(ncol * self.nout)      # staging for the largest possible window
        self._gv = np.zeros(ncol * self.nout)
        self._ie = np.zeros(ncol, np.int32)
        self._nrays_resident = 0
        self._nrr = 0

    def close(self):
        if self._h:
            self._L.mct_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_model(self, points, params, want_maps=True):
        points, params = _f64(points), _f64(params)
        g = self.grid
        inval = C.c_int32(0)
        pv = np.zeros((g.nx, g.ny, self.nout)) if want_maps else None
        gv = np.zeros((g.nx, g.ny, self.nout)) if want_maps else None
        ie = np.zeros((g.nx, g.ny), np.int32) if want_maps else None
        rc = _check(self._L.mct_session_set_model(self._h, points.ctypes.data, params.ctypes.data, len(points),
                                                  pv.ctypes.data if want_maps else None, gv.ctypes.data if want_maps else None,
                                                  ie.ctypes.data if want_maps else None, C.addressof(inval)),
                    allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))
        return dict(pvel=pv, gvel=gv, ierr=ie, model_invalid=inval.value, rc=rc)

    def propose(self, points, params, box, pm=None):
        points, params, box = _f64(points), _f64(params), _f64(box)
        pmv = None if pm is None else _f64(pm)
        win = np.zeros(4, np.int32)
        inval = C.c_int32(0)
        rc = _check(self._L.mct_session_propose(self._h, points.ctypes.data, params.ctypes.data, len(points), box.ctypes.data,
                                                None if pmv is None else pmv.ctypes.data, win.ctypes.data, self._pv.ctypes.data,
                                                self._gv.ctypes.data, self._ie.ctypes.data, C.addressof(inval)),
                    allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))
        wx, wy = int(win[1] - win[0] + 1), int(win[3] - win[2] + 1)
        n = max(wx, 0) * max(wy, 0)
        pv = self._pv[: n * self.nout].reshape(max(wx, 0), max(wy, 0), self.nout)
        gv = self._gv[: n * self.nout].reshape(max(wx, 0), max(wy, 0), self.nout)
        ie = self._ie[:n].reshape(max(wx, 0), max(wy, 0))
        return dict(window=tuple(int(v) for v in win), pvel=pv, gvel=gv, ierr=ie, model_invalid=inval.value, rc=rc)

    def accept(self):
        _check(self._L.mct_session_accept(self._h))

    def reject(self):
        _check(self._L.mct_session_reject(self._h))

    def get_model(self):
        g = self.grid
        vp, vs, rho = np.zeros(g.shape), np.zeros(g.shape), np.zeros(g.shape)
        sid = np.zeros(g.shape, np.int32)
        _check(self._L.mct_session_get_model(self._h, vp.ctypes.data, vs.ctypes.data, rho.ctypes.data, sid.ctypes.data))
        return vp, vs, rho, sid

    def group_times(self, ray_points=None, ray_offsets=None, nrays=0, pending=False):
        """CalGroupTime through like%gvel (group map when phaseGroup == 1, else phase map) of the current model, or of
        the pending proposal (window maps overlaid): (np, nrays) travel times.  ray_points None: resident rays."""
        fn = self._L.mct_session_group_times_pending if pending else self._L.mct_session_group_times
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        if ray_points is None:
            nrays = self._nrays_resident
            t = np.zeros((len(self.freqs), nrays))
            _check(fn(self._h, None, None, 0, t.ctypes.data))
            return t
        pts = _f64(ray_points)
        off = np.ascontiguousarray(ray_offsets, dtype=np.int64)
        t = np.zeros((len(self.freqs), nrays))
        _check(fn(self._h, pts.ctypes.data, off.ctypes.data, nrays, t.ctypes.data))
        return t

    def set_rays(self, ray_points, ray_offsets, nrays):
        pts = _f64(ray_points)
        off = np.ascontiguousarray(ray_offsets, dtype=np.int64)
        self._L.mct_session_set_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _check(self._L.mct_session_set_rays(self._h, pts.ctypes.data, off.ctypes.data, nrays))
        self._nrays_resident = nrays

    def set_data(self, ttime, raystat, sigdep=0, nrays_total=None, srdist=None):
        """ttime (np,3,nrr), raystat (np,2,nrr) [C order == Fortran (nrr,3,np) / (nrr,2,np)], srdist (np,nrr)."""
        tt = _f64(ttime)
        rs = np.ascontiguousarray(raystat, dtype=np.int32)
        sd = None if srdist is None else _f64(srdist)
        nrr = tt.shape[-1]
        if nrays_total is None:
            nrays_total = int((rs[:, 0, :] == 1).sum())
        self._nrr = nrr
        self._L.mct_session_set_data.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _check(self._L.mct_session_set_data(self._h, nrr, sigdep, nrays_total, tt.ctypes.data, rs.ctypes.data, _ptr(sd)))

    def likelihood(self, pending=False, rays=None, snoise0=None, snoise1=None, want_arrays=False):
        """surf_likelihood's tail on the resident maps: dict(like, misfit, unweighted_misfit[, phase_time, sigma])."""
        out = np.zeros(3)
        n0 = None if snoise0 is None else _f64(snoise0)
        n1 = None if snoise1 is None else _f64(snoise1)
        pt = np.zeros((len(self.freqs), self._nrr)) if want_arrays else None
        sg = np.zeros((len(self.freqs), self._nrr)) if want_arrays else None
        self._L.mct_session_likelihood.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_void_p]
        if rays is None:
            rp, ro, nr = None, None, 0
        else:
            rp, ro, nr = _f64(rays[0]), np.ascontiguousarray(rays[1], dtype=np.int64), int(rays[2])
        rc = _check(self._L.mct_session_likelihood(self._h, 1 if pending else 0, _ptr(rp), _ptr(ro), nr, _ptr(n0), _ptr(n1),
                                                   out.ctypes.data, _ptr(pt), _ptr(sg)), allow=(MCT_E_ZERO_NOISE,))
        return dict(like=out[0], misfit=out[1], unweighted_misfit=out[2], phase_time=pt, sigma=sg, rc=rc)

    def set_fm2d(self, src, rcv, opts):
        """sources (nsrc,2), receivers (nrc,2) as (x, y); opts = fm2d_opts(...)"""
        src, rcv = _f64(src), _f64(rcv)
        sx, sz, rx, rz = _f64(src[:, 0].copy()), _f64(src[:, 1].copy()), _f64(rcv[:, 0].copy()), _f64(rcv[:, 1].copy())
        vp = C.c_void_p
        self._L.mct_session_set_fm2d.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.POINTER(mct_fm2d_opts)]
        _check(self._L.mct_session_set_fm2d(self._h, sx.ctypes.data, sz.ctypes.data, len(src), rx.ctypes.data, rz.ctypes.data, len(rcv),
                                            C.byref(opts)))

    def likelihood_fm2d(self, pending=False, snoise0=None, snoise1=None, want_arrays=False):
        """surf_likelihood with curved rays (fast marching) on the resident maps: dict(like, misfit, unweighted_misfit[, ...])."""
        out = np.zeros(3)
        n0 = None if snoise0 is None else _f64(snoise0)
        n1 = None if snoise1 is None else _f64(snoise1)
        pt = np.zeros((len(self.freqs), self._nrr)) if want_arrays else None
        sg = np.zeros((len(self.freqs), self._nrr)) if want_arrays else None
        vp = C.c_void_p
        self._L.mct_session_likelihood_fm2d.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        rc = _check(self._L.mct_session_likelihood_fm2d(self._h, 1 if pending else 0, _ptr(n0), _ptr(n1), out.ctypes.data, _ptr(pt), _ptr(sg)),
                    allow=(MCT_E_ZERO_NOISE,))
        return dict(like=out[0], misfit=out[1], unweighted_misfit=out[2], phase_time=pt, sigma=sg, rc=rc)

    def stat_accumulate(self):
        self._L.mct_session_stat_accumulate.argtypes = [C.c_void_p]
        _check(self._L.mct_session_stat_accumulate(self._h))

    def stat_get(self):
        g = self.grid
        a = [np.zeros(g.shape) for _ in range(4)]
        n = C.c_int64(0)
        self._L.mct_session_stat_get.argtypes = [C.c_void_p] * 5 + [C.POINTER(C.c_int64)]
        _check(self._L.mct_session_stat_get(self._h, *[x.ctypes.data for x in a], C.byref(n)))
        return a, n.value

    def get_maps(self):
        g = self.grid
        pv, gv = np.zeros((g.nx, g.ny, self.nout)), np.zeros((g.nx, g.ny, self.nout))
        ie = np.zeros((g.nx, g.ny), np.int32)
        _check(self._L.mct_session_get_maps(self._h, pv.ctypes.data, gv.ctypes.data, ie.ctypes.data))
        return pv, gv, ie


def surf_misfit(time, ttime, raystat, sigdep=0, nrays_total=None, snoise0=None, snoise1=None, srdist=None):
    """mct_surf_misfit: the Gaussian misfit sums of likelihood_surf.F90:356-404 for HOST ray times (e.g. fm2d's).
    time (np,nrr), ttime (np,3,nrr), raystat (np,2,nrr), srdist (np,nrr) [C order == Fortran (nrr,..,np)]."""
    L = lib()
    t, tt = _f64(time), _f64(ttime)
    rs = np.ascontiguousarray(raystat, dtype=np.int32)
    np_, nrr = t.shape
    if nrays_total is None:
        nrays_total = int((rs[:, 0, :] == 1).sum())
    out = np.zeros(3)
    sg = np.zeros((np_, nrr))
    n0 = None if snoise0 is None else _f64(snoise0)
    n1 = None if snoise1 is None else _f64(snoise1)
    sd = None if srdist is None else _f64(srdist)
    L.mct_surf_misfit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
    rc = _check(L.mct_surf_misfit(t.ctypes.data, nrr, np_, sigdep, nrays_total, tt.ctypes.data, rs.ctypes.data, _ptr(n0), _ptr(n1),
                                  _ptr(sd), out.ctypes.data, sg.ctypes.data), allow=(MCT_E_ZERO_NOISE,))
    return dict(like=out[0], misfit=out[1], unweighted_misfit=out[2], sigma=sg, rc=rc)


# ---- multi-GPU data plane (mct_comm_*, include/mctomo_b200.h) -------------------------------------------------------
MCT_COMM_ID_BYTES = 128


def slab_bounds(nx: int, nranks: int, rank: int):
    """mct_slab_bounds: (ix0, ix1, per), 1-based inclusive; needs no device."""
    a, b, p = C.c_int(0), C.c_int(0), C.c_int(0)
    L = lib()
    L.mct_slab_bounds.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _check(L.mct_slab_bounds(nx, nranks, rank, C.byref(a), C.byref(b), C.byref(p)))
    return a.value, b.value, p.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(MCT_COMM_ID_BYTES)
    _check(lib().mct_comm_unique_id(buf))
    return buf.raw


def comm_init(id128: bytes, rank: int, nranks: int):
    L = lib()
    L.mct_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int]
    assert len(id128) == MCT_COMM_ID_BYTES
    _check(L.mct_comm_init(id128, rank, nranks))


def comm_set_mode(mode: int):
    """0: contiguous x-slabs; 1: balanced (every n-th distinct column of the sorted list)."""
    L = lib()
    L.mct_comm_set_mode.argtypes = [C.c_int]
    _check(L.mct_comm_set_mode(mode))


def comm_destroy():
    _check(lib().mct_comm_destroy())


def comm_info():
    r, n, v = C.c_int(0), C.c_int(1), C.c_int(0)
    rc = lib().mct_comm_info(C.byref(r), C.byref(n), C.byref(v))
    return dict(rank=r.value, nranks=n.value, nccl_version=v.value, active=(rc == 0))


def comm_init_torch(dist, device):
    """Bootstrap for a torch.distributed host program: rank 0 draws the id, one broadcast moves the 128 bytes."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.zeros(MCT_COMM_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)


def allgather_inplace(d_buf, bytes_per_rank: int, stream):
    L = lib()
    L.mct_allgather_inplace.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    _check(L.mct_allgather_inplace(d_buf, bytes_per_rank, stream))


def allreduce_flags(d_flags, n: int, stream):
    L = lib()
    L.mct_allreduce_flags.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    _check(L.mct_allreduce_flags(d_flags, n, stream))


def forward_sharded_dev(grid: Grid, freqs, opts, d_vp, d_vs, d_rho, d_sites, d_pvel, d_gvel, d_ierr, d_flags, stream,
                        derive_vp_rho=True):
    """mct_forward_sharded_dev: this rank's x-slab of the resident nuclei set + in-place all-gather of the maps."""
    L = lib()
    vp = C.c_void_p
    L.mct_forward_sharded_dev.argtypes = [C.POINTER(mct_grid), C.c_int, vp, C.c_int, C.POINTER(mct_disp_opts)] + [vp] * 9
    freqs = _f64(freqs)
    return _check(L.mct_forward_sharded_dev(C.byref(grid.c()), 1 if derive_vp_rho else 0, freqs.ctypes.data, len(freqs),
                                            C.byref(opts), d_vp, d_vs, d_rho, d_sites, d_pvel, d_gvel, d_ierr, d_flags, stream),
                  allow=(MCT_E_GRT_NEEDED, MCT_E_TOO_MANY_LAYERS, MCT_E_FLUID_BELOW_TOP))


def comm_last_ms() -> float:
    v = C.c_double(0.0)
    _check(lib().mct_comm_last_ms(C.byref(v)))
    return v.value


# ---- the 2-D product and point location ---------------------------------------------------------------
def voronoi_to_grid_2d(points2, params, nx, ny, xmin, ymin, dx, dy):
    """kdtree_to_grid of the 2-D variant (mcmc2d/mcmc.f90:1469-1526): returns vp, vs, rho (nx, ny) and sites_id."""
    p2, par = _f64(points2), _f64(params)
    vp, vs, rho = (np.zeros((nx, ny)) for _ in range(3))
    sid = np.zeros((nx, ny), np.int32)
    L = lib()
    L.mct_voronoi_to_grid_2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_void_p] * 4
    _check(L.mct_voronoi_to_grid_2d(p2.ctypes.data, par.ctypes.data, len(p2), nx, ny, xmin, ymin, dx, dy, vp.ctypes.data,
                                    vs.ctypes.data, rho.ctypes.data, sid.ctypes.data))
    return vp, vs, rho, sid


def nearest_nucleus(points, queries):
    """kdtree_locate for a batch (dim 2 or 3): 1-based index of the nearest nucleus of every query point."""
    p, q = _f64(points), _f64(queries)
    out = np.zeros(len(q), np.int32)
    L = lib()
    L.mct_nearest_nucleus.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    _check(L.mct_nearest_nucleus(p.ctypes.data, p.shape[1], len(p), q.ctypes.data, len(q), out.ctypes.data))
    return out


def sites_locate(points, sites_id, grid: Grid, queries):
    """sites_locate (src/likelihood_body.F90:799-831) for a batch of 3-D points; sites_id: host (nx,ny,nz) int32."""
    p, q = _f64(points), _f64(queries)
    sid = np.ascontiguousarray(sites_id, dtype=np.int32)
    out = np.zeros(len(q), np.int32)
    L = lib()
    L.mct_sites_locate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(mct_grid), C.c_void_p, C.c_int64, C.c_void_p]
    _check(L.mct_sites_locate(p.ctypes.data, len(p), sid.ctypes.data, C.byref(grid.c()), q.ctypes.data, len(q), out.ctypes.data))
    return out
