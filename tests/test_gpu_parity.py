"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Gates (north_star): cell assignment bit-exact with kdtree2 (ties included); phase and group velocity
within 1e-5 km/s of surfdisp96; identical ierr / mode counts.  Against the oracle in PORTABLE math mode
(same sin/cos/exp as the device) the stronger statement is tested: every output and both work counters
are bit-identical.
"""
import os

import numpy as np
import pytest

import oracle_lib as orc
from mctomo_b200 import synth
from mctomo_b200.capi import Grid, disp_opts

pytestmark = pytest.mark.gpu

TOL = 1e-5  # km/s, north_star


def _empty_model(grid):
    return (np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32))


def _both_k1(mct, pts, par, grid, box, pm=None, init=None):
    """GPU (all four kernel shapes: tree walk per node, round-1 column scan, box lists with the float64 walk, box lists
    with the float32 walk = production) and oracle."""
    outs = []
    for mode in (1, 2, 3, 0):
        mct.set_k1_mode(mode)
        vp, vs, rho, sid = _empty_model(grid) if init is None else [a.copy() for a in init]
        mct.kdtree_to_grid(pts, par, grid, box, vp, vs, rho, sid, pm=pm)
        outs.append((vp, vs, rho, sid))
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert np.array_equal(x, y), f"the nearest-nucleus kernels disagree in {(x != y).sum()} nodes"
    vp, vs, rho, sid = _empty_model(grid) if init is None else [a.copy() for a in init]
    orc.kdtree_to_grid(pts, par, grid, box, vp, vs, rho, sid, pm=pm)
    return [outs[3], (vp, vs, rho, sid)]


def _assert_k1_equal(a, b):
    for x, y, name in zip(a, b, ("vp", "vs", "rho", "sites_id")):
        assert np.array_equal(x, y), f"{name} differs in {(x != y).sum()} nodes"


@pytest.mark.parametrize("ncells,seed", [(13, 1), (25, 2), (300, 3), (1000, 4)])
def test_k1_full_box_bit_exact(mct, ncells, seed):
    grid = synth.make_grid(33, 29, 41)
    pts, par = synth.generate_model(grid, ncells, seed)
    g, o = _both_k1(mct, pts, par, grid, grid.full_box())
    _assert_k1_equal(g, o)
    assert g[3].min() >= 1 and g[3].max() <= ncells


def test_k1_sub_box_leaves_outside_untouched(mct):
    grid = synth.make_grid(40, 36, 30)
    pts, par = synth.generate_model(grid, 200, 7)
    rng = np.random.default_rng(0)
    init = (rng.uniform(1, 2, grid.shape), rng.uniform(1, 2, grid.shape), rng.uniform(1, 2, grid.shape),
            rng.integers(1, 200, grid.shape).astype(np.int32))
    box = np.array([-1.3, -2.2, 3.1, 2.4, 0.7, 9.9])
    g, o = _both_k1(mct, pts, par, grid, box, init=init)
    _assert_k1_equal(g, o)
    w = mct.box_window(grid, box)
    assert np.array_equal(w, orc.box_window(grid, box))
    mask = np.ones(grid.shape, bool)
    mask[w[0] - 1:w[1], w[2] - 1:w[3], w[4] - 1:w[5]] = False
    assert np.array_equal(g[0][mask], init[0][mask]) and np.array_equal(g[3][mask], init[3][mask])


def test_k1_pm_mode(mct):
    """value move: only nodes carrying the old (vp,vs) of the perturbed cell are reassigned (mcmc_loc2.f90:2055)."""
    grid = synth.make_grid(30, 30, 24)
    pts, par = synth.generate_model(grid, 150, 11)
    base = _empty_model(grid)
    orc.kdtree_to_grid(pts, par, grid, grid.full_box(), *base)
    cell = 17
    pm = par[cell].copy()
    par2 = par.copy()
    par2[cell] = [pm[0] * 1.1, pm[1] * 1.1, pm[2]]
    g, o = _both_k1(mct, pts, par2, grid, grid.full_box(), pm=pm, init=base)
    _assert_k1_equal(g, o)
    assert (g[1] != base[1]).sum() == (base[3] == cell + 1).sum() > 0


def test_k1_ties_lattice_and_duplicates(mct):
    """Nuclei on a lattice whose nodes fall exactly half-way: the winner is kdtree2's traversal order."""
    nuc = np.stack(np.meshgrid(np.arange(5.), np.arange(5.), np.arange(5.), indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(5)
    nuc = rng.permutation(np.concatenate([nuc, nuc[:10]]))  # + exact duplicates
    par = np.stack([np.arange(len(nuc)) + 1.0, np.arange(len(nuc)) + 0.5, np.ones(len(nuc))], 1)
    grid = Grid(9, 9, 9, 0.0, 4.0, 0.0, 4.0, 0.0, 4.0)  # spacing 0.5: every other node is equidistant
    g, o = _both_k1(mct, nuc, par, grid, grid.full_box())
    _assert_k1_equal(g, o)


def test_k1_many_nuclei_and_thin_window(mct):
    """5000 nuclei (config C5's count: the per-column candidate stage overflows for some columns and falls back to
    the tree walk) and a window one node thick."""
    grid = synth.make_grid(20, 18, 80)
    pts, par = synth.generate_model(grid, 5000, 31)
    g, o = _both_k1(mct, pts, par, grid, grid.full_box())
    _assert_k1_equal(g, o)
    box = np.array([-5.0, -5.0, 6.0, 5.0, 5.0, 6.05])
    g, o = _both_k1(mct, pts, par, grid, box)
    _assert_k1_equal(g, o)


@pytest.mark.parametrize("shape", [(9, 7, 41), (5, 33, 121), (3, 3, 1), (17, 16, 7), (6, 5, 150)])
def test_k1_odd_shapes_and_windows(mct, shape):
    """Odd nz (columns start on odd element offsets: the 16-byte stores shift by one node per column), a single plane,
    columns taller than 128 nodes (segments longer than 8), and sub-windows starting on odd z."""
    nx, ny, nz = shape
    grid = synth.make_grid(nx, ny, nz) if nz > 1 else Grid(nx, ny, 2, -5, 5, -5, 5, 0, 12)
    pts, par = synth.generate_model(grid, 90, 5 + nz)
    g, o = _both_k1(mct, pts, par, grid, grid.cover_box())
    _assert_k1_equal(g, o)
    rng = np.random.default_rng(nz)
    init = (rng.uniform(1, 2, grid.shape), rng.uniform(1, 2, grid.shape), rng.uniform(1, 2, grid.shape),
            rng.integers(1, 90, grid.shape).astype(np.int32))
    for box in ([-3.1, -2.0, 1.3, 2.2, 3.9, 7.7], [-5.0, -5.0, 0.31, 5.0, 5.0, 0.32], [0.1, 0.1, 5.0, 0.2, 0.2, 11.9]):
        g, o = _both_k1(mct, pts, par, grid, np.array(box), init=init)
        _assert_k1_equal(g, o)


def test_k1_batch_with_odd_model_stride(mct):
    """A batch whose models have an odd number of nodes: model b's arrays start 8 bytes off a 16-byte boundary for odd b."""
    from mctomo_b200 import capi
    grid = synth.make_grid(7, 5, 9)   # 315 nodes
    freqs = synth.freqs(3)
    models = [synth.generate_model(grid, 20 + 3 * b, 50 + b) for b in range(3)]
    pts, par, off = capi.pack_models(models)
    r = mct.forward_eval_batch(pts, par, off, grid, freqs, disp_opts(), want_model=True)
    for b in range(3):
        o = orc.forward_eval(*models[b], grid, freqs)
        assert np.array_equal(r["sites_id"][b], o["sites_id"]) and np.array_equal(r["vs"][b], o["vs"])
        assert np.array_equal(r["rho"][b], o["rho"]) and np.array_equal(r["pvel"][b], o["pvel"])


def test_k1_one_child_tree_nodes(mct):
    """Nuclei sharing a coordinate (a line, a plane, a cluster of duplicates): kdtree2's split leaves one child empty,
    the node keeps its single child and is scanned as a terminal (kdtree2.f90:818-826,1388).  The oracle is pinned on
    these sets against kdtree2.o (fixtures collinear / plane / dup12); the library must accept them and agree."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kdtree2_ref.npz"))
    grid = Grid(21, 19, 23, -0.1, 1.1, -0.1, 1.1, -0.1, 1.1)
    for case in ("collinear", "plane", "dup12"):
        nuc = np.ascontiguousarray(g[f"{case}_points"])
        par = np.stack([np.arange(len(nuc)) + 1.0, np.arange(len(nuc)) + 0.5, np.ones(len(nuc))], 1)
        a, o = _both_k1(mct, nuc, par, grid, grid.full_box())
        _assert_k1_equal(a, o)


def test_2d_product_and_point_location(mct):
    """kdtree_to_grid of the 2-D variant (mcmc2d/mcmc.f90:1469-1526), kdtree_locate (:1528-1551) and
    sites_locate (src/likelihood_body.F90:799-831).  2-D cells against fixtures made with the reference's kdtree2.o run
    with dim = 2 (random nuclei on grid nodes; a lattice with exact ties and duplicates)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kdtree2_ref.npz"))
    p2 = np.ascontiguousarray(g["grid2d_points"])
    par = np.stack([np.arange(len(p2)) + 1.0, np.arange(len(p2)) + 0.5, np.ones(len(p2))], 1)
    vp, vs, rho, sid = mct.voronoi_to_grid_2d(p2, par, 41, 37, -5.0, -5.0, 0.25, 10.0 / 36)
    # the fixture's queries are linspace nodes; the library's are xmin + (i-1)*dx: compare through the oracle on ITS nodes
    q = np.stack(np.meshgrid(-5.0 + np.arange(41) * 0.25, -5.0 + np.arange(37) * (10.0 / 36), indexing="ij"), -1).reshape(-1, 2)
    ref, _ = orc.kd_nearest(p2, q)
    assert np.array_equal(sid.reshape(-1), ref) and np.array_equal(vs.reshape(-1), par[ref - 1, 1])
    assert (sid.reshape(-1) == g["grid2d_idx"]).mean() > 0.99           # same cells as the fixture wherever the nodes coincide bitwise
    for case in ("grid2d", "lattice2d"):                                 # kdtree_locate on the fixtures' own query points
        got = mct.nearest_nucleus(g[f"{case}_points"], g[f"{case}_queries"])
        assert np.array_equal(got, g[f"{case}_idx"]), case
    got3 = mct.nearest_nucleus(g["lattice_ties_points"], g["lattice_ties_queries"])
    assert np.array_equal(got3, g["lattice_ties_idx"])
    # the 2-D product's surface-wave likelihood (mcmc2d/likelihood_surf_phase.F90:150-190): one dispersion curve per x
    # node of a (depth, x) profile = the 3-D call with ny = 1 and no check_model
    nxp, nzp = 37, 33
    p2d = np.column_stack([np.random.default_rng(4).uniform(-5, 5, 60), np.random.default_rng(5).uniform(0, 12, 60)])  # (x, depth)
    vsn = 2.0 + p2d[:, 1] / 3.0
    par2 = np.column_stack([1.73 * vsn, vsn, 1.74 * (1.73 * vsn) ** 0.25])
    vp2, vs2, rho2, sid2 = mct.voronoi_to_grid_2d(p2d, par2, nxp, nzp, -5.0, 0.0, 10.0 / (nxp - 1), 12.0 / (nzp - 1))
    prof = Grid(nxp, 1, nzp, -5.0, 5.0, 0.0, 1.0, 0.0, 12.0)
    fr = synth.freqs(6)
    pv, gv, ie, inval, rc = mct.surf_dispersion(vp2.reshape(nxp, 1, nzp), vs2.reshape(nxp, 1, nzp), rho2.reshape(nxp, 1, nzp), prof,
                                                (1, nxp, 1, 1), fr, disp_opts(phaseGroup=1), check=False)
    po, go, io, _, _ = orc.surf_dispersion(vp2.reshape(nxp, 1, nzp), vs2.reshape(nxp, 1, nzp), rho2.reshape(nxp, 1, nzp), prof,
                                           (1, nxp, 1, 1), fr, phaseGroup=1)
    assert np.array_equal(pv, po) and np.array_equal(gv, go) and np.array_equal(ie, io) and (ie <= 1).all()
    # sites_locate
    grid = synth.make_grid(14, 13, 11)
    pts, par3 = synth.generate_model(grid, 30, 8)
    m = _empty_model(grid)
    orc.kdtree_to_grid(pts, par3, grid, grid.cover_box(), *m)
    qq = np.random.default_rng(3).uniform([-5.5, -5.5, -0.5], [5.5, 5.5, 12.5], (5000, 3))
    assert np.array_equal(mct.sites_locate(pts, m[3], grid, qq), orc.sites_locate(pts, m[3], grid, qq))


def test_k1_degenerate_nuclei_reported(mct):
    grid = synth.make_grid(8, 8, 8)
    pts = np.tile(np.array([[0.1, 0.2, 3.0]]), (40, 1))
    par = np.ones((40, 3))
    with pytest.raises(mct.MctError) as e:
        mct.kdtree_to_grid(pts, par, grid, grid.full_box(), *_empty_model(grid))
    assert e.value.code == mct.MCT_E_DEGENERATE_NUCLEI


def _model(grid, ncells, seed, math_mode=orc.PORTABLE):
    pts, par = synth.generate_model(grid, ncells, seed)
    vp, vs, rho, sid = _empty_model(grid)
    orc.kdtree_to_grid(pts, par, grid, grid.full_box(), vp, vs, rho, sid)
    vp, rho = orc.vs2vp_rho(vs, math_mode)
    return vp, vs, rho


def _check_disp(mct, grid, vp, vs, rho, window, freqs, raylov, pg, nmodes, variant="likelihood"):
    """Runs the dispersion block with EVERY kernel shape (one thread per column; 2..32 lanes per column; 2 and 4
    warps per column); all must agree bit for bit with each other and with the oracle, counters included."""
    opts = disp_opts(raylov=raylov, phaseGroup=pg, nmodes=nmodes, variant=variant)
    res = []
    for mode, lanes in ((1, 0), (2, 256), (2, 128), (2, 64), (2, 32), (2, 16), (2, 8), (2, 4), (2, 2)):
        mct.set_k2_mode(mode)
        mct.set_k2_lanes(lanes)
        mct.reset_stats()
        res.append(mct.surf_dispersion(vp, vs, rho, grid, window, freqs, opts) + (mct.stats(),))
    mct.set_k2_mode(0)
    mct.set_k2_lanes(0)
    for r in res[1:]:
        for a, b in zip(res[0][:3], r[:3]):
            assert np.array_equal(a, b), "thread-per-column and lane-cooperative kernels disagree"
        assert res[0][5]["n_dltar"] == r[5]["n_dltar"] and res[0][5]["n_layer_steps"] == r[5]["n_layer_steps"]
    pv, gv, ie, inval, rc, st = res[1]
    kw = dict(raylov=raylov, phaseGroup=pg, nmodes=nmodes, layer_eps=opts.layer_eps, water_thresh=opts.water_thresh,
              preset=opts.preset)
    po, go, io, cnt, nun = orc.surf_dispersion(vp, vs, rho, grid, window, freqs, math_mode=orc.PORTABLE, **kw)
    assert inval == 0
    assert np.array_equal(ie, io), "ierr differs"
    assert np.array_equal(pv, po), f"phase velocity not bit-identical: max |d| = {np.abs(pv - po).max()}"
    assert np.array_equal(gv, go), f"group velocity not bit-identical: max |d| = {np.abs(gv - go).max()}"
    assert st["n_dltar"] == cnt[0] and st["n_layer_steps"] == cnt[1], (st, cnt)
    # the faithful (libm) restatement: the north_star tolerance
    pl, gl, il, _, _ = orc.surf_dispersion(vp, vs, rho, grid, window, freqs, math_mode=orc.LIBM, **kw)
    assert np.array_equal(ie, il)
    assert np.abs(pv - pl).max() <= TOL
    assert np.abs(gv - gl).max() <= TOL
    return pv, gv, ie


@pytest.mark.parametrize("raylov,pg", [(1, 0), (1, 1), (0, 0), (0, 1)])
def test_k2_fundamental(mct, raylov, pg):
    grid = synth.make_grid(12, 10, 40)
    vp, vs, rho = _model(grid, 120, 21)
    pv, gv, ie = _check_disp(mct, grid, vp, vs, rho, (1, 12, 1, 10), synth.freqs(20), raylov, pg, 0)
    assert (ie == 0).all() and (pv < 7).all() and (pv > 1).all()


@pytest.mark.parametrize("raylov,pg,nmodes", [(1, 0, 2), (1, 1, 2), (0, 1, 3), (1, 0, 1)])
def test_k2_overtones_mode_counts(mct, raylov, pg, nmodes):
    grid = synth.make_grid(9, 8, 30)
    vp, vs, rho = _model(grid, 80, 22)
    pv, gv, ie = _check_disp(mct, grid, vp, vs, rho, (1, 9, 1, 8), synth.freqs(16), raylov, pg, nmodes)
    # "mode counts": found/not-found pattern already compared bit-wise; make sure overtones exist at all
    if nmodes > 1:
        assert (pv.reshape(9, 8, nmodes, 16)[:, :, 1, :4] > 0).any()


def test_k2_window_water_scaling(mct):
    grid = synth.make_grid(14, 13, 25, waterDepth=0.8, scaling=1.0)
    vp, vs, rho = _model(grid, 90, 23)
    _check_disp(mct, grid, vp, vs, rho, (3, 9, 2, 12), synth.freqs(10), 1, 1, 0)
    grid2 = Grid(14, 13, 25, -5.0, 5.0, -5.0, 5.0, 0.0, 12.0, waterDepth=0.0, scaling=2.0)
    vp, vs, rho = _model(grid2, 90, 24)
    _check_disp(mct, grid2, vp, vs, rho, (1, 14, 5, 13), synth.freqs(10), 1, 0, 0)


def test_k2_modelling_variant(mct):
    """forward_modelling.f90's constants: EPS=1e-5, no check_model, preset 1000."""
    grid = synth.make_grid(10, 9, 20)
    vp, vs, rho = _model(grid, 60, 25)
    _check_disp(mct, grid, vp, vs, rho, (1, 10, 1, 9), synth.freqs(8), 1, 1, 0, variant="modelling")


def test_k2_check_model_and_lvl_columns(mct):
    grid = synth.make_grid(8, 7, 20)
    vp, vs, rho = _model(grid, 50, 26)
    freqs = synth.freqs(6)
    opts = disp_opts()
    bad = vs.copy()
    bad[5, 3, 10] = bad[5, 3, 0] * 0.5  # deeper vs below the top vs: check_model rejects the whole model
    pv, gv, ie, inval, rc = mct.surf_dispersion(vp, bad, rho, grid, (1, 8, 1, 7), freqs, opts)
    assert inval == 1 == orc.check_model(bad, grid)
    # without check_model (program modelling has none) the same column takes the GRT branch: ierr = 2
    pv, gv, ie, inval, rc = mct.surf_dispersion(vp, bad, rho, grid, (1, 8, 1, 7), freqs, opts, check=False)
    po, go, io, cnt, nun = orc.surf_dispersion(vp, bad, rho, grid, (1, 8, 1, 7), freqs)
    assert np.array_equal(ie, io) and np.array_equal(pv, po)
    assert rc == mct.MCT_E_GRT_NEEDED or nun == 0


def test_surfmodes_batch(mct):
    rng = np.random.default_rng(3)
    cols, offs = [], [0]
    for c in range(37):
        n = int(rng.integers(1, 9))
        vs = np.sort(rng.uniform(1.5, 4.5, n))
        th = rng.uniform(0.3, 3.0, n)
        th[-1] = 0
        vp = 1.73 * vs
        cols.append(np.stack([th, vp, vs, 1.74 * vp ** 0.25], 1))
        offs.append(offs[-1] + n)
    a = np.concatenate(cols)
    freqs = synth.freqs(12)
    for raylov, pg, nm in [(1, 1, 0), (0, 0, 2)]:
        opts = disp_opts(raylov=raylov, phaseGroup=pg, nmodes=nm)
        ph, gr, ie, rc = mct.surfmodes_batch(a[:, 0], a[:, 1], a[:, 2], a[:, 3], offs, freqs, opts)
        for c in range(37):
            s = slice(offs[c], offs[c + 1])
            rc0, p0, g0, e0, _ = orc.surfmodes(a[s, 0], a[s, 1], a[s, 2], a[s, 3], freqs, raylov, pg, nm)
            assert rc0 == 0 and e0 == ie[c]
            assert np.array_equal(ph[c], p0) and np.array_equal(gr[c], g0)


def test_forward_eval_fused(mct):
    grid, pts, par, freqs = synth.config("C2")
    grid = synth.make_grid(24, 20, 40)
    pts, par = synth.generate_model(grid, 300, 1002)
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    r = mct.forward_eval(pts, par, grid, freqs, opts, want_model=True)
    o = orc.forward_eval(pts, par, grid, freqs)
    assert r["model_invalid"] == 0 == o["model_invalid"]
    for k in ("sites_id", "vs", "vp", "rho", "ierr", "pvel"):
        assert np.array_equal(r[k], o[k]), k
    ol = orc.forward_eval(pts, par, grid, freqs, math_mode=orc.LIBM)
    assert np.abs(r["pvel"] - ol["pvel"]).max() <= TOL
    assert np.abs(r["rho"] - ol["rho"]).max() <= 1e-15 * 4


def test_vs2vp_rho(mct):
    vs = np.random.default_rng(9).uniform(0.5, 6.0, 100003)
    vp, rho = mct.vs2vp_rho(vs)
    vo, ro = orc.vs2vp_rho(vs, orc.PORTABLE)
    assert np.array_equal(vp, vo) and np.array_equal(rho, ro)
    vl, rl = orc.vs2vp_rho(vs, orc.LIBM)
    assert np.array_equal(vp, vl) and np.abs(rho - rl).max() <= np.spacing(rl.max())


@pytest.mark.parametrize("emax", [8, 60, 390, 1000])
def test_division_selftest(mct, emax):
    """The kernel's shared-reciprocal division must be bit-identical to IEEE division (what the oracle does)."""
    tested, bad = mct.selftest_division(emax)
    assert tested > 10 ** 9 and bad == 0, f"{bad} of {tested} quotients differ from IEEE division"


@pytest.mark.parametrize("raylov", [1, 0])
def test_full_size_c3_sampled_against_oracle(mct, raylov):
    """BASELINE config C3 at full size (256x256x60, 40 periods, phase+group, fundamental + 1st overtone):
    every column is solved on the GPU; 300 randomly chosen columns are re-solved by the oracle and must be
    bit-identical; size-independent properties are checked on all 65536 columns."""
    grid, pts, par, freqs = synth.config("C3")
    opts = disp_opts(raylov=raylov, phaseGroup=1, nmodes=2)
    r = mct.forward_eval(pts, par, grid, freqs, opts, want_model=True)
    assert r["model_invalid"] == 0 and r["rc"] == 0
    np_ = len(freqs)
    pv = r["pvel"].reshape(grid.nx, grid.ny, 2, np_)
    gv = r["gvel"].reshape(grid.nx, grid.ny, 2, np_)
    ie = r["ierr"]
    assert set(np.unique(ie)) <= {0, 1}
    # fundamental found everywhere, velocities inside the model's range, no mode crossing where the overtone exists
    assert (pv[:, :, 0] > 1.5).all() and (pv[:, :, 0] < 6.1).all()
    found = pv[:, :, 1] > 0
    assert found.any() and (pv[:, :, 1][found] > pv[:, :, 0][found]).all()
    # once an overtone is lost at some period it stays lost (ift, surfdisp96.f:566,688-695)
    assert (np.diff(found.astype(np.int8), axis=-1) <= 0).all()
    # cell map: brute-force check of a sample of nodes (ties have measure zero for random nuclei)
    rng = np.random.default_rng(0)
    ii = rng.integers(0, grid.nx, 2000); jj = rng.integers(0, grid.ny, 2000); kk = rng.integers(0, grid.nz, 2000)
    q = np.stack([grid.xmin + ii * grid.dx, grid.ymin + jj * grid.dy, grid.zmin + kk * grid.dz], 1)
    d = ((q[:, None, :] - pts[None]) ** 2).sum(-1)
    assert np.array_equal(r["sites_id"][ii, jj, kk], d.argmin(1) + 1)
    # sampled columns against the oracle
    for i, j in zip(rng.integers(0, grid.nx, 300), rng.integers(0, grid.ny, 300)):
        n, (th, al, be, rk) = orc.convert_column(r["vp"][i, j], r["vs"][i, j], r["rho"][i, j], grid.dz)
        rc, p0, g0, e0, _ = orc.surfmodes(th, al, be, rk, freqs, raylov, 1, 2)
        assert rc == 0 and e0 == ie[i, j]
        assert np.array_equal(p0, r["pvel"][i, j]) and np.array_equal(g0, r["gvel"][i, j])


def test_full_size_c2x32_sampled_against_oracle(mct):
    """The very workload bench.py times: BASELINE config C2 (64x64x40, 20 periods, Rayleigh phase, 300 nuclei) as a
    batch of 32 models through mct_forward_eval_batch.  Model 0 and model 17 are compared with the oracle in full
    (PORTABLE: bit-identical incl. both work counters; LIBM: float32-identical count reported, phase within 1e-5 km/s);
    ten random columns of every other model are re-solved by the oracle."""
    from mctomo_b200 import capi
    grid = synth.make_grid(64, 64, 40)
    freqs = synth.freqs(20)
    models = [synth.generate_model(grid, 300, 1000 + 2 + b) for b in range(32)]   # bench.py's seeds on rank 0
    pts, par, off = capi.pack_models(models)
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    mct.reset_stats()
    r = mct.forward_eval_batch(pts, par, off, grid, freqs, opts, want_model=True)
    st = mct.stats()
    assert r["pvel"].shape == (32, 64, 64, 20) and not r["ierr"].any() and not np.any(r["model_invalid"])
    assert (r["pvel"] > 1.5).all() and (r["pvel"] < 6.1).all() and (np.diff(r["pvel"], axis=-1) > -1e-3).all()
    assert st["n_columns"] == 32 * 4096 and 0 < st["n_columns_solved"] <= st["n_columns"]
    nd = nl = 0
    for b in (0, 17):
        o = orc.forward_eval(*models[b], grid, freqs)
        assert np.array_equal(r["sites_id"][b], o["sites_id"]) and np.array_equal(r["vs"][b], o["vs"])
        assert np.array_equal(r["pvel"][b], o["pvel"]) and np.array_equal(r["ierr"][b], o["ierr"])
        ol = orc.forward_eval(*models[b], grid, freqs, math_mode=orc.LIBM)
        diff = r["pvel"][b] != ol["pvel"]
        print(f"model {b}: {int(diff.sum())} of {diff.size} phase velocities differ from the libm-mode oracle "
              f"(max {np.abs(r['pvel'][b] - ol['pvel']).max():.3g} km/s)")
        assert diff.mean() <= 1e-4 and np.abs(r["pvel"][b] - ol["pvel"]).max() <= TOL
    rng = np.random.default_rng(32)
    for b in range(32):
        for i, j in zip(rng.integers(0, 64, 10), rng.integers(0, 64, 10)):
            n, (th, al, be, rk) = orc.convert_column(r["vp"][b, i, j], r["vs"][b, i, j], r["rho"][b, i, j], grid.dz)
            rc, p0, g0, e0, _ = orc.surfmodes(th, al, be, rk, freqs, 1, 0, 0)
            assert rc == 0 and e0 == 0 and np.array_equal(p0, r["pvel"][b, i, j])


@pytest.mark.parametrize("ncells", [25, 100, 300])
def test_full_size_c1_example1_as_shipped(mct, ncells):
    """BASELINE config C1: example1's grid (101x101x121, examples/example1/MCTomo.inp:20-22) with the 11 frequencies of
    examples/example1/otimes.dat:2 AS SHIPPED (0.333333, 0.166667, ... not exact reciprocals), Rayleigh phase, 25 / 100 /
    300 cells (the prior's bounds).  Every one of the 10 201 columns is compared with the oracle: outputs, ierr and both
    work counters; with few cells most columns are duplicates of each other, which the library folds (exactly)."""
    grid = synth.make_grid(101, 101, 121)
    freqs = synth.example1_freqs()
    assert freqs[3] == 0.333333 and len(freqs) == 11
    pts, par = synth.generate_model(grid, ncells, 1001 + ncells)
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    o = orc.forward_eval(pts, par, grid, freqs)
    res = {}
    for dd in (True, False):
        mct.set_dedup(dd)
        mct.reset_stats()
        res[dd] = (mct.forward_eval(pts, par, grid, freqs, opts, want_model=True), mct.stats(), mct.last_launch())
    mct.set_dedup(True)
    for dd in (True, False):
        r, st, li = res[dd]
        assert np.array_equal(r["sites_id"], o["sites_id"])
        assert np.array_equal(r["pvel"], o["pvel"]) and np.array_equal(r["ierr"], o["ierr"]), f"dedup={dd}"
        assert st["n_dltar"] == o["counters"][0] and st["n_layer_steps"] == o["counters"][1] and st["n_columns"] == 10201
    on, off_ = res[True][1], res[False][1]
    assert off_["n_columns_solved"] == 10201 and off_["n_dltar_executed"] == off_["n_dltar"]
    vs = o["vs"].astype(np.float32).reshape(-1, grid.nz)
    distinct = len(np.unique(vs, axis=0))
    assert on["n_columns_solved"] == res[True][2]["columns_solved"]
    assert abs(on["n_columns_solved"] - distinct) <= 2   # (the key is the layered stack; float32 cell sequences agree with it
    #                                                       unless two nuclei's vs collide within float32 rounding)
    assert on["n_dltar_executed"] < on["n_dltar"] or distinct == 10201
    print(f"C1 {ncells} cells: {distinct} distinct columns of 10201, kernel {res[True][2]['kernel']}")
    ol = orc.forward_eval(pts, par, grid, freqs, math_mode=orc.LIBM)
    diff = res[True][0]["pvel"] != ol["pvel"]
    print(f"  {int(diff.sum())} of {diff.size} phase velocities differ from the libm-mode oracle")
    assert diff.mean() <= 1e-4 and np.abs(res[True][0]["pvel"] - ol["pvel"]).max() <= TOL


def test_dedup_forced_duplicates_all_kernel_shapes(mct):
    """A model of 5 cells on a 40x40 grid: ~20 distinct columns among 1600.  Every kernel shape must reproduce the
    oracle for EVERY column (the duplicates' outputs are copied from their representative), with represented counters
    equal to the oracle's and executed counters equal to the distinct columns' share; phase+group, two modes; a second
    model in the batch that check_model rejects must stay untouched."""
    import torch
    grid = synth.make_grid(40, 40, 30)
    freqs = synth.freqs(7)
    pts, par = synth.generate_model(grid, 5, 77)
    o = orc.forward_eval(pts, par, grid, freqs, phaseGroup=1, nmodes=2)
    opts = disp_opts(raylov=1, phaseGroup=1, nmodes=2)
    vs32 = o["vs"].astype(np.float32).reshape(-1, grid.nz)
    distinct = len(np.unique(vs32, axis=0))
    assert distinct < 800
    for lanes in (0, 2, 8, 32, 128):
        mct.set_k2_lanes(lanes)
        for mode in ((0, 1) if lanes == 0 else (0,)):
            mct.set_k2_mode(mode, -1)
            mct.reset_stats()
            r = mct.forward_eval(pts, par, grid, freqs, opts)
            st = mct.stats()
            assert np.array_equal(r["pvel"], o["pvel"]) and np.array_equal(r["gvel"], o["gvel"]) and np.array_equal(r["ierr"], o["ierr"])
            assert st["n_dltar"] == o["counters"][0] and st["n_layer_steps"] == o["counters"][1]
            assert abs(st["n_columns_solved"] - distinct) <= 1 and st["n_columns"] == 1600
    mct.set_k2_lanes(0)
    mct.set_k2_mode(0, -1)
    # batch of two: the second model is invalid (fast cell on top) -> its outputs stay as they were
    from mctomo_b200 import capi
    bad = par.copy()
    bad[int(np.argmin(pts[:, 2])), 1] = 9.0
    p2, a2, off = capi.pack_models([(pts, par), (pts, bad)])
    out = {"pvel": np.full((2, 40, 40, 14), -3.0), "gvel": np.full((2, 40, 40, 14), -3.0), "ierr": np.full((2, 40, 40), -3, np.int32)}
    r = mct.forward_eval_batch(p2, a2, off, grid, freqs, opts, out=out)
    assert list(r["model_invalid"]) == [0, 1]
    assert np.array_equal(r["pvel"][0], o["pvel"])


def test_full_size_c5_sampled_against_oracle(mct):
    """BASELINE config C5 at full size (1024x1024x80, 60 periods = the NP limit, 5000 nuclei, Rayleigh phase): all
    1 048 576 columns are solved on one GPU; 200 random columns are re-solved by the oracle (bit-identical), the cell
    map is checked by brute force on a sample of nodes, and the maps must be complete and inside the model's range."""
    grid, pts, par, freqs = synth.config("C5")
    assert (grid.nx, grid.ny, grid.nz, len(freqs), len(pts)) == (1024, 1024, 80, 60, 5000)
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    r = mct.forward_eval(pts, par, grid, freqs, opts, want_model=True)
    assert r["model_invalid"] == 0 and r["rc"] == 0
    pv, ie = r["pvel"], r["ierr"]
    assert pv.shape == (1024, 1024, 60) and not ie.any()
    assert (pv > 1.5).all() and (pv < 6.1).all()
    assert (np.diff(pv, axis=-1) > -1e-3).all()          # normal dispersion for depth-increasing velocities
    rng = np.random.default_rng(5)
    ii = rng.integers(0, grid.nx, 1500); jj = rng.integers(0, grid.ny, 1500); kk = rng.integers(0, grid.nz, 1500)
    q = np.stack([grid.xmin + ii * grid.dx, grid.ymin + jj * grid.dy, grid.zmin + kk * grid.dz], 1)
    d = ((q[:, None, :] - pts[None]) ** 2).sum(-1)
    assert np.array_equal(r["sites_id"][ii, jj, kk], d.argmin(1) + 1)
    for i, j in zip(rng.integers(0, grid.nx, 200), rng.integers(0, grid.ny, 200)):
        n, (th, al, be, rk) = orc.convert_column(r["vp"][i, j], r["vs"][i, j], r["rho"][i, j], grid.dz)
        rc, p0, g0, e0, _ = orc.surfmodes(th, al, be, rk, freqs, 1, 0, 0)
        assert rc == 0 and e0 == 0
        assert np.array_equal(p0, pv[i, j])


def test_full_size_c4_chains_per_gpu(mct):
    """BASELINE config C4, one GPU's share at full size: 8 chains x 128x128x50, 20 periods, ncells ~ U{25..300} per
    chain, in ONE batched call; 25 random columns of every chain are re-solved by the oracle (bit-identical)."""
    grid = synth.make_grid(128, 128, 50)
    freqs = synth.freqs(20)
    rng = np.random.default_rng(1004)
    models = [synth.generate_model(grid, int(rng.integers(25, 301)), 1004 + c) for c in range(8)]
    pts, par, off = mct.pack_models(models)
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    r = mct.forward_eval_batch(pts, par, off, grid, freqs, opts, want_model=True)
    assert r["rc"] == 0 and not r["model_invalid"].any() and not r["ierr"].any()
    for b in range(8):
        for i, j in zip(rng.integers(0, grid.nx, 25), rng.integers(0, grid.ny, 25)):
            n, (th, al, be, rk) = orc.convert_column(r["vp"][b, i, j], r["vs"][b, i, j], r["rho"][b, i, j], grid.dz)
            rc, p0, g0, e0, _ = orc.surfmodes(th, al, be, rk, freqs, 1, 0, 0)
            assert rc == 0 and e0 == 0 and np.array_equal(p0, r["pvel"][b, i, j])


def test_forward_eval_batch_of_chains(mct):
    """Several independent models (chains) with different numbers of nuclei in ONE call (config C4's shape)."""
    grid = synth.make_grid(20, 18, 30)
    freqs = synth.freqs(12)
    models = [synth.generate_model(grid, n, 500 + n) for n in (25, 300, 77, 150)]
    opts = disp_opts(raylov=1, phaseGroup=1, nmodes=0)
    pts, par, off = mct.pack_models(models)
    for k1 in (0, 1):
        mct.set_k1_mode(k1)
        r = mct.forward_eval_batch(pts, par, off, grid, freqs, opts, want_model=True)
        assert r["rc"] == 0 and not r["model_invalid"].any()
        for b, (p, q) in enumerate(models):
            o = orc.forward_eval(p, q, grid, freqs, phaseGroup=1)
            for k in ("sites_id", "vs", "vp", "rho"):
                assert np.array_equal(r[k][b], o[k]), (b, k)
            assert np.array_equal(r["ierr"][b], o["ierr"])
            assert np.array_equal(r["pvel"][b], o["pvel"]) and np.array_equal(r["gvel"][b], o["gvel"])
    mct.set_k1_mode(0)
    # one model of the batch is invalid (a deeper vs below the top vs): only that model is rejected
    bad = [(p.copy(), q.copy()) for p, q in models]
    top = np.argmin(bad[2][0][:, 2])
    bad[2][1][top, 1] = 9.0  # the shallowest nucleus gets the highest vs
    pts, par, off = mct.pack_models(bad)
    r = mct.forward_eval_batch(pts, par, off, grid, freqs, opts)
    assert list(r["model_invalid"]) == [0, 0, 1, 0]
    o = orc.forward_eval(bad[3][0], bad[3][1], grid, freqs, phaseGroup=1)
    assert np.array_equal(r["pvel"][3], o["pvel"])


def test_edge_cases_single_cell_halfspace_and_limits(mct):
    """ncells = 1 (every column is a bare half-space, mmax = 1), the NP = 60 period limit, bad arguments."""
    grid = synth.make_grid(5, 4, 12)
    pts = np.array([[0.3, -0.2, 5.0]])
    par = np.array([[1.73 * 3.0, 3.0, 2.5]])
    freqs = 1.0 / np.linspace(0.5, 30.0, 60)
    for raylov in (1, 0):
        opts = disp_opts(raylov=raylov, phaseGroup=1, nmodes=0)
        r = mct.forward_eval(pts, par, grid, freqs, opts, want_model=True)
        o = orc.forward_eval(pts, par, grid, freqs, raylov=raylov, phaseGroup=1)
        assert (r["sites_id"] == 1).all()
        assert np.array_equal(r["ierr"], o["ierr"]) and np.array_equal(r["pvel"], o["pvel"]) and np.array_equal(r["gvel"], o["gvel"])
        assert (r["ierr"] == (0 if raylov == 1 else 1)).all()  # no Love wave in a half-space
    with pytest.raises(mct.MctError) as e:
        mct.forward_eval(pts, par, grid, 1.0 / np.linspace(0.5, 30.0, 61), disp_opts())
    assert e.value.code == mct.MCT_E_INVALID_ARG
    with pytest.raises(mct.MctError):
        mct.surf_dispersion(np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), grid, (0, 3, 1, 2), freqs[:5], disp_opts())


def test_water_layer_love_and_rayleigh_overtones(mct):
    grid = synth.make_grid(9, 7, 30, waterDepth=1.2)
    vp, vs, rho = _model(grid, 70, 77)
    _check_disp(mct, grid, vp, vs, rho, (1, 9, 1, 7), synth.freqs(12), 0, 1, 2)
    _check_disp(mct, grid, vp, vs, rho, (2, 8, 1, 7), synth.freqs(12), 1, 1, 2)


def test_failing_columns_ierr_and_zeroed_group(mct):
    """Very long periods push the root to betmx: surfdisp96 gives up (ierr = 1, cg zeroed from the failing
    period on, surfdisp96.f:333-376).  The GPU must fail in exactly the same places."""
    grid = synth.make_grid(8, 6, 20)
    vp, vs, rho = _model(grid, 40, 78)
    freqs = 1.0 / np.geomspace(0.5, 3000.0, 24)
    pv, gv, ie = _check_disp(mct, grid, vp, vs, rho, (1, 8, 1, 6), freqs, 0, 1, 0)
    assert (ie == 1).any() and (gv == 0.0).any() and (pv == 100.0).any()


def test_assemble_vel_dev(mct):
    import torch
    nx, ny, np_ = 9, 7, 5
    rng = np.random.default_rng(4)
    st = torch.cuda.Stream()
    for win in [(1, 9, 1, 7), (3, 6, 2, 5), (1, 4, 3, 7), (5, 9, 1, 2)]:
        wx, wy = win[1] - win[0] + 1, win[3] - win[2] + 1
        pvel = rng.uniform(1, 5, (wx, wy, np_))
        vel0 = rng.uniform(10, 20, (nx + 2, ny + 2, np_))
        ref = vel0.copy()
        orc.assemble_vel(pvel, np_, nx, ny, win, ref)
        d_p = torch.from_numpy(pvel).cuda()
        d_v = torch.from_numpy(vel0.copy()).cuda()
        with torch.cuda.stream(st):
            mct.assemble_vel_dev(d_p.data_ptr(), np_, nx, ny, win, d_v.data_ptr(), st.cuda_stream)
        st.synchronize()
        assert np.array_equal(d_v.cpu().numpy(), ref), win


@pytest.mark.parametrize("seed,raylov,pg,nm", [(1, 1, 0, 0), (2, 1, 1, 2), (3, 0, 1, 0), (4, 0, 0, 3), (5, 1, 1, 0)])
def test_fuzz_random_layer_stacks(mct, seed, raylov, pg, nm):
    """1500 random pre-layered columns per case through mct_surfmodes_batch: thin and very thick layers (the
    exp-skipping branches p >= 16, exa >= 60), vp/vs from 1.45 to 2.6, strong and tiny velocity contrasts,
    water layers, periods from 0.1 s to 80 s.  Every column must match the oracle bit for bit, including the
    columns the reference routes to the GRT branch (ierr = 2) and the ones whose search fails (ierr = 1)."""
    rng = np.random.default_rng(seed)
    ncol = 1500
    cols, offs = [], [0]
    for c in range(ncol):
        n = int(rng.integers(1, 14))
        water = rng.random() < 0.15 and n >= 2
        vs = np.sort(rng.uniform(0.3, 5.0, n))
        if rng.random() < 0.2:                      # near-identical neighbours
            vs[1:] = vs[:-1] + rng.uniform(1e-7, 1e-3, n - 1)
        if rng.random() < 0.1 and n >= 3:           # a low-velocity layer somewhere: GRT branch in the reference
            k = int(rng.integers(1, n - 1))
            vs[k] = vs[0] * rng.uniform(0.5, 0.99)
        ratio = rng.uniform(1.45, 2.6, n)
        vp = vs * ratio
        rho = rng.uniform(1.0, 3.5, n)
        th = rng.choice([rng.uniform(1e-3, 0.05), rng.uniform(0.1, 3.0), rng.uniform(5.0, 60.0)], size=n)
        th[-1] = 0.0
        if water:
            vs[0], vp[0], rho[0], th[0] = 0.0, 1.5, 1.0, rng.uniform(0.05, 4.0)
        cols.append(np.stack([th, vp, vs, rho], 1))
        offs.append(offs[-1] + n)
    a = np.concatenate(cols)
    periods = np.sort(rng.choice(np.geomspace(0.1, 80.0, 40), 14, replace=False))
    freqs = 1.0 / periods
    opts = disp_opts(raylov=raylov, phaseGroup=pg, nmodes=nm)
    for mode, lanes in ((1, 0), (2, 128), (2, 32), (2, 8), (2, 2)):
        mct.set_k2_mode(mode)
        mct.set_k2_lanes(lanes)
        ph, gr, ie, rc = mct.surfmodes_batch(a[:, 0], a[:, 1], a[:, 2], a[:, 3], offs, freqs, opts)
        seen = {0: 0, 1: 0, 2: 0}
        for c in range(ncol):
            s = slice(offs[c], offs[c + 1])
            rc0, p0, g0, e0, _ = orc.surfmodes(a[s, 0], a[s, 1], a[s, 2], a[s, 3], freqs, raylov, pg, nm)
            if rc0 == 2:
                assert ie[c] == 2, c
                seen[2] += 1
                continue
            assert rc0 == 0 and e0 == ie[c], (c, rc0, e0, ie[c])
            assert np.array_equal(ph[c], p0) and np.array_equal(gr[c], g0), (c, np.abs(ph[c] - p0).max())
            seen[e0] += 1
        assert seen[0] + seen[1] > 1000 and seen[0] > 100 and seen[2] > 20
    mct.set_k2_mode(0)
    mct.set_k2_lanes(0)


def test_windowed_maps_and_window_scoped_check(mct):
    """mct_vs2vp_rho_window touches only the window; check_scope = 1 looks only at the window's columns."""
    grid = synth.make_grid(14, 12, 20)
    vp, vs, rho = _model(grid, 50, 90)
    vp2 = np.full(grid.shape, -1.0)
    rho2 = np.full(grid.shape, -1.0)
    w = (3, 9, 2, 7, 4, 15)
    mct.vs2vp_rho_window(vs, vp2, rho2, grid, w)
    inside = np.zeros(grid.shape, bool)
    inside[w[0] - 1:w[1], w[2] - 1:w[3], w[4] - 1:w[5]] = True
    assert np.array_equal(vp2[inside], vp[inside]) and np.array_equal(rho2[inside], rho[inside])
    assert (vp2[~inside] == -1).all() and (rho2[~inside] == -1).all()
    # an invalid column OUTSIDE the window: scope 0 (reference) rejects, scope 1 does not and solves the window
    bad = vs.copy()
    bad[12, 10, 8] = 0.5 * bad[12, 10, 0]
    freqs = synth.freqs(6)
    win = (3, 9, 2, 7)
    r0 = mct.surf_dispersion(vp, bad, rho, grid, win, freqs, disp_opts(check_scope=0))
    r1 = mct.surf_dispersion(vp, bad, rho, grid, win, freqs, disp_opts(check_scope=1))
    assert r0[3] == 1 and r1[3] == 0
    po, go, io, cnt, nun = orc.surf_dispersion(vp, bad, rho, grid, win, freqs)
    assert np.array_equal(r1[0], po) and np.array_equal(r1[2], io)
    # ... and an invalid column INSIDE the window is caught by both
    bad2 = vs.copy()
    bad2[4, 4, 8] = 0.5 * bad2[4, 4, 0]
    assert mct.surf_dispersion(vp, bad2, rho, grid, win, freqs, disp_opts(check_scope=1))[3] == 1


def test_sample_postprocessor_accumulation(mct):
    """program sample's loop (src/sample.f90:481-487): regrid every kept sample over the full grid and accumulate
    sum(vs), sum(vs^2), sum(vp), sum(vp^2) -- here entirely on the device; compared with the oracle's regrids
    accumulated by numpy in the same order (bit-identical sums)."""
    import torch
    grid = synth.make_grid(22, 19, 25)
    n = grid.nx * grid.ny * grid.nz
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream()
    d = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.zeros(n, dtype=torch.int32, device=dev)]
    acc = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(4)]
    ref = [np.zeros(grid.shape) for _ in range(4)]
    for k in range(7):
        pts, par = synth.generate_model(grid, 30 + 11 * k, 300 + k)
        with torch.cuda.stream(st):
            mct.voronoi_to_grid_dev(pts, par, grid, grid.cover_box(), *[t.data_ptr() for t in d], st.cuda_stream)
            mct.accumulate_stats_dev(d[1].data_ptr(), d[0].data_ptr(), *[t.data_ptr() for t in acc], n, st.cuda_stream)
        vp, vs, rho, sid = _empty_model(grid)
        orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
        ref[0] = ref[0] + vs
        ref[1] = ref[1] + vs * vs
        ref[2] = ref[2] + vp
        ref[3] = ref[3] + vp * vp
    st.synchronize()
    for a, r in zip(acc, ref):
        assert np.array_equal(a.cpu().numpy().reshape(grid.shape), r)


def test_status_codes_and_degenerate_inputs(mct):
    """Conditions the reference answers with `stop`, array overruns or empty loops: too many layers (ierr = 3), a fluid
    layer below the top (ierr = 4), boxes outside the grid (nothing is touched), no nuclei (argument error)."""
    freqs = synth.freqs(4)
    opts = disp_opts()
    # 230 distinct layers > NL = 200 (surfdisp96.f:57) next to an ordinary column and one with a fluid layer in the middle
    n_big = 230
    th = np.concatenate([np.full(n_big - 1, 0.05), [0.0], [1.0, 2.0, 0.0], [1.0, 1.0, 2.0, 0.0]])
    vs = np.concatenate([np.linspace(1.0, 4.5, n_big), [2.0, 3.0, 4.0], [2.0, 0.0, 3.0, 4.0]])
    vp = 1.73 * vs
    vp[n_big + 3 + 1] = 1.5
    rho = np.full_like(vs, 2.5)
    offs = [0, n_big, n_big + 3, n_big + 7]
    ph, gr, ie, rc = mct.surfmodes_batch(th, vp, vs, rho, offs, freqs, opts)
    assert list(ie) == [3, 0, 4]
    assert rc in (mct.MCT_E_TOO_MANY_LAYERS, mct.MCT_E_FLUID_BELOW_TOP)
    assert (ph[0] == opts.preset).all() and (ph[2] == opts.preset).all()
    rc0, p0, g0, e0, _ = orc.surfmodes(th[n_big:n_big + 3], vp[n_big:n_big + 3], vs[n_big:n_big + 3], rho[n_big:n_big + 3], freqs, 1, 0, 0)
    assert np.array_equal(ph[1], p0)                      # the ordinary column is solved regardless of its neighbours
    # boxes that miss the grid: the Fortran loops run zero times, nothing is written
    grid = synth.make_grid(8, 7, 10)
    pts, par = synth.generate_model(grid, 20, 3)
    base = [np.full(grid.shape, -7.0), np.full(grid.shape, -7.0), np.full(grid.shape, -7.0), np.full(grid.shape, -7, np.int32)]
    for box in ([grid.xmax + 1, grid.ymin, grid.zmin, grid.xmax + 2, grid.ymax, grid.zmax],
                [grid.xmin, grid.ymin, grid.zmax + 1, grid.xmax, grid.ymax, grid.zmax + 5],
                [grid.xmin, grid.ymin, grid.zmin, grid.xmin - 3, grid.ymax, grid.zmax]):
        arrs = [a.copy() for a in base]
        mct.kdtree_to_grid(pts, par, grid, np.array(box, float), *arrs)
        ref = [a.copy() for a in base]
        orc.kdtree_to_grid(pts, par, grid, np.array(box, float), *ref)
        for a, b in zip(arrs, ref):
            assert np.array_equal(a, b)
    with pytest.raises(mct.MctError) as e:
        mct.kdtree_to_grid(pts[:0], par[:0], grid, grid.cover_box(), *[a.copy() for a in base])
    assert e.value.code == mct.MCT_E_INVALID_ARG


def test_maximum_layer_count(mct):
    """Columns with exactly NL = 200 layers (surfdisp96.f:57) and 199: solved, bit-identical to the oracle, in the
    thread-per-column kernel and in the cooperative ones (7 passes of the layer-parallel evaluation)."""
    cols, offs = [], [0]
    for n in (200, 199, 64, 33):
        vs = np.linspace(1.2, 4.6, n) + 0.003 * np.sin(np.arange(n))
        vs = np.sort(vs)
        th = np.full(n, 0.06)
        th[-1] = 0.0
        vp = 1.8 * vs
        cols.append(np.stack([th, vp, vs, np.full(n, 2.6)], 1))
        offs.append(offs[-1] + n)
    a = np.concatenate(cols)
    freqs = 1.0 / np.array([0.3, 1.0, 4.0, 12.0])
    for raylov in (1, 0):
        opts = disp_opts(raylov=raylov, phaseGroup=1, nmodes=0)
        for mode, lanes in ((1, 0), (2, 128), (2, 32), (2, 4)):
            mct.set_k2_mode(mode)
            mct.set_k2_lanes(lanes)
            ph, gr, ie, rc = mct.surfmodes_batch(a[:, 0], a[:, 1], a[:, 2], a[:, 3], offs, freqs, opts)
            for c in range(4):
                s = slice(offs[c], offs[c + 1])
                rc0, p0, g0, e0, _ = orc.surfmodes(a[s, 0], a[s, 1], a[s, 2], a[s, 3], freqs, raylov, 1, 0)
                assert rc0 == 0 and e0 == ie[c], (raylov, lanes, c)
                assert np.array_equal(ph[c], p0) and np.array_equal(gr[c], g0), (raylov, lanes, c)
    mct.set_k2_mode(0)
    mct.set_k2_lanes(0)
