"""Pins the kd-tree oracle (oracle/kdtree2_ref.c) against the reference's own kdtree2.o:
committed fixtures generated from that object (tools/make_golden_kdtree.py) and, when the object is
present (build container), live runs through oracle/_ref/kdtree2_ref."""
import os

import numpy as np
import pytest

import oracle_lib as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kdtree2_ref.npz")
CASES = ["random300", "vertices2", "lattice_ties", "tiny13", "tiny14", "collinear", "plane", "dup12", "grid2d", "lattice2d"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_fixture(case):
    g = np.load(GOLD)
    idx, dis = orc.kd_nearest(g[f"{case}_points"], g[f"{case}_queries"])
    assert np.array_equal(idx, g[f"{case}_idx"]), "nearest index differs from kdtree2.o"
    assert np.array_equal(dis, g[f"{case}_dis"]), "squared distance not bit-identical to kdtree2.o"


def test_fixture_really_contains_ties():
    g = np.load(GOLD)
    p, q = g["lattice_ties_points"], g["lattice_ties_queries"]
    d = ((q[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    ntie = (np.isclose(d, d.min(1, keepdims=True), rtol=0, atol=0).sum(1) > 1).sum()
    assert ntie > 300
    # the winner is neither "lowest index" nor "highest index": it is traversal order
    lo = d.argmin(1) + 1
    hi = d.shape[1] - d[:, ::-1].argmin(1)
    idx = g["lattice_ties_idx"]
    assert (idx != lo).any() and (idx != hi).any()


@pytest.mark.skipif(not orc.have_ref_binary(), reason="oracle/_ref/kdtree2_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,seed", [(13, 0), (14, 1), (40, 2), (300, 3), (2500, 4)])
def test_oracle_matches_reference_live(n, seed):
    rng = np.random.default_rng(seed)
    pts = rng.uniform([-5, -5, 0], [5, 5, 12], (n, 3))
    pts[: n // 4] = np.round(pts[: n // 4], 1)  # coarse coordinates: repeated values along a dimension
    q = np.concatenate([rng.uniform([-6, -6, -1], [6, 6, 13], (3000, 3)), np.round(rng.uniform(-5, 5, (2000, 3)), 1)])
    i1, d1 = orc.kd_nearest(pts, q)
    i2, d2 = orc.ref_kd_nearest(pts, q)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_degenerate_nuclei_detected():
    """Only nuclei coincident in EVERY dimension (more than a bucket of them) make kdtree2's build recurse for ever;
    a split that merely leaves one child empty is legal (see the collinear/plane fixtures)."""
    with pytest.raises(ValueError):
        orc.kd_nearest(np.ones((40, 3)), np.zeros((1, 3)))


@pytest.mark.skipif(not orc.have_ref_binary(), reason="oracle/_ref/kdtree2_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(6))
def test_one_child_nodes_match_reference_live(seed):
    """Random mixtures of points sharing one or two coordinates (one-child nodes, kdtree2.f90:818-826)."""
    rng = np.random.default_rng(100 + seed)
    n_line, n_plane, n_free = rng.integers(14, 60, 3)
    line = np.column_stack([np.full(n_line, 0.5), np.full(n_line, rng.uniform()), rng.uniform(0, 1, n_line)])
    plane = np.column_stack([rng.uniform(0, 1, n_plane), np.full(n_plane, 0.125), rng.uniform(0, 1, n_plane)])
    pts = rng.permutation(np.concatenate([line, plane, rng.uniform(0, 1, (n_free, 3))]))
    q = np.concatenate([rng.uniform(-0.2, 1.2, (3000, 3)), np.round(rng.uniform(0, 1, (1000, 3)), 1)])
    i1, d1 = orc.kd_nearest(pts, q)
    i2, d2 = orc.ref_kd_nearest(pts, q)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_kdtree_to_grid_window_and_pm():
    from mctomo_b200 import synth
    grid = synth.make_grid(12, 11, 10)
    pts, par = synth.generate_model(grid, 50, 5)
    vp, vs, rho = (np.zeros(grid.shape) for _ in range(3))
    sid = np.zeros(grid.shape, np.int32)
    orc.kdtree_to_grid(pts, par, grid, grid.full_box(), vp, vs, rho, sid)
    # brute force agrees wherever there is no tie
    x = grid.xmin + np.arange(grid.nx) * grid.dx
    y = grid.ymin + np.arange(grid.ny) * grid.dy
    z = grid.zmin + np.arange(grid.nz) * grid.dz
    q = np.stack(np.meshgrid(x, y, z, indexing="ij"), -1).reshape(-1, 3)
    d = ((q[:, None, :] - pts[None]) ** 2).sum(-1)
    assert np.array_equal(sid.reshape(-1), d.argmin(1) + 1)
    assert np.array_equal(vs, par[sid - 1, 1])
    # sub-box leaves the outside untouched
    box = np.array([-2.0, -1.0, 2.0, 1.5, 3.0, 7.0])
    w = orc.box_window(grid, box)
    v2 = np.full(grid.shape, -1.0)
    s2 = np.full(grid.shape, -1, np.int32)
    orc.kdtree_to_grid(pts, par, grid, box, v2.copy(), v2, v2.copy(), s2)
    inside = np.zeros(grid.shape, bool)
    inside[w[0] - 1:w[1], w[2] - 1:w[3], w[4] - 1:w[5]] = True
    assert (s2[~inside] == -1).all() and np.array_equal(s2[inside], sid[inside])


def test_full_box_rounding_quirk_is_reference_behaviour():
    """floor((zmax-zmin)/dz)+1 can land on nz-1: the reference's "whole grid" box then skips the last plane
    (src/mcmc_loc2.f90:2034-2045).  Reproduced, not fixed; cover_box() addresses every node."""
    from mctomo_b200 import synth
    grid = synth.make_grid(256, 256, 60)           # config C3: 12/(12/59) = 58.999999999999993
    w = orc.box_window(grid, grid.full_box())
    assert w[5] == 59 and w[1] == 256
    assert list(orc.box_window(grid, grid.cover_box())) == [1, 256, 1, 256, 1, 60]
    grid = synth.make_grid(101, 101, 121)          # example1's grid is unaffected
    assert list(orc.box_window(grid, grid.full_box())) == [1, 101, 1, 101, 1, 121]


def test_sites_locate_restatement():
    """sites_locate (likelihood_body.F90:799-831): inside a cell's interior the eight-node shortcut and the tree agree;
    the shortcut is taken (no tree search) exactly when the eight nodes around the point carry one index."""
    from mctomo_b200 import synth
    grid = synth.make_grid(14, 13, 11)
    pts, par = synth.generate_model(grid, 30, 8)
    vp, vs, rho = (np.zeros(grid.shape) for _ in range(3))
    sid = np.zeros(grid.shape, np.int32)
    orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
    rng = np.random.default_rng(3)
    q = rng.uniform([-5.5, -5.5, -0.5], [5.5, 5.5, 12.5], (4000, 3))
    got = orc.sites_locate(pts, sid, grid, q)
    near, _ = orc.kd_nearest(pts, q)
    inside = (np.abs(q[:, :2]) <= 5).all(1) & (q[:, 2] >= 0) & (q[:, 2] <= 12)
    assert np.array_equal(got[inside], near[inside])          # convex cells: 8 equal corners => the box is inside the cell
    assert got.min() >= 1 and got.max() <= len(pts)
