"""Proposal replay: the call pattern the rjMCMC sampler issues (reference src/mcmc_loc2.f90:183-400).

Each proposal perturbs ONE Voronoi cell -- birth, death, move, value change -- then calls
    kdtree_to_grid(RTI, grid, bnd_box, model[, pm])        on the perturbed cell's bounding box (host arrays, in place)
    likelihood -> vs2vp_3d / vp2rho_3d over the grid, surf_likelihood on the box (+1 column halo)
The GPU library (host-pointer C ABI) and the oracle replay the same chain of proposals on their own copies of
the model; after every proposal all arrays must be bit-identical.  The bounding box, which the reference gets
from CGAL (src/cgal_delaunay.cpp:276-300), is computed here from the exact before/after cell maps."""
import numpy as np
import pytest

import oracle_lib as orc
from mctomo_b200 import synth
from mctomo_b200.capi import disp_opts

pytestmark = pytest.mark.gpu


def _cell_box(grid, sid_a, sid_b, ids):
    """Coordinate box of all nodes that belong to any of `ids` (1-based) before or after the move."""
    m = np.zeros(grid.shape, bool)
    for i in ids:
        m |= (sid_a == i) | (sid_b == i)
    if not m.any():
        return None
    ii, jj, kk = np.nonzero(m)
    lo = np.array([grid.xmin + ii.min() * grid.dx, grid.ymin + jj.min() * grid.dy, grid.zmin + kk.min() * grid.dz])
    hi = np.array([grid.xmin + ii.max() * grid.dx, grid.ymin + jj.max() * grid.dy, grid.zmin + kk.max() * grid.dz])
    pad = 1e-9
    return np.concatenate([lo - pad, hi + pad])


def _window(grid, box, expand=1):
    """likelihood_surf.F90:155-158,166-169"""
    ix0 = int(np.floor((box[0] - grid.xmin) / grid.dx)) + 1 - expand
    ix1 = int(np.floor((box[3] - grid.xmin) / grid.dx)) + 1 + expand
    iy0 = int(np.floor((box[1] - grid.ymin) / grid.dy)) + 1 - expand
    iy1 = int(np.floor((box[4] - grid.ymin) / grid.dy)) + 1 + expand
    return max(ix0, 1), min(ix1, grid.nx), max(iy0, 1), min(iy1, grid.ny)


def test_proposal_replay_matches_oracle(mct):
    rng = np.random.default_rng(2026)
    grid = synth.make_grid(31, 29, 41)
    freqs = synth.example1_freqs()
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    pts, par = synth.generate_model(grid, 60, 7)
    # two copies of the model state: GPU-driven and oracle-driven
    G = [np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32)]
    O = [a.copy() for a in G]
    mct.kdtree_to_grid(pts, par, grid, grid.cover_box(), *G)
    orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), *O)
    for a, b in zip(G, O):
        assert np.array_equal(a, b)
    vsmin, vsmax = 2.0, 6.0
    kinds = ["move", "value", "birth", "death", "move", "value", "move", "birth", "value", "death", "move", "move"]
    n_checked = 0
    for step, kind in enumerate(kinds):
        pts2, par2 = pts.copy(), par.copy()
        pm = None
        if kind == "move":
            i = int(rng.integers(len(pts)))
            pts2[i] = np.clip(pts[i] + rng.normal(0, 0.6, 3), [grid.xmin, grid.ymin, grid.zmin], [grid.xmax, grid.ymax, grid.zmax])
            ids = [i + 1]
        elif kind == "value":
            i = int(rng.integers(len(pts)))
            pm = par[i].copy()                      # the OLD (vp, vs, rho) of the cell: mcmc_loc2.f90:395
            vs_new = pm[1] * (1 + 0.02 * rng.normal())
            par2[i] = [vs_new * float(np.float32(1.73)), vs_new, pm[2]]
            ids = [i + 1]
        elif kind == "birth":
            p = rng.uniform([grid.xmin, grid.ymin, grid.zmin], [grid.xmax, grid.ymax, grid.zmax])
            vs_new = vsmin + (p[2] - grid.zmin) * (vsmax - vsmin) / (grid.zmax - grid.zmin)
            pts2 = np.vstack([pts, p])
            par2 = np.vstack([par, [vs_new * float(np.float32(1.73)), vs_new, 2.5]])
            ids = [len(pts2)]
        else:  # death: remove cell i; indices above i shift down (mcmc_loc2.f90:276-278) -> regrid the dead cell's box
            i = int(rng.integers(len(pts)))
            pts2 = np.delete(pts, i, 0)
            par2 = np.delete(par, i, 0)
            ids = [i + 1]
        # exact before/after cell maps give the box CGAL would return
        full_new = [np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32)]
        orc.kdtree_to_grid(pts2, par2, grid, grid.cover_box(), *full_new)
        if kind == "death":
            box = _cell_box(grid, O[3], O[3], ids)
            for M in (G, O):                        # where(sites_id > iremove) sites_id -= 1
                M[3][M[3] > ids[0]] -= 1
        else:
            box = _cell_box(grid, O[3], full_new[3], ids)
        if box is None:
            pts, par = pts2, par2
            continue
        mct.kdtree_to_grid(pts2, par2, grid, box, *G, pm=pm)
        orc.kdtree_to_grid(pts2, par2, grid, box, *O, pm=pm)
        for a, b, name in zip(G, O, ("vp", "vs", "rho", "sites_id")):
            assert np.array_equal(a, b), f"step {step} ({kind}): {name} differs"
        if kind != "value":  # the sub-box update must reproduce the full regrid (consistency of the box)
            assert np.array_equal(O[3], full_new[3]), f"step {step} ({kind}): box did not cover the change"
        # likelihood (datatype 2): property maps over the grid, dispersion on the box + halo
        vpg, rhog = mct.vs2vp_rho(G[1])
        vpo, rhoo = orc.vs2vp_rho(O[1])
        assert np.array_equal(vpg, vpo) and np.array_equal(rhog, rhoo)
        win = _window(grid, box)
        pv, gv, ie, inval, rc = mct.surf_dispersion(vpg, G[1], rhog, grid, win, freqs, opts)
        assert inval == orc.check_model(O[1], grid)
        if not inval:
            po, go, io, cnt, nun = orc.surf_dispersion(vpo, O[1], rhoo, grid, win, freqs)
            assert np.array_equal(ie, io) and np.array_equal(pv, po), f"step {step} ({kind}): dispersion differs"
            n_checked += 1
        pts, par = pts2, par2
    assert n_checked >= 6
