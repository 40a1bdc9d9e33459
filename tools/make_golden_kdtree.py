#!/usr/bin/env python
"""Generates tests/golden/kdtree2_ref.npz from the REFERENCE'S OWN kd-tree object.

Runs in the build container only (needs /root/reference and oracle/_ref/kdtree2_ref, see
oracle/build_ref.sh).  Each case stores the nuclei, the query points and what
kdtree2_n_nearest(nn=1) of utils/libutils.a:kdtree2.o returned (1-based index, squared distance).
Cases:
  random300   : 300 uniform nuclei in example1's box, 4000 uniform queries (some outside the box)
  vertices2   : the 2476 accepted nuclei of plot/run_time_vertices_2.txt (a real MCTomo model;
                only the nuclei are read, write_vertices format read_write.f90:197-203), queries on a
                coarse grid of its box x[0,5] y[0,14] z[0,2]
  lattice_ties: 5x5x5 integer lattice + 10 exact duplicates, shuffled; queries on the half-integer
                grid, where up to 8 nuclei are exactly equidistant -- pins the tie-breaking
  tiny13/14   : the leaf/first-split boundary of bucket_size = 12
  collinear   : 20 distinct nuclei sharing x = y (a line along z) + 20 scattered: the inherited, over-estimated box
                picks a dimension in which a node's >= 14 points are all equal, the split leaves one child empty and
                kdtree2 keeps the one-child node (kdtree2.f90:818-826), which its search scans as a terminal (:1388)
  plane       : 60 nuclei on the plane x = 0.3 + 5 off it (same situation one dimension up), and 30 duplicates of
                ONE point among 50 others (coincident in all dimensions but fewer than the whole range)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc  # noqa: E402


def main():
    assert orc.have_ref_binary(), "run oracle/build_ref.sh first"
    rng = np.random.default_rng(20261017)
    cases = {}
    pts = rng.uniform([-5, -5, 0], [5, 5, 12], (300, 3))
    cases["random300"] = (pts, rng.uniform([-6, -6, -1], [6, 6, 13], (4000, 3)))
    txt = np.loadtxt("/root/reference/plot/run_time_vertices_2.txt", skiprows=1)
    nuc = np.ascontiguousarray(txt[:, :3])
    gx, gy, gz = np.meshgrid(np.linspace(0, 5, 11), np.linspace(0, 14, 29), np.linspace(0, 2, 9), indexing="ij")
    cases["vertices2"] = (nuc, np.stack([gx, gy, gz], -1).reshape(-1, 3))
    lat = np.stack(np.meshgrid(np.arange(5.), np.arange(5.), np.arange(5.), indexing="ij"), -1).reshape(-1, 3)
    lat = rng.permutation(np.concatenate([lat, lat[:10]]))
    hx = np.stack(np.meshgrid(*[np.arange(0, 4.01, 0.5)] * 3, indexing="ij"), -1).reshape(-1, 3)
    cases["lattice_ties"] = (lat, hx)
    for n in (13, 14):
        cases[f"tiny{n}"] = (rng.uniform(0, 1, (n, 3)), rng.uniform(-0.2, 1.2, (500, 3)))
    line = np.stack([np.full(20, 0.25), np.full(20, 0.25), np.linspace(0.0, 1.0, 20)], -1)
    cases["collinear"] = (rng.permutation(np.concatenate([line, rng.uniform(0, 1, (20, 3))])), rng.uniform(-0.2, 1.2, (1500, 3)))
    plane = np.concatenate([np.column_stack([np.full(60, 0.3), rng.uniform(0, 1, (60, 2))]), rng.uniform(0, 1, (5, 3))])
    cases["plane"] = (rng.permutation(plane), rng.uniform(-0.2, 1.2, (1500, 3)))
    dup = np.concatenate([np.tile(np.array([[0.4, 0.6, 0.5]]), (12, 1)), rng.uniform(0, 1, (50, 3))])
    cases["dup12"] = (rng.permutation(dup), rng.uniform(-0.2, 1.2, (1500, 3)))
    # the 2-D product (mcmc2d/mcmc.f90:1469-1556): kdtree2 with dim = 2 -- random nuclei on a grid of query nodes, and a
    # lattice with half-integer queries (exact ties) plus duplicates
    p2 = rng.uniform([-5, -5], [5, 5], (180, 2))
    g2 = np.stack(np.meshgrid(np.linspace(-5, 5, 41), np.linspace(-5, 5, 37), indexing="ij"), -1).reshape(-1, 2)
    cases["grid2d"] = (p2, g2)
    lat2 = np.stack(np.meshgrid(np.arange(6.), np.arange(6.), indexing="ij"), -1).reshape(-1, 2)
    lat2 = rng.permutation(np.concatenate([lat2, lat2[:7]]))
    cases["lattice2d"] = (lat2, np.stack(np.meshgrid(np.arange(0, 5.01, .5), np.arange(0, 5.01, .5), indexing="ij"), -1).reshape(-1, 2))
    out = {}
    for name, (p, q) in cases.items():
        idx, dis = orc.ref_kd_nearest(p, q)
        out[f"{name}_points"] = p
        out[f"{name}_queries"] = q.astype(np.float64)
        out[f"{name}_idx"] = idx
        out[f"{name}_dis"] = dis
        print(name, len(p), "nuclei", len(q), "queries")
    path = os.path.join(ROOT, "tests", "golden", "kdtree2_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
