"""Synthetic Voronoi models of the reference's own initial-model family and the BASELINE configs.

generate_model (reference src/initialise.f90:339-380): `ncells` nuclei uniform in the box,
vs_i = vsmin + (z_i - zmin)(vsmax - vsmin)/(zmax - zmin), vp = 1.73 vs (vs2vp), nucleus rho from
vp2rho = 2.35 + 0.036 (vp-3)^2 (src/utils.f90:90-123).  Box, velocity range and periods follow
examples/example1/MCTomo.inp:11-18,88-89.  Pure numpy; no GPU code here.
"""
from __future__ import annotations

import numpy as np

from .capi import Grid

# name -> (nx, ny, nz, np, ncells); SURVEY.md section 8 / BASELINE.md
CONFIGS = {
    "C1": dict(nx=101, ny=101, nz=121, np=11, ncells=300),
    "C2": dict(nx=64, ny=64, nz=40, np=20, ncells=300),
    "C3": dict(nx=256, ny=256, nz=60, np=40, ncells=1000),
    "C4": dict(nx=128, ny=128, nz=50, np=20, ncells=300),
    "C5": dict(nx=1024, ny=1024, nz=80, np=60, ncells=5000),
}


def make_grid(nx, ny, nz, waterDepth=0.0, scaling=1.0) -> Grid:
    return Grid(nx, ny, nz, -5.0, 5.0, -5.0, 5.0, 0.0, 12.0, waterDepth=waterDepth, scaling=scaling)


def periods(np_: int) -> np.ndarray:
    """np periods linearly spaced 0.5 .. 10 s, ascending."""
    return np.linspace(0.5, 10.0, np_)


def freqs(np_: int) -> np.ndarray:
    return 1.0 / periods(np_)


# examples/example1/otimes.dat:2 as shipped (six decimals: 0.333333, not 1/3); read_times_3 reads them with a
# list-directed read into real(ii10) (src/likelihood_settings.f90:332-421), i.e. the nearest double of each decimal
EXAMPLE1_FREQS_SHIPPED = (2.000000, 1.000000, 0.500000, 0.333333, 0.250000, 0.200000, 0.166667, 0.142857, 0.125000,
                          0.111111, 0.100000)


def example1_freqs() -> np.ndarray:
    """The 11 frequencies of examples/example1/otimes.dat:2 exactly as the file ships them (periods ~0.5,1,2,...,10 s)."""
    return np.array(EXAMPLE1_FREQS_SHIPPED, dtype=np.float64)


def generate_model(grid: Grid, ncells: int, seed: int, vsmin=2.0, vsmax=6.0):
    """Returns (points (ncells,3), params (ncells,3) = vp,vs,rho)."""
    rng = np.random.default_rng(seed)
    lo = np.array([grid.xmin, grid.ymin, grid.zmin])
    hi = np.array([grid.xmax, grid.ymax, grid.zmax])
    pts = rng.uniform(lo, hi, size=(ncells, 3))
    vs = vsmin + (pts[:, 2] - grid.zmin) * (vsmax - vsmin) / (grid.zmax - grid.zmin)
    vp = vs * float(np.float32(1.730))
    rho = float(np.float32(2.35)) + float(np.float32(0.036)) * (vp - 3) ** 2
    return np.ascontiguousarray(pts), np.ascontiguousarray(np.stack([vp, vs, rho], axis=1))


def config(name: str, seed_offset: int = 0):
    """(grid, points, params, freqs) of a BASELINE config; RNG = default_rng(1000 + id + seed_offset)."""
    c = CONFIGS[name]
    grid = make_grid(c["nx"], c["ny"], c["nz"])
    cid = int(name[1:])
    pts, par = generate_model(grid, c["ncells"], 1000 + cid + seed_offset)
    f = example1_freqs() if name == "C1" else freqs(c["np"])
    return grid, pts, par, f
