"""Known answers for the oracle's CalGroupTime / GetVelocity (reference src/likelihood_surf.F90:454-521)."""
import numpy as np

import oracle_lib as orc
from mctomo_b200 import synth


def test_constant_map_gives_length_over_velocity():
    grid = synth.make_grid(12, 9, 5)
    vel = np.full((grid.nx, grid.ny, 3), 2.5)
    line = np.array([[-4.0, -3.0], [0.0, 0.0], [4.0, 3.0]])
    off = np.concatenate([[0], np.cumsum([3, 0, 1] * 3)]).astype(np.int64)     # per period: a ray, an empty one, a 1-point one
    pts = np.concatenate([np.concatenate([line, line[:1]]) for _ in range(3)])
    t = orc.cal_group_time(vel, grid, pts, off, 3)
    assert t.shape == (3, 3)
    assert np.allclose(t[:, 0], 10.0 / 2.5, rtol=1e-14) and (t[:, 1:] == 0).all()


def test_bilinear_interpolation_reproduces_a_bilinear_field():
    """v = a + b x + c y + d x y is reproduced exactly by the 4-node stencil, so the trapezoid rule on a fine straight
    ray converges to the line integral of 1/v; period slices are independent."""
    grid = synth.make_grid(21, 17, 5)
    x = grid.xmin + np.arange(grid.nx) * grid.dx
    y = grid.ymin + np.arange(grid.ny) * grid.dy
    X, Y = np.meshgrid(x, y, indexing="ij")
    vel = np.stack([3.0 + 0.1 * X + 0.05 * Y + 0.01 * X * Y, np.full_like(X, 4.0)], axis=-1)
    n = 4001
    s = np.linspace(0.0, 1.0, n)[:, None]
    a, b = np.array([-4.5, -2.0]), np.array([3.5, 4.0])
    ray = a + s * (b - a)
    off = np.array([0, n, 2 * n], np.int64)
    t = orc.cal_group_time(vel, grid, np.concatenate([ray, ray]), off, 1)
    v = lambda p: 3.0 + 0.1 * p[:, 0] + 0.05 * p[:, 1] + 0.01 * p[:, 0] * p[:, 1]
    L = np.linalg.norm(b - a)
    fine = a + np.linspace(0.0, 1.0, 200001)[:, None] * (b - a)
    exact = np.trapezoid(1.0 / v(fine), dx=L / 200000)
    assert abs(t[0, 0] - exact) < 1e-7 and abs(t[1, 0] - L / 4.0) < 1e-12
