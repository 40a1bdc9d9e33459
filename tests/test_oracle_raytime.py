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


def test_misfit_restatement_against_numpy():
    """orc_surf_misfit (likelihood_surf.F90:356-404) against a direct numpy statement of the same sums (sequential
    order kept with math.fsum-free python loops), both noise models, rays without data, zero-noise condition."""
    import math
    rng = np.random.default_rng(8)
    np_, nrr = 4, 15
    time = rng.uniform(1, 5, (np_, nrr))
    ttime = np.zeros((np_, 3, nrr))
    ttime[:, 0] = rng.uniform(1, 5, (np_, nrr))
    ttime[:, 1] = rng.uniform(0.1, 0.4, (np_, nrr))
    raystat = np.zeros((np_, 2, nrr), np.int32)
    raystat[:, 0] = rng.uniform(size=(np_, nrr)) < 0.7
    srdist = rng.uniform(1, 9, (np_, nrr))
    sn0, sn1 = rng.uniform(0.01, 0.1, np_), rng.uniform(0.01, 0.1, np_)
    for sigdep in (0, 1):
        got = orc.surf_misfit(time, ttime, raystat, sigdep=sigdep, snoise0=sn0, snoise1=sn1, srdist=srdist, math_mode=orc.LIBM)
        like = mis = unw = 0.0
        sig = np.ones((np_, nrr))
        for i in range(np_):
            for r in range(nrr):
                if raystat[i, 0, r] == 1:
                    sg = sn0[i] * srdist[i, r] + sn1[i] if sigdep else ttime[i, 1, r]
                    sig[i, r] = sg
                    d2 = (time[i, r] - ttime[i, 0, r]) ** 2
                    like += d2 / (2 * sg * sg)
                    mis += d2 / (sg * sg)
                    unw += d2
        slog = 0.0
        for v in sig.reshape(-1):
            slog += math.log(v)
        nr = int((raystat[:, 0] == 1).sum())
        like = like + slog + float(np.float32(nr) / np.float32(2.0)) * math.log(float(np.float32(6.283185)))
        assert np.array_equal(got["sigma"], sig)
        assert got["like"] == like and got["misfit"] == mis and got["unweighted_misfit"] == unw and got["rc"] == 0
        port = orc.surf_misfit(time, ttime, raystat, sigdep=sigdep, snoise0=sn0, snoise1=sn1, srdist=srdist)
        assert abs(port["like"] - like) <= 1e-12 * abs(like) and port["misfit"] == mis
