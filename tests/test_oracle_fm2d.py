"""The oracle's restatement of the 2-D fast-marching solver (oracle/fm2d_ref.c) against physics.

PARITY UNPINNED (no Fortran compiler here, no golden travel times in the reference): the checks are analytic -- a
homogeneous medium (t = d / v) and a constant vertical gradient (t = acosh(1 + g^2 d^2 / (2 v1 v2)) / g) -- at the
accuracy fast marching has on example1's 101 x 101 grid, plus the properties the scheme guarantees: refinement and the
mixed-order stencils reduce the error, every receiver time is positive and no smaller than d / vmax."""
import numpy as np
import pytest

import oracle_lib as orc

NX = NY = 101
X0 = Y0 = -5.0
DX = 0.1
SRC = np.array([[0.03, 0.07], [-3.2, 2.5], [4.9, -4.9], [-5.0, 5.0]])  # inside, inside, near a corner, ON a corner
RNG = np.random.default_rng(0)
RCV = np.vstack([RNG.uniform(-4.9, 4.9, (9, 2)), [[0.05, 0.09]]])    # the last one within a cell of source 1


def dist():
    return np.sqrt(((SRC[:, None, :] - RCV[None, :, :]) ** 2).sum(-1))


@pytest.mark.parametrize("fom,asgr,tol_mean,tol_max", [(0, 0, 0.03, 0.08), (0, 1, 0.02, 0.04), (1, 0, 0.015, 0.06), (1, 1, 0.005, 0.02)])
def test_homogeneous_medium(fom, asgr, tol_mean, tol_max):
    vel = np.full((NX + 2, NY + 2), 3.0)
    err, tt, field, cnt = orc.fm2d_times(SRC, RCV, np.ones((4, 10), np.int32), vel, X0, Y0, DX, DX, fom=fom, asgr=asgr, want_field=True)
    assert err == 0 and cnt[0] >= 4 * NX * NY
    rel = np.abs(tt * 3.0 - dist()) / dist()
    assert rel.mean() < tol_mean and rel.max() < tol_max, (rel.mean(), rel.max())
    assert (field >= 0).all() and (tt > 0).all()


def test_refinement_and_mixed_order_reduce_the_error():
    vel = np.full((NX + 2, NY + 2), 3.0)
    e = {}
    for fom in (0, 1):
        for asgr in (0, 1):
            _, tt, _, _ = orc.fm2d_times(SRC[:2], RCV[:9], np.ones((2, 9), np.int32), vel, X0, Y0, DX, DX, fom=fom, asgr=asgr)
            d = dist()[:2, :9]
            e[fom, asgr] = (np.abs(tt * 3.0 - d) / d).mean()
    assert e[1, 1] < e[1, 0] < e[0, 0] and e[1, 1] < e[0, 1] < e[0, 0]


def test_constant_gradient_medium():
    g = 0.2
    yy = Y0 + (np.arange(NY + 2) - 1) * DX
    vel = np.tile(2 + g * (yy + 5), (NX + 2, 1))
    err, tt, _, _ = orc.fm2d_times(SRC[:3], RCV[:9], np.ones((3, 9), np.int32), vel, X0, Y0, DX, DX)
    v1, v2 = 2 + g * (SRC[:3, 1] + 5), 2 + g * (RCV[:9, 1] + 5)
    d2 = ((SRC[:3, None, :] - RCV[None, :9, :]) ** 2).sum(-1)
    ta = np.arccosh(1 + g * g * d2 / (2 * v1[:, None] * v2[None, :])) / g
    assert err == 0 and (np.abs(tt - ta) / ta).max() < 0.02
    assert (tt >= np.sqrt(d2) / vel.max() - 1e-12).all()


def test_sources_without_data_are_skipped_and_receivers_without_data_untouched():
    vel = np.full((NX + 2, NY + 2), 2.5)
    srs = np.ones((4, 10), np.int32)
    srs[2, :] = 0      # no valid ray: the source is skipped (it is not the first)
    srs[1, 3] = 0
    err, tt, _, cnt = orc.fm2d_times(SRC, RCV, srs, vel, X0, Y0, DX, DX)
    assert err == 0 and (tt[2] == -1.0).all() and tt[1, 3] == -1.0 and (tt[[0, 1, 3]][:, [0, 1, 2]] > 0).all()
    srs[:] = 0         # the first source is marched even without data
    err, tt, _, cnt2 = orc.fm2d_times(SRC, RCV, srs, vel, X0, Y0, DX, DX)
    assert err == 0 and (tt == -1.0).all() and 0 < cnt2[0] < cnt[0]


def test_errors_are_reported_not_fatal():
    vel = np.full((NX + 2, NY + 2), 2.5)
    err, _, _, _ = orc.fm2d_times(np.array([[9.0, 0.0]]), RCV, np.ones((1, 10), np.int32), vel, X0, Y0, DX, DX)
    assert err == 1   # source outside the model: the Fortran STOPs
    err, _, _, _ = orc.fm2d_times(SRC[:1], RCV, np.ones((1, 10), np.int32), vel, X0, Y0, DX, DX, snb=0.001)
    assert err == 2   # narrow band larger than snb*nx*ny: the Fortran overruns btg


def test_rays_are_straight_in_a_homogeneous_medium():
    vel = np.full((NX + 2, NY + 2), 3.0)
    src, rcv = SRC[:3], RCV[:9]
    err, tt, npts, pts, ln, crazy = orc.fm2d_rays(src, rcv, np.ones((3, 9), np.int32), vel, X0, Y0, DX, DX)
    assert err == 0 and crazy == 0
    d = np.sqrt(((src[:, None, :] - rcv[None, :, :]) ** 2).sum(-1)).ravel()
    for s in range(27):
        p = pts[s, :npts[s]]
        assert np.array_equal(p[0], rcv[s % 9]) and np.array_equal(p[-1], src[s // 9])     # receiver first, source last
        assert abs(ln[s] - d[s]) < 0.02 * d[s] + 0.11                                    # step = dx/2, last hop up to dx
        # every point stays close to the straight line
        a, b = rcv[s % 9], src[s // 9]
        t = (b - a) / np.linalg.norm(b - a)
        off = np.abs((p - a) @ np.array([-t[1], t[0]]))
        assert off.max() < 0.25, (s, off.max())
        seg = np.linalg.norm(np.diff(p, axis=0), axis=1)
        assert np.allclose(seg[:-1], 0.05, atol=1e-9)                                     # dpl = half the node spacing


def test_rays_in_a_gradient_medium_carry_the_travel_time():
    g = 0.2
    yy = Y0 + (np.arange(NY + 2) - 1) * DX
    vel = np.tile(2 + g * (yy + 5), (NX + 2, 1))
    src, rcv = SRC[:2], RCV[:9]
    err, tt, npts, pts, ln, crazy = orc.fm2d_rays(src, rcv, np.ones((2, 9), np.int32), vel, X0, Y0, DX, DX)
    assert err == 0 and crazy == 0
    err0, tt0, _, _ = orc.fm2d_times(src, rcv, np.ones((2, 9), np.int32), vel, X0, Y0, DX, DX)
    assert np.array_equal(tt, tt0)                                                        # the ray tracing does not touch the times
    for s in range(18):
        p = pts[s, :npts[s]]
        mid = 0.5 * (p[1:] + p[:-1])
        seg = np.linalg.norm(np.diff(p, axis=0), axis=1)
        t_ray = (seg / (2 + g * (mid[:, 1] + 5))).sum()                                    # slowness integrated along the traced ray
        assert abs(t_ray - tt.ravel()[s]) < 0.03 * tt.ravel()[s] + 0.02, (s, t_ray, tt.ravel()[s])
        assert ln[s] >= np.linalg.norm(p[0] - p[-1]) - 1e-9                               # curved: no shorter than the chord


def test_ray_slots_follow_raystat_and_pairs_without_data_get_no_ray():
    vel = np.full((NX + 2, NY + 2), 2.5)
    srs = np.ones((2, 4), np.int32)
    srs[1, 2] = 0
    srsv = np.arange(8, 0, -1, dtype=np.int32).reshape(2, 4)        # reversed slots
    err, tt, npts, pts, ln, crazy = orc.fm2d_rays(SRC[:2], RCV[:4], srs, vel, X0, Y0, DX, DX, srsv=srsv)
    assert err == 0
    assert npts[8 - 7] == 0 and (np.delete(npts, 1) >= 2).all()    # pair (source 2, receiver 3) -> slot 7 - ... stays empty
    assert np.array_equal(pts[7, 0], RCV[0]) and np.array_equal(pts[0, 0], RCV[3])


def test_a_source_in_the_last_cell_row_returns_the_previous_sources_field():
    """Documents a property of the reference, reproduced by the restatement: with source-grid refinement on, travel()'s edge test
    compares the REFINED extent with the COARSE index vnr (fm2d_ttime.f90:76-87), so a source in the model's last cell row or
    column stops its own refined march after one or two nodes; ttn is one array over the source loop (fm2d_wrapper.f90), so what
    the reference then returns for that source is the previous source's field.  The device path reports it (MCT_E_FM2D_STALE)."""
    rng = np.random.default_rng(5)
    vel = 3.0 + 0.3 * rng.random((19, 40))               # 17 x 38 nodes + the replicated edge
    src = np.array([[0.8, 3.1], [1.5, 37 * 0.25 - 0.07], [2.9, 5.0]])
    rcv = np.array([[0.3, 0.4], [3.1, 8.0]])
    srs = np.ones((3, 2), np.int32)
    unreached = orc.fm2d_unreached(3)
    err, tt, field, _ = orc.fm2d_times(src, rcv, srs, vel, 0.0, 0.0, 0.2, 0.25, sgdl=3, sgs=4, want_field=True)
    orc.fm2d_disarm()
    assert err == 0 and unreached[0] == 0 and unreached[2] == 0 and unreached[1] > 600
    assert (field[1] == field[0]).sum() > 600
    # alone (or first), the same source leaves the field at its initial zeros
    unreached = orc.fm2d_unreached(1)
    err, tt1, field1, _ = orc.fm2d_times(src[1:2], rcv, srs[1:2], vel, 0.0, 0.0, 0.2, 0.25, sgdl=3, sgs=4, want_field=True)
    orc.fm2d_disarm()
    assert (field1[0] == 0).sum() > 600 and unreached[0] > 600
    # without refinement the march is complete
    unreached = orc.fm2d_unreached(3)
    orc.fm2d_times(src, rcv, srs, vel, 0.0, 0.0, 0.2, 0.25, asgr=0)
    orc.fm2d_disarm()
    assert not unreached.any()
