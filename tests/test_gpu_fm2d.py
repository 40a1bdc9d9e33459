"""GPU parity of the 2-D fast-marching travel times (mctomo_b200/csrc/k6_fm2d.cuh) against oracle/fm2d_ref.c through
the C ABI: receiver times AND the whole travel-time field bit-identical, node / stencil counters identical, for both
stencil orders, with and without source-grid refinement, diced grids, sources on the model's edge, receivers inside the
source cell, sources without data, several periods in one call.  The oracle of this path is itself pinned on the
reference's own Fortran 90, translated mechanically (tests/test_oracle_fm2d_vs_reference.py), and on analytic travel times
(tests/test_oracle_fm2d.py); the last tests of this file hold the device to the reference-derived fixtures directly."""
import numpy as np
import pytest

import oracle_lib as orc
from test_oracle_fm2d import SRC, RCV

pytestmark = pytest.mark.gpu


def _maps(nmaps, nx, ny, seed):
    """smooth random phase-velocity maps with the replicated edge of like%vel"""
    rng = np.random.default_rng(seed)
    out = np.zeros((nmaps, nx + 2, ny + 2))
    x, y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny), indexing="ij")
    for m in range(nmaps):
        v = 2.5 + 0.2 * m
        for _ in range(5):
            kx, ky, ph, a = rng.uniform(1, 6), rng.uniform(1, 6), rng.uniform(0, 6.28), rng.uniform(0.05, 0.25)
            v = v + a * np.sin(kx * x * 3 + ky * y * 3 + ph)
        out[m, 1:-1, 1:-1] = v
        out[m, 0, :] = out[m, 1, :]; out[m, -1, :] = out[m, -2, :]
        out[m, :, 0] = out[m, :, 1]; out[m, :, -1] = out[m, :, -2]
    return out


def _compare(mct, src, rcv, srs, vel, x0, y0, dx, dy, **kw):
    o = mct.fm2d_opts(gridx=kw.get("gdx", 1), gridy=kw.get("gdz", 1), sgref=kw.get("asgr", 1), sgdic=kw.get("sgdl", 4),
                      sgext=kw.get("sgs", 8), order=kw.get("fom", 1), band=kw.get("snb", 0.5))
    mct.reset_stats()
    tt, field = mct.fm2d_times(src, rcv, srs, vel, x0, y0, dx, dy, o, want_field=True)
    st = mct.fm2d_stats()
    tot = np.zeros(2, np.int64)
    stale = False
    for m in range(vel.shape[0]):
        srs_m = srs[m] if np.ndim(srs) == 3 else srs
        unreached = orc.fm2d_unreached(len(src))
        err, to, fo, cnt = orc.fm2d_times(src, rcv, srs_m, vel[m], x0, y0, dx, dy, want_field=True, **kw)
        orc.fm2d_disarm()
        assert err == 0
        tot += cnt
        marched = [i for i in range(len(src)) if i == 0 or srs_m[i].any()]
        # a source whose march dies at the refined-grid edge test returns, in the reference, the PREVIOUS source's field: not a
        # function of the problem's inputs -- excluded from the comparison, the device must flag it (MCT_E_FM2D_STALE)
        good = [i for i in marched if unreached[i] == 0]
        stale = stale or len(good) < len(marched)
        assert np.array_equal(tt[m][good], to[good]), (m, np.abs(tt[m][good] - to[good]).max())
        assert np.array_equal(field[m][good], fo[good]), m
    assert st["accepted"] == tot[0] and st["updates"] == tot[1], (st, tot)
    assert mct.LAST_FM2D_RC == (mct.MCT_E_FM2D_STALE if stale else 0)
    return tt


@pytest.mark.parametrize("fom,asgr", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_fm2d_example1_geometry_bit_identical(mct, fom, asgr):
    vel = _maps(3, 101, 101, 5)
    srs = np.ones((4, 10), np.int32)
    srs[2, :] = 0
    srs[1, 3] = 0
    tt = _compare(mct, SRC, RCV, srs, vel, -5.0, -5.0, 0.1, 0.1, fom=fom, asgr=asgr)
    assert (tt[:, 2] == -1.0).all() and (tt[:, 1, 3] == -1.0).all()


def test_fm2d_diced_anisotropic_grid_and_per_period_raystat(mct):
    vel = _maps(2, 37, 53, 9)
    rng = np.random.default_rng(3)
    src = np.column_stack([rng.uniform(0.0, 7.2, 6), rng.uniform(10.0, 23.0, 6)])
    src[0] = [0.0, 10.0]            # on the model's corner
    src[1] = [7.19, 22.99]
    rcv = np.column_stack([rng.uniform(0.0, 7.2, 17), rng.uniform(10.0, 23.0, 17)])
    srs = (rng.uniform(size=(2, 6, 17)) < 0.7).astype(np.int32)
    _compare(mct, src, rcv, srs, vel, 0.0, 10.0, 0.2, 0.25, gdx=2, gdz=3, sgdl=3, sgs=5)
    _compare(mct, src, rcv, srs, vel, 0.0, 10.0, 0.2, 0.25, gdx=1, gdz=1, sgdl=8, sgs=12, fom=0)  # refined grid larger than the model's


def test_fm2d_errors_are_reported(mct):
    vel = _maps(1, 21, 21, 1)
    o = mct.fm2d_opts()
    with pytest.raises(RuntimeError):
        mct.fm2d_times(np.array([[9.0, 0.0]]), RCV[:3] * 0.1, np.ones((1, 3), np.int32), vel, -1.0, -1.0, 0.1, 0.1, o)
    with pytest.raises(RuntimeError):
        mct.fm2d_times(np.array([[0.0, 0.0]]), RCV[:3] * 0.1, np.ones((1, 3), np.int32), vel, -1.0, -1.0, 0.1, 0.1, mct.fm2d_opts(band=0.001))


@pytest.mark.parametrize("asgr,fom", [(1, 1), (0, 1), (1, 0)])
def test_fm2d_ray_geometry_bit_identical(mct, asgr, fom):
    """rpaths (uar = 0, group-velocity data): every ray's points, point count and length, the crazy-ray count and the travel
    times equal the restatement's; slots follow raystat(:,2,:); pairs without data get no ray."""
    vel = _maps(2, 101, 101, 11)
    rng = np.random.default_rng(4)
    nsrc, nrc = 5, 12
    src = rng.uniform(-4.8, 4.8, (nsrc, 2))
    rcv = rng.uniform(-4.8, 4.8, (nrc, 2))
    rcv[0] = src[0] + [0.03, 0.02]            # inside the source cell: a two-point ray
    srs = (rng.uniform(size=(2, nsrc, nrc)) < 0.85).astype(np.int32)
    srsv = np.stack([rng.permutation(nsrc * nrc).reshape(nsrc, nrc) + 1 for _ in range(2)]).astype(np.int32)
    o = mct.fm2d_opts(sgref=asgr, order=fom)
    got = mct.fm2d_rays(src, rcv, srs, vel, -5.0, -5.0, 0.1, 0.1, o, srsv=srsv)
    cap = got["pts"].shape[2]
    for m in range(2):
        err, tt, npts, pts, ln, crazy = orc.fm2d_rays(src, rcv, srs[m], vel[m], -5.0, -5.0, 0.1, 0.1, asgr=asgr, fom=fom, srsv=srsv[m], cap=cap)
        assert err == 0
        assert np.array_equal(got["ttime"][m], tt) and np.array_equal(got["npts"][m], npts) and got["crazy"][m] == crazy
        for s in range(nsrc * nrc):
            assert np.array_equal(got["pts"][m, s, :npts[s]], pts[s, :npts[s]]), (m, s)
        assert np.array_equal(got["length"][m], ln)
        if asgr == 1:  # without source refinement the field around the source is first-order only, its gradient can vanish at a
            assert (npts[srsv[m][srs[m] == 1] - 1] >= 2).all() and crazy == 0  # node (0/0 in the Fortran): counted as crazy rays
        assert npts.sum() > 0


def test_fm2d_device_entry_reads_like_vel_in_place(mct):
    """mct_fm2d_times_dev on the Fortran's like%vel(np, ny+2, nx+2) layout (element stride np, map stride 1) and on
    dat%raystat(nrr, 2, np) (map stride 2*nrr), asynchronous on a user stream: equal to the host entry point."""
    import torch
    nmaps, nx, ny = 3, 41, 35
    vel = _maps(nmaps, nx, ny, 21)                                   # (nmaps, nx+2, ny+2)
    rng = np.random.default_rng(2)
    nsrc, nrc = 4, 6
    src = np.column_stack([rng.uniform(0.1, 3.9, nsrc), rng.uniform(0.1, 3.3, nsrc)])
    rcv = np.column_stack([rng.uniform(0.1, 3.9, nrc), rng.uniform(0.1, 3.3, nrc)])
    raystat = np.zeros((nmaps, 2, nsrc * nrc), np.int32)
    raystat[:, 0, :] = rng.uniform(size=(nmaps, nsrc * nrc)) < 0.8
    raystat[:, 1, :] = 7
    o = mct.fm2d_opts(sgdic=2, sgext=6)
    want, _ = mct.fm2d_times(src, rcv, raystat[:, 0, :].reshape(nmaps, nsrc, nrc), vel, 0.0, 0.0, 0.1, 0.1, o)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d_vel = torch.from_numpy(np.ascontiguousarray(np.transpose(vel, (1, 2, 0)))).cuda()   # C (nx+2, ny+2, np) == Fortran (np, ny+2, nx+2)
        d_src = torch.from_numpy(np.concatenate([src[:, 0], src[:, 1]])).cuda()
        d_rcv = torch.from_numpy(np.concatenate([rcv[:, 0], rcv[:, 1]])).cuda()
        d_srs = torch.from_numpy(raystat).cuda()
        d_tt = torch.full((nmaps, nsrc, nrc), -1.0, dtype=torch.float64, device="cuda")
        d_err = torch.zeros(nmaps * nsrc, dtype=torch.int32, device="cuda")
        mct.fm2d_times_dev(d_src.data_ptr(), nsrc, d_rcv.data_ptr(), nrc, d_srs.data_ptr(), 2 * nsrc * nrc, d_vel.data_ptr(), nmaps, 1, nmaps,
                           nx, ny, 0.0, 0.0, 0.1, 0.1, o, d_tt.data_ptr(), d_err.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert int(d_err.abs().sum()) == 0
    assert np.array_equal(d_tt.cpu().numpy(), want)


def test_fm2d_source_in_the_last_cell_row_is_flagged(mct):
    """The reference's refined march tests `ix == nnx` with the REFINED extent against the COARSE vnr (fm2d_ttime.f90:76-87), so
    for a source in the model's last cell row/column the source cell itself counts as a refined-grid edge: the march stops after
    one or two nodes and the Fortran returns the previous source's field (one ttn array is shared over the source loop).  The
    restatement reproduces that; the device cannot (its problems are independent) and says so: status 6 / MCT_E_FM2D_STALE,
    every other source bit-identical."""
    vel = _maps(1, 17, 38, 3)
    src = np.array([[0.8, 3.1], [1.5, 37 * 0.25 - 0.07], [2.9, 5.0]])     # the second one: inside the last cell row
    rcv = np.array([[0.3, 0.4], [3.1, 8.0], [2.0, 2.2]])
    srs = np.ones((3, 3), np.int32)
    unreached = orc.fm2d_unreached(3)
    err, to, fo, _ = orc.fm2d_times(src, rcv, srs, vel[0], 0.0, 0.0, 0.2, 0.25, sgdl=3, sgs=4, want_field=True)
    orc.fm2d_disarm()
    assert err == 0 and unreached[1] > 0 and unreached[0] == 0 and unreached[2] == 0
    assert (fo[1] == fo[0]).sum() > 600                                         # the reference's "field" of source 2 is source 1's
    tt, field = mct.fm2d_times(src, rcv, srs, vel, 0.0, 0.0, 0.2, 0.25, mct.fm2d_opts(sgdic=3, sgext=4), want_field=True)
    assert mct.LAST_FM2D_RC == mct.MCT_E_FM2D_STALE
    assert np.array_equal(tt[0][[0, 2]], to[[0, 2]]) and np.array_equal(field[0][[0, 2]], fo[[0, 2]])
    assert (field[0][1] == 0).sum() > 600                                       # unreached nodes read 0 on the device


def test_fm2d_device_equals_the_reference_derived_fixtures_directly(mct):
    """The device against numbers that come from the reference's own statements (tests/golden/fm2d_travel_ref.npz: whole calls of
    modrays for travel times through the mechanical translation, tools/make_golden_fm2d_ref.py) -- without the restatement in
    between: receiver times of every source whose march completes, and the digest of the fields where all of them do."""
    import os
    from test_oracle_fm2d_vs_reference import GOLD, times_cases, field_digest
    g = np.load(GOLD)
    n = nfield = 0
    for k, (src, rcv, srs, vel, gox, goz, dvx, dvz, kw) in enumerate(times_cases(int(g["seed"]), int(g["nt"]))):
        unreached = orc.fm2d_unreached(len(src))
        orc.fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, **kw)          # (only to know which marches die)
        orc.fm2d_disarm()
        o = mct.fm2d_opts(gridx=kw["gdx"], gridy=kw["gdz"], sgref=kw["asgr"], sgdic=kw["sgdl"], sgext=kw["sgs"], order=kw["fom"])
        tt, field = mct.fm2d_times(src, rcv, srs[None], vel[None], gox, goz, dvx, dvz, o, ttime=np.full((1,) + srs.shape, -1.0), want_field=True)
        good = [i for i in range(len(src)) if unreached[i] == 0]
        assert np.array_equal(tt[0][good], g[f"t{k}_tt"][good]), (k, kw)
        if len(good) == len(src):
            assert field_digest(field[0], srs) == g[f"t{k}_digest"].tobytes(), (k, kw)
            nfield += 1
        n += len(good)
    assert n > 40 and nfield > 12


def test_fm2d_device_rays_equal_the_reference_derived_fixtures_directly(mct):
    """Group-velocity data: the device's ray points, point counts, crazy-ray count and times against the fixtures made by the
    translated modrays + rpaths (no restatement in between)."""
    from test_oracle_fm2d_vs_reference import GOLD, rays_cases, check_rays_against_fixture
    g = np.load(GOLD)
    rays = 0
    for k, (src, rcv, srs, vel, gox, goz, dvx, dvz, kw, cap) in enumerate(rays_cases(int(g["seed"]), int(g["nr"]))):
        o = mct.fm2d_opts(gridx=kw["gdx"], gridy=kw["gdz"], sgref=1, sgdic=kw["sgdl"], sgext=kw["sgs"], order=kw["fom"])
        r = mct.fm2d_rays(src, rcv, srs[None], vel[None], gox, goz, dvx, dvz, o, cap=cap)
        tt = np.where(srs == 1, r["ttime"][0], -1.0)
        rays += check_rays_against_fixture(g, k, tt, r["npts"][0], r["pts"][0], int(r["crazy"][0]))
    assert rays > 60
