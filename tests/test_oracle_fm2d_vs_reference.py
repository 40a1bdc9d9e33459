"""Pins the march of the fm2d oracle (oracle/fm2d_ref.c: travel, fouds1, fouds2, addtree, downtree, updtree, bilinear) against
the REFERENCE'S OWN fm2d/fm2d_ttime.f90:

  * tests/golden/fm2d_travel_ref.npz -- calls of `travel` whose outputs (travel-time field, node status, the heap left behind)
    were produced by the reference source itself, translated statement by statement to C by oracle/f90toc.py and compiled
    with gcc (tools/make_golden_fm2d_ref.py).  The fixtures travel; this test runs everywhere.
  * live, where oracle/_ref/libfm2d_ttime_f2c.so exists: fresh random media every run, including homogeneous ones (every
    wavefront node ties with others: the heap's arrangement decides the order) and strong contrasts (negative discriminants).
Every output must be BIT-IDENTICAL, for the first-order and the mixed-order operators, for a whole-grid march (urg = 0),
for the window-limited march of the refined source grid (urg = 1, stops at the first edge node) and for its continuation
from the nodes left in the narrow band (urg = 2).  Likewise gridder, bsplrefine and srtimes of fm2dray_cartesian.f90, each
called alone (live only).  What stays by restatement only: modrays' own glue between these calls (the refinement window,
the refined -> coarse mapping, the narrow-band completion: fm2dray_cartesian.f90:262-420) and rpaths, pinned on analytic
media in test_oracle_fm2d.py."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fm2d_travel_ref.npz")


def medium(rng, nnx, nnz, kind):
    if kind == "homogeneous":
        return np.full((nnx, nnz), 3.0)
    x, z = np.meshgrid(np.linspace(0, 1, nnx), np.linspace(0, 1, nnz), indexing="ij")
    v = 3.0 + 0.0 * x
    for _ in range(4):
        kx, kz, ph, a = rng.uniform(1, 9), rng.uniform(1, 9), rng.uniform(0, 6.28), rng.uniform(0.05, 0.3)
        v = v + a * np.sin(kx * x * 3 + kz * z * 3 + ph)
    if kind == "rough":                        # cell-scale contrasts of a factor 4: stencils without a real root
        v = v * np.where(rng.random((nnx, nnz)) < 0.3, 0.35, 1.0)
    return v


def cases(seed, n):
    """(veln, gox, goz, dnx, dnz, fom, scx, scz, window): the grid sizes, spacings and source positions of a case"""
    rng = np.random.default_rng(seed)
    for k in range(n):
        nnx, nnz = int(rng.integers(5, 42)), int(rng.integers(5, 34))
        dnx, dnz = float(rng.uniform(0.1, 0.6)), float(rng.uniform(0.1, 0.6))
        if k % 5 == 0:
            dnz = dnx
        gox, goz = float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))
        kind = ("smooth", "homogeneous", "rough")[k % 3]
        veln = medium(rng, nnx, nnz, kind)
        mode = k % 4
        if mode == 0:                          # anywhere inside
            scx, scz = gox + rng.uniform(0, (nnx - 1) * dnx), goz + rng.uniform(0, (nnz - 1) * dnz)
        elif mode == 1:                        # exactly on a node
            scx, scz = gox + int(rng.integers(0, nnx)) * dnx, goz + int(rng.integers(0, nnz)) * dnz
        elif mode == 2:                        # in the last cell row / column
            scx, scz = gox + (nnx - 1) * dnx - rng.uniform(0, dnx), goz + (nnz - 1) * dnz - rng.uniform(0, dnz)
        else:                                  # on the first edge
            scx, scz = gox, goz + rng.uniform(0, (nnz - 1) * dnz)
        isx = min(int((scx - gox) / dnx) + 1, nnx - 1)
        isz = min(int((scz - goz) / dnz) + 1, nnz - 1)
        ext = int(rng.integers(1, 5))
        # a window as modrays builds it (clipped at the grid), seen through travel's own edge test
        window = (max(isx - ext, 1), min(isx + ext, nnx), max(isz - ext, 1), min(isz + ext, nnz))
        yield veln, gox, goz, dnx, dnz, k % 2, float(scx), float(scz), window


def run_case(impl, case):
    """whole-grid march; window-limited march; its continuation"""
    veln, gox, goz, dnx, dnz, fom, scx, scz, window = case
    out = [orc.fm2d_travel(impl, veln, gox, goz, dnx, dnz, fom, scx, scz, urg=0)]
    r1 = orc.fm2d_travel(impl, veln, gox, goz, dnx, dnz, fom, scx, scz, urg=1, window=window)
    out.append(r1)
    # what modrays does between the two marches, reduced to its effect on the status array: alive nodes with a far neighbour
    # go back into the narrow band (status 1), everything else that is not alive becomes far
    t, s = r1[1], r1[2].copy()
    alive = s == 0
    s[:] = -1
    s[alive] = 0
    pad = np.pad(s, 1, constant_values=0)
    far_nb = (pad[:-2, 1:-1] == -1) | (pad[2:, 1:-1] == -1) | (pad[1:-1, :-2] == -1) | (pad[1:-1, 2:] == -1)
    s[alive & far_nb] = 1
    out.append(orc.fm2d_travel(impl, veln, gox, goz, dnx, dnz, fom, scx, scz, urg=2, ttn=t, nsts=s))
    return out


def same(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def test_restatement_reproduces_the_reference_fixtures_bit_for_bit():
    g = np.load(GOLD)
    n = int(g["n"])
    seed = int(g["seed"])
    stopped_early = ties = 0
    for k, case in enumerate(cases(seed, n)):
        got = run_case("port", case)
        for j, r in enumerate(got):
            assert r[0] == int(g[f"{k}_{j}_rc"]), (k, j)
            assert np.array_equal(r[1], g[f"{k}_{j}_ttn"]), f"case {k} march {j}: {(r[1] != g[f'{k}_{j}_ttn']).sum()} travel times differ"
            assert np.array_equal(r[2], g[f"{k}_{j}_nsts"]), (k, j)
            assert np.array_equal(r[3], g[f"{k}_{j}_heap"]), (k, j)
        stopped_early += len(got[1][3]) > 0
        ties += case[0].std() == 0
    assert n >= 36 and stopped_early >= 10 and ties >= 10            # the fixture covers the early stop and the tie-breaking


@pytest.mark.skipif(not orc.have_fm2d_reference(), reason="oracle/_ref/libfm2d_ttime_f2c.so not built (needs /root/reference)")
def test_restatement_equals_the_translated_reference_on_fresh_media():
    seed = orc.live_seed("test_restatement_equals_the_translated_reference_on_fresh_media")
    n = nodes = 0
    for case in cases(seed, 120):
        ref = run_case("reference", case)
        got = run_case("port", case)
        for j in range(3):
            assert same(ref[j], got[j]), f"seed {seed} case {n} march {j}"
        n += 1
        nodes += case[0].size
    assert n == 120 and nodes > 30000


@pytest.mark.skipif(not orc.have_fm2d_reference(), reason="oracle/_ref/libfm2d_ttime_f2c.so not built (needs /root/reference)")
def test_source_outside_the_grid_is_the_fortran_stop():
    veln = np.full((8, 9), 3.0)
    for impl in ("reference", "port"):
        rc, t, s, h = orc.fm2d_travel(impl, veln, 0.0, 0.0, 0.5, 0.5, 1, 9.0, 1.0)
        assert rc == 1 and not t.any() and len(h) == 0


needs_ref = pytest.mark.skipif(not orc.have_fm2d_reference(), reason="oracle/_ref/libfm2d_ttime_f2c.so not built (needs /root/reference)")


@needs_ref
def test_gridder_and_bsplrefine_equal_the_translated_reference():
    rng = np.random.default_rng(orc.live_seed("test_gridder_and_bsplrefine_equal_the_translated_reference"))
    for k in range(60):
        nvx, nvz = int(rng.integers(2, 14)), int(rng.integers(2, 12))
        gdx, gdz = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        velv = 2.0 + rng.random((nvx + 2, nvz + 2))
        a, b = orc.fm2d_gridder("reference", velv, gdx, gdz), orc.fm2d_gridder("port", velv, gdx, gdz)
        assert a.shape == ((nvx - 1) * gdx + 1, (nvz - 1) * gdz + 1) and np.array_equal(a, b) and a.min() > 1.0, k
        nnx, nnz = a.shape
        sgdl, ext = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        isx, isz = int(rng.integers(1, max(2, nnx))), int(rng.integers(1, max(2, nnz)))
        window = (max(isx - ext, 1), min(isx + ext, nnx), max(isz - ext, 1), min(isz + ext, nnz))
        a, b = orc.fm2d_bsplrefine("reference", velv, gdx, gdz, sgdl, window), orc.fm2d_bsplrefine("port", velv, gdx, gdz, sgdl, window)
        assert np.array_equal(a, b) and a.min() > 1.0, (k, window, sgdl)      # (every refined node is written)


@needs_ref
def test_srtimes_equals_the_translated_reference():
    rng = np.random.default_rng(orc.live_seed("test_srtimes_equals_the_translated_reference"))
    n_near = 0
    for k in range(80):
        nnx, nnz = int(rng.integers(4, 30)), int(rng.integers(4, 30))
        dnx, dnz = float(rng.uniform(0.1, 0.6)), float(rng.uniform(0.1, 0.6))
        gox, goz = float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))
        veln, ttn = 2.0 + rng.random((nnx, nnz)), 10.0 * rng.random((nnx, nnz))
        scx, scz = gox + rng.uniform(0, (nnx - 1) * dnx), goz + rng.uniform(0, (nnz - 1) * dnz)
        nrc = 12
        rcv = np.stack([gox + rng.uniform(0, (nnx - 1) * dnx, nrc), goz + rng.uniform(0, (nnz - 1) * dnz, nrc)], axis=1)
        rcv[0] = (scx + 0.3 * min(dnx, dnz), scz)                        # closer than a node spacing: the near-source formula
        rcv[0, 0] = min(rcv[0, 0], gox + (nnx - 1) * dnx)
        rcv[1] = (gox + (nnx - 1) * dnx, goz + (nnz - 1) * dnz)          # the far corner: last cell row and column
        srs = (rng.random(nrc) < 0.8).astype(np.int32)
        srs[:2] = 1
        ra, ta = orc.fm2d_srtimes("reference", veln, ttn, gox, goz, dnx, dnz, float(scx), float(scz), rcv, srs)
        rb, tb = orc.fm2d_srtimes("port", veln, ttn, gox, goz, dnx, dnz, float(scx), float(scz), rcv, srs)
        assert ra == 0 and rb == 0 and np.array_equal(ta, tb), k
        assert np.array_equal(ta == -1.0, srs == 0)                       # receivers without data are not touched
        n_near += 1
    rcv[2] = (gox - 1.0, goz)                                             # a receiver outside: STOP / error 3
    srs[2] = 1
    ra, _ = orc.fm2d_srtimes("reference", veln, ttn, gox, goz, dnx, dnz, float(scx), float(scz), rcv, srs)
    rb, _ = orc.fm2d_srtimes("port", veln, ttn, gox, goz, dnx, dnz, float(scx), float(scz), rcv, srs)
    assert ra == 1 and rb == 3


def _times_case(rng, k):
    nvx, nvz = int(rng.integers(3, 26)), int(rng.integers(3, 22))
    if k % 4 == 1:
        nvx, nvz = nvx + 9, nvz + 9                                    # room for a refined grid that needs no reallocation
    gdx, gdz = (1, 1) if k % 3 else (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    dvx, dvz = float(rng.uniform(0.2, 1.0)), float(rng.uniform(0.2, 1.0))
    gox, goz = float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))
    x, z = np.meshgrid(np.linspace(0, 1, nvx + 2), np.linspace(0, 1, nvz + 2), indexing="ij")
    vel = 3.0 + 0.4 * np.sin(5 * x + 1.3 * k) * np.cos(4 * z) + 0.2 * rng.random((nvx + 2, nvz + 2))
    if k % 5 == 0:
        vel[:] = 3.0                                                   # homogeneous: ties
    nsrc, nrc = int(rng.integers(1, 6)), int(rng.integers(1, 7))
    ex, ez = (nvx - 1) * dvx, (nvz - 1) * dvz
    src = np.stack([gox + rng.uniform(0, ex, nsrc), goz + rng.uniform(0, ez, nsrc)], axis=1)
    if k % 4 == 1:
        src[-1] = (gox + ex - 0.3 * dvx / gdx, goz + 0.5 * ez)         # inside the last cell column: the march dies, the field is stale
    rcv = np.stack([gox + rng.uniform(0, ex, nrc), goz + rng.uniform(0, ez, nrc)], axis=1)
    srs = (rng.random((nsrc, nrc)) < 0.75).astype(np.int32)
    if k % 6 == 2 and nsrc > 1:
        srs[1] = 0                                                     # a source without data is skipped
    kw = dict(gdx=gdx, gdz=gdz, asgr=int(k % 7 != 3), sgdl=int(rng.integers(1, 6)), sgs=int(rng.integers(1, 9)), fom=int(k % 2))
    if k % 4 == 1:
        kw.update(sgdl=int(rng.integers(1, 4)), sgs=int(rng.integers(1, 3)))
    return src, rcv, srs, vel, gox, goz, dvx, dvz, kw


def times_cases(seed, n):
    """whole-call cases whose outcome is defined in Fortran: no heap overrun, no dead march on reallocated arrays"""
    rng = np.random.default_rng(seed)
    k = got = 0
    while got < n:
        case = _times_case(rng, k)
        k += 1
        src, rcv, srs, vel, gox, goz, dvx, dvz, kw = case
        unreached = orc.fm2d_unreached(len(src))
        err, _, field, _ = orc.fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, want_field=True, **kw)
        orc.fm2d_disarm()
        realloc = kw["asgr"] == 1 and (2 * kw["sgs"] * kw["sgdl"] + 1 > min(field.shape[1:]))
        if err == 0 and not (realloc and unreached.max() > 0):
            got += 1
            yield case


def field_digest(field, srs):
    marched = [i for i in range(len(srs)) if i == 0 or srs[i].any()]
    return hashlib.sha256(np.ascontiguousarray(field[marched]).tobytes()).digest()


def test_restatement_reproduces_the_reference_fixtures_of_whole_calls():
    g = np.load(GOLD)
    nt = int(g["nt"])
    stale = 0
    for k, (src, rcv, srs, vel, gox, goz, dvx, dvz, kw) in enumerate(times_cases(int(g["seed"]), nt)):
        unreached = orc.fm2d_unreached(len(src))
        err, tt, field, _ = orc.fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, want_field=True, **kw)
        orc.fm2d_disarm()
        assert err == int(g[f"t{k}_err"]) == 0
        assert np.array_equal(tt, g[f"t{k}_tt"]), (k, kw)
        assert field_digest(field, srs) == g[f"t{k}_digest"].tobytes(), (k, kw)
        stale += int(unreached.max() > 0)
    assert nt >= 24 and stale >= 1


@needs_ref
def test_travel_times_of_whole_calls_equal_the_translated_modrays():
    """End to end for phase-velocity data: orc_fm2d_times (the oracle every GPU test compares with) against gridder + the body
    of modrays' source loop + travel + bsplrefine + srtimes as the reference's own statements execute them -- receiver times
    and the whole field of every marched source, sources whose march dies and return the previous source's field included
    (the driver keeps ONE ttn array over the source loop, as modrays does); tiny models where the refined grid outgrows the
    propagation grid and modrays reallocates its arrays; source-grid refinement on and off; dicing 1..3."""
    seed = orc.live_seed("test_travel_times_of_whole_calls_equal_the_translated_modrays")
    rng = np.random.default_rng(seed)
    n = stale = big = overrun = 0
    for k in range(90):
        src, rcv, srs, vel, gox, goz, dvx, dvz, kw = _times_case(rng, k)
        unreached = orc.fm2d_unreached(len(src))
        e0, t0, f0, _ = orc.fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, want_field=True, **kw)
        orc.fm2d_disarm()
        e1, t1, f1 = orc.fm2d_times_reference(src, rcv, srs, vel, gox, goz, dvx, dvz, **kw)
        if e0 == 2:        # narrow band larger than snb * nnx * nnz (tiny grids): the Fortran overruns btg(maxbt) -- undefined
            overrun += 1
            continue
        assert e0 == e1 == 0, (seed, k, e0, e1)
        nnx, nnz = f0.shape[1:]
        # does modrays reallocate veln / ttn / nsts for the refined grid (fm2dray_cartesian.f90:316-330)?  Then what a dead march
        # returns is a fresh ALLOCATE's content -- undefined in Fortran (zeros in the translation, the previous source's field in
        # the restatement): not compared.  Without reallocation the stale field itself is pinned.
        realloc = kw["asgr"] == 1 and (2 * kw["sgs"] * kw["sgdl"] + 1 > min(nnx, nnz))
        cmp = [i for i in range(len(src)) if (i == 0 or srs[i].any()) and (unreached[i] == 0 or not realloc)]
        assert np.array_equal(t0[cmp], t1[cmp]), (seed, k, kw)
        assert np.array_equal(f0[cmp], f1[cmp]), (seed, k, kw)
        n += 1
        stale += int(any(unreached[i] > 0 for i in cmp))
        big += int(realloc)
    assert n + overrun == 90 and overrun <= 10 and stale >= 2 and big >= 5, (seed, n, overrun, stale, big)


def rays_cases(seed, n):
    """whole-call cases for group-velocity data: refinement on, receivers off the last cell row / column (rpaths STOPs there),
    no heap overrun, no dead march"""
    rng = np.random.default_rng(seed + 1)
    k = got = 0
    while got < n:
        src, rcv, srs, vel, gox, goz, dvx, dvz, kw = _times_case(rng, k)
        k += 1
        kw["asgr"] = 1
        nnx, nnz = (vel.shape[0] - 3) * kw["gdx"] + 1, (vel.shape[1] - 3) * kw["gdz"] + 1
        rcv[:, 0] = np.minimum(rcv[:, 0], gox + (nnx - 1.001) * dvx / kw["gdx"])
        rcv[:, 1] = np.minimum(rcv[:, 1], goz + (nnz - 1.001) * dvz / kw["gdz"])
        unreached = orc.fm2d_unreached(len(src))
        err = orc.fm2d_times(src, rcv, srs, vel, gox, goz, dvx, dvz, **kw)[0]
        orc.fm2d_disarm()
        if err == 0 and unreached.max() == 0:
            got += 1
            yield src, rcv, srs, vel, gox, goz, dvx, dvz, kw, nnx * nnz + 2


def ray_digests(npts, pts):
    return np.stack([np.frombuffer(hashlib.sha256(np.ascontiguousarray(pts[s_, :npts[s_]]).tobytes()).digest(), dtype=np.uint8)
                     for s_ in range(len(npts))])


def jumps_of(npts, pts, dpl):
    """slots whose reference ray contains the corner jump that follows a 0/0 gradient (see the live test below)"""
    out = np.zeros(len(npts), bool)
    for s_ in range(len(npts)):
        q = pts[s_, :npts[s_]]
        if len(q) > 1:
            out[s_] = np.sqrt(((q[1:] - q[:-1]) ** 2).sum(1)).max() > 3.0 * dpl
    return out


def check_rays_against_fixture(g, k, tt, npts, pts, crazy):
    """shared by the CPU test (restatement) and the GPU test (device): times, point counts, every point, crazy-ray count"""
    assert np.array_equal(tt, g[f"r{k}_tt"]), k
    ref_n, ref_d, jump = g[f"r{k}_npts"], g[f"r{k}_digest"], g[f"r{k}_jump"]
    d = ray_digests(npts, pts)
    for s_ in range(len(ref_n)):
        if jump[s_]:
            assert npts[s_] == 1, (k, s_)              # ended as a crazy ray instead of following the jump
        else:
            assert npts[s_] == ref_n[s_] and np.array_equal(d[s_], ref_d[s_]), (k, s_, npts[s_], ref_n[s_])
    assert crazy == int(g[f"r{k}_crazy"]) + int(jump.sum()), k
    return int((ref_n > 1).sum())


def test_restatement_reproduces_the_reference_ray_fixtures():
    g = np.load(GOLD)
    rays = 0
    for k, (src, rcv, srs, vel, gox, goz, dvx, dvz, kw, cap) in enumerate(rays_cases(int(g["seed"]), int(g["nr"]))):
        err, tt, npts, pts, _, crazy = orc.fm2d_rays(src, rcv, srs, vel, gox, goz, dvx, dvz, cap=cap, **kw)
        assert err == 0
        rays += check_rays_against_fixture(g, k, tt, npts, pts, crazy)
    assert rays > 60


@needs_ref
def test_rays_of_whole_calls_equal_the_translated_modrays():
    """Group-velocity data (uar = 0): rpaths too is the reference's own statements here (fm2dray_cartesian.f90:801-1456; its
    T_RAY container replaced by the driver's store).  Every ray point, the point counts, the crazy-ray count and the receiver
    times of orc_fm2d_rays must equal the translation's.  Source-grid refinement on, as shipped: without it the Fortran reads
    a variable it never assigned (ipzr) and divides 0 by 0 next to the source -- outcomes the restatement flags instead."""
    seed = orc.live_seed("test_rays_of_whole_calls_equal_the_translated_modrays")
    rng = np.random.default_rng(seed)
    n = rays = crazies = jumps = 0
    for k in range(110):
        src, rcv, srs, vel, gox, goz, dvx, dvz, kw = _times_case(rng, k)
        kw["asgr"] = 1
        nnx, nnz = (vel.shape[0] - 3) * kw["gdx"] + 1, (vel.shape[1] - 3) * kw["gdz"] + 1
        # rpaths refuses receivers in the last cell row / column (ipx >= nnx: STOP): keep them off it
        rcv[:, 0] = np.minimum(rcv[:, 0], gox + (nnx - 1.001) * dvx / kw["gdx"])
        rcv[:, 1] = np.minimum(rcv[:, 1], goz + (nnz - 1.001) * dvz / kw["gdz"])
        cap = nnx * nnz + 2                                            # the Fortran's own limit (maxrp + 1 points)
        unreached = orc.fm2d_unreached(len(src))
        e0, t0, n0, p0, l0, c0 = orc.fm2d_rays(src, rcv, srs, vel, gox, goz, dvx, dvz, cap=cap, **kw)
        orc.fm2d_disarm()
        if e0 == 2 or unreached.max() > 0:                             # heap overrun / a dead march: undefined in the Fortran (see above)
            continue
        e1, t1, n1, p1, c1 = orc.fm2d_rays_reference(src, rcv, srs, vel, gox, goz, dvx, dvz, cap=cap, **kw)
        assert e0 == e1 == 0, (seed, k, e0, e1)
        assert np.array_equal(t0, t1), (seed, k, kw)
        dpl = 0.5 * min(dvx / kw["gdx"], dvz / kw["gdz"])
        zero_grad = 0
        for slot in range(len(n0)):
            if n0[slot] == n1[slot]:
                assert np.array_equal(p0[slot, :n0[slot]], p1[slot, :n1[slot]]), (seed, k, slot, kw)
                continue
            # The one outcome the restatement (and the device) deliberately do not reproduce: a gradient that vanishes exactly
            # (symmetry, in a homogeneous medium) is 0/0 in the Fortran; the NaN point becomes INT_MIN + 1 in floor(), the
            # boundary clamp then puts the "ray" on the model's corner, from where it walks on to the source.  The restatement
            # ends such a ray as a crazy one (one point).  Recognised here by the jump.
            q = p1[slot, :n1[slot]]
            steps = np.sqrt(((q[1:] - q[:-1]) ** 2).sum(1))
            assert n0[slot] == 1 and steps.max() > 3.0 * dpl, (seed, k, slot, kw, n0[slot], n1[slot])
            zero_grad += 1
        assert c0 == c1 + zero_grad, (seed, k, c0, c1, zero_grad)
        n += 1
        rays += int((n0 > 1).sum())
        crazies += c1
        jumps += zero_grad
    print(f"rays: {n} calls, {rays} rays traced, {crazies} crazy in the reference, {jumps} zero-gradient jumps")
    assert n >= 50 and rays > 250 and jumps < rays // 8, (seed, n, rays, crazies, jumps)
