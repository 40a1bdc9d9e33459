"""Resident session (mct_session_*): a random walk of rjMCMC-style proposals -- moves, value changes, births,
deaths, accepted and rejected, valid and invalid -- must leave the device-resident model and maps identical, bit
for bit, to what the oracle computes from scratch for the chain's current nuclei, and every proposal's window
maps must equal the oracle's (reference call sequence: src/mcmc_loc2.f90:199-228,556-566,
src/likelihood.f90:75-83, src/likelihood_surf.F90:155-231)."""
import numpy as np
import pytest

import oracle_lib as orc
from mctomo_b200 import synth
from mctomo_b200.capi import disp_opts

pytestmark = pytest.mark.gpu

F1730 = float(np.float32(1.730))


def _full_model(grid, pts, par):
    vp, vs, rho = np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape)
    sid = np.zeros(grid.shape, np.int32)
    orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
    vp, rho = orc.vs2vp_rho(vs)
    return vp, vs, rho, sid


def _box_of(grid, mask):
    ii, jj, kk = np.nonzero(mask)
    lo = np.array([grid.xmin + ii.min() * grid.dx, grid.ymin + jj.min() * grid.dy, grid.zmin + kk.min() * grid.dz])
    hi = np.array([grid.xmin + ii.max() * grid.dx, grid.ymin + jj.max() * grid.dy, grid.zmin + kk.max() * grid.dz])
    return np.concatenate([lo - 1e-9, hi + 1e-9])


@pytest.mark.parametrize("raylov,pg,nm", [(1, 1, 0), (0, 0, 2)])
def test_session_random_walk(mct, raylov, pg, nm):
    grid = synth.make_grid(20, 18, 30)
    freqs = synth.freqs(8)
    opts = disp_opts(raylov=raylov, phaseGroup=pg, nmodes=nm)
    okw = dict(raylov=raylov, phaseGroup=pg, nmodes=nm)
    rng = np.random.default_rng(77 + raylov)
    pts, par = synth.generate_model(grid, 40, 4242)
    S = mct.Session(grid, freqs, opts)
    r0 = S.set_model(pts, par)
    cur = _full_model(grid, pts, par)
    o = orc.surf_dispersion(cur[0], cur[1], cur[2], grid, (1, grid.nx, 1, grid.ny), freqs, **okw)
    assert r0["model_invalid"] == 0 and orc.check_model(cur[1], grid) == 0
    assert np.array_equal(r0["pvel"], o[0]) and np.array_equal(r0["gvel"], o[1]) and np.array_equal(r0["ierr"], o[2])
    maps = [o[0].copy(), o[1].copy(), o[2].copy()]
    seen = dict(accept=0, reject=0, invalid=0, value=0, birth=0, death=0, move=0)
    lo = np.array([grid.xmin, grid.ymin, grid.zmin])
    hi = np.array([grid.xmax, grid.ymax, grid.zmax])
    for step in range(36):
        kind = ("move", "move", "vmove", "value", "birth", "death")[step % 6]
        n = len(pts)
        pts2, par2, pm = pts.copy(), par.copy(), None
        i = int(rng.integers(n))
        if kind == "move":
            pts2[i] = np.clip(pts[i] + rng.normal(0, 0.5, 3) * np.array([1.0, 1.0, 0.0]), lo, hi)
        elif kind == "vmove":     # vertical moves break vs(z) monotonicity now and then: check_model rejects
            pts2[i] = np.clip(pts[i] + rng.normal(0, 2.5, 3) * np.array([0.2, 0.2, 1.0]), lo, hi)
        elif kind == "value":
            par2[i, 1] = par[i, 1] * (1.0 + 0.01 * rng.normal())
            par2[i, 0] = 1.73 * par2[i, 1]
            pm = np.array([par[i, 1] * F1730, par[i, 1], 0.0])  # what the gridded model holds for that cell
        elif kind == "birth":
            p = lo + rng.random(3) * (hi - lo)
            vs = 2.0 + (p[2] - grid.zmin) * 4.0 / (grid.zmax - grid.zmin)
            pts2 = np.vstack([pts, p])
            par2 = np.vstack([par, [1.73 * vs, vs, 2.4]])
        else:                     # death of the last nucleus: the numbering of the others is unchanged
            pts2, par2 = pts[:-1].copy(), par[:-1].copy()
            i = n - 1
        new = _full_model(grid, pts2, par2)
        changed = (new[3] != cur[3]) | (new[1] != cur[1])
        if kind == "value":
            changed |= cur[3] == i + 1
        if not changed.any():
            continue
        box = _box_of(grid, changed)
        r = S.propose(pts2, par2, box, pm=pm)
        w = orc.box_window(grid, box)
        win = (max(w[0] - 1, 1), min(w[1] + 1, grid.nx), max(w[2] - 1, 1), min(w[3] + 1, grid.ny))
        assert r["window"] == tuple(int(v) for v in win)
        inval = orc.check_model(new[1], grid)
        assert r["model_invalid"] == inval
        # the proposed model is on the device
        got = S.get_model()
        for a, b in zip(got, new):
            assert np.array_equal(a, b), kind
        if not inval:
            ow = orc.surf_dispersion(new[0], new[1], new[2], grid, win, freqs, **okw)
            assert np.array_equal(r["pvel"], ow[0]) and np.array_equal(r["gvel"], ow[1]) and np.array_equal(r["ierr"], ow[2]), kind
        accept = bool(rng.random() < 0.6) and not inval
        if inval:
            seen["invalid"] += 1
        if accept:
            S.accept()
            pts, par, cur = pts2, par2, new
            sx, sy = slice(win[0] - 1, win[1]), slice(win[2] - 1, win[3])
            maps[0][sx, sy], maps[1][sx, sy], maps[2][sx, sy] = ow[0], ow[1], ow[2]
            seen["accept"] += 1
        else:
            S.reject()
            seen["reject"] += 1
        seen[kind if kind != "vmove" else "move"] += 1
        got = S.get_model()
        for a, b in zip(got, cur):
            assert np.array_equal(a, b), (kind, accept)
        gm = S.get_maps()
        assert np.array_equal(gm[0], maps[0]) and np.array_equal(gm[1], maps[1]) and np.array_equal(gm[2], maps[2])
    # after the walk the resident maps are the from-scratch maps of the current nuclei
    o = orc.surf_dispersion(cur[0], cur[1], cur[2], grid, (1, grid.nx, 1, grid.ny), freqs, **okw)
    gm = S.get_maps()
    assert np.array_equal(gm[0], o[0]) and np.array_equal(gm[1], o[1]) and np.array_equal(gm[2], o[2])
    assert seen["accept"] >= 8 and seen["reject"] >= 6 and seen["invalid"] >= 1 and seen["value"] >= 3
    S.close()


def test_session_protocol_errors(mct):
    grid = synth.make_grid(8, 8, 10)
    S = mct.Session(grid, synth.freqs(4), disp_opts())
    pts, par = synth.generate_model(grid, 10, 1)
    with pytest.raises(mct.MctError):
        S.propose(pts, par, grid.cover_box())          # no current model yet
    S.set_model(pts, par)
    with pytest.raises(mct.MctError):
        S.accept()                                     # nothing pending
    S.propose(pts, par, grid.cover_box())
    with pytest.raises(mct.MctError):
        S.propose(pts, par, grid.cover_box())          # previous proposal unresolved
    S.reject()
    S.close()


def _random_rays(grid, np_, nrays, rng):
    """Packed rays in the shape fm2d hands to CalGroupTime: wiggly paths of 2-40 points, plus degenerate ones
    (0 or 1 point -> time 0), points on the domain edges and slightly outside (the stencil clamps)."""
    pts, off = [], [0]
    for ip in range(np_):
        for r in range(nrays):
            n = int(rng.choice([0, 1, 2, 3, 17, 40]))
            a = np.array([rng.uniform(grid.xmin, grid.xmax), rng.uniform(grid.ymin, grid.ymax)])
            b = np.array([rng.uniform(grid.xmin, grid.xmax), rng.uniform(grid.ymin, grid.ymax)])
            if n:
                s = np.linspace(0.0, 1.0, n)[:, None]
                p = a + s * (b - a) + rng.normal(0, 0.05, (n, 2))
                if r % 5 == 0:
                    p[0] = [grid.xmin, grid.ymax]              # exactly on a corner
                if r % 7 == 0:
                    p[-1] = [grid.xmax + 0.3, grid.ymin - 0.2]  # outside: index clamps, weights extrapolate
                pts.append(p)
            off.append(off[-1] + n)
    return (np.concatenate(pts) if pts else np.zeros((0, 2))), np.array(off, np.int64)


def test_group_times_on_resident_map(mct):
    """CalGroupTime/GetVelocity on the device (likelihood_surf.F90:454-521) == the oracle, bit for bit."""
    grid = synth.make_grid(20, 18, 30)
    freqs = synth.freqs(6)
    S = mct.Session(grid, freqs, disp_opts(raylov=1, phaseGroup=1, nmodes=0))
    pts, par = synth.generate_model(grid, 40, 4242)
    r0 = S.set_model(pts, par)
    rng = np.random.default_rng(9)
    nrays = 23
    rp, ro = _random_rays(grid, len(freqs), nrays, rng)
    t = S.group_times(rp, ro, nrays)
    ref = orc.cal_group_time(r0["gvel"], grid, rp, ro, nrays)
    assert t.shape == ref.shape == (len(freqs), nrays)
    assert np.array_equal(t, ref)
    assert (t[ro[1:].reshape(len(freqs), nrays) - ro[:-1].reshape(len(freqs), nrays) < 2] == 0).all() and (t > 0).any()
    # a straight ray through a constant map: time = length / velocity
    import torch
    const = torch.full((grid.nx * grid.ny * 6,), 2.5, dtype=torch.float64, device="cuda")
    line = np.array([[-4.0, -3.0], [0.0, 0.0], [4.0, 3.0]])
    off = np.array([0] + [3] * 6, np.int64)
    off = np.concatenate([[0], np.cumsum([3] * 6)]).astype(np.int64)
    tl = mct.group_times_dev(const.data_ptr(), 6, grid, np.tile(line, (6, 1)), off, 1)
    assert np.allclose(tl, 10.0 / 2.5, rtol=1e-14)
    S.close()


def _straight_rays(grid, src, rev, np_, raystat):
    """setup_straightRays (likelihood_surf.F90:408-452): points every dl = min(dx,dy)/2 from source to receiver."""
    dl = min(grid.dx, grid.dy) / 2
    pts, off, dist = [], [0], np.zeros((np_, len(src) * len(rev)))
    for i in range(np_):
        n = 0
        for j in range(len(src)):
            for k in range(len(rev)):
                dx, dy = rev[k, 0] - src[j, 0], rev[k, 1] - src[j, 1]
                ds = np.sqrt(dx ** 2 + dy ** 2)
                dist[i, n] = ds
                if raystat[i, 0, n] == 1:
                    m = int(np.floor(ds / dl)) + 1
                    p = np.zeros((m, 2))
                    for l in range(1, m):
                        p[l - 1] = [src[j, 0] + (l - 1) * dl * dx / ds, src[j, 1] + (l - 1) * dl * dy / ds]
                    p[m - 1] = rev[k]
                    pts.append(p)
                    off.append(off[-1] + m)
                else:
                    off.append(off[-1])
                n += 1
    return np.concatenate(pts), np.array(off, np.int64), dist


@pytest.mark.parametrize("phaseGroup,sigdep", [(1, 0), (0, 1)])
def test_session_likelihood_straight_rays(mct, phaseGroup, sigdep):
    """The whole surface-wave likelihood of the reference's straight-ray mode (likelihood_surf.F90:226-243,356-404) on
    the resident maps: CalGroupTime through like%gvel (= pvel when phaseGroup /= 1: ADVICE r1) + sigma + the Gaussian
    sums, for the current model and for a PENDING proposal (window maps overlaid), bit-identical to the restatement."""
    grid = synth.make_grid(22, 20, 24)
    freqs = synth.freqs(5)
    np_ = len(freqs)
    S = mct.Session(grid, freqs, disp_opts(raylov=1, phaseGroup=phaseGroup, nmodes=0))
    pts, par = synth.generate_model(grid, 30, 77)
    r0 = S.set_model(pts, par)
    rng = np.random.default_rng(5)
    src = rng.uniform([-4, -4], [4, 4], (4, 2))
    rev = rng.uniform([-4, -4], [4, 4], (6, 2))
    nrr = 24
    raystat = np.zeros((np_, 2, nrr), np.int32)
    raystat[:, 0, :] = rng.uniform(size=(np_, nrr)) < 0.8
    ttime = np.zeros((np_, 3, nrr))
    ttime[:, 0, :] = rng.uniform(1.0, 4.0, (np_, nrr))
    ttime[:, 1, :] = rng.uniform(0.05, 0.3, (np_, nrr))
    rp, ro, dist = _straight_rays(grid, src, rev, np_, raystat)
    sn0, sn1 = rng.uniform(0.01, 0.05, np_), rng.uniform(0.02, 0.1, np_)
    S.set_rays(rp, ro, nrr)
    S.set_data(ttime, raystat, sigdep=sigdep, srdist=dist)
    kw = dict(snoise0=sn0, snoise1=sn1) if sigdep else {}

    def ref(maps):
        t = orc.cal_group_time(maps, grid, rp, ro, nrr)
        return t, orc.surf_misfit(t, ttime, raystat, sigdep=sigdep, srdist=dist, **kw)

    cur_map = r0["gvel"] if phaseGroup == 1 else r0["pvel"]
    t_ref, m_ref = ref(cur_map)
    got = S.likelihood(want_arrays=True, **kw)
    assert np.array_equal(got["phase_time"], t_ref) and np.array_equal(got["sigma"], m_ref["sigma"])
    for k in ("like", "misfit", "unweighted_misfit"):
        assert got[k] == m_ref[k], k
    assert np.array_equal(S.group_times(), t_ref)
    # pending proposal: move one nucleus, evaluate its likelihood before deciding
    pts2 = pts.copy()
    pts2[3] += [0.7, -0.5, 0.4]
    box = np.array([-5.0, -5.0, 0.0, 5.0, 5.0, 12.0])
    pr = S.propose(pts2, par, box)
    assert pr["model_invalid"] == 0
    full = mct.forward_eval(pts2, par, grid, freqs, disp_opts(raylov=1, phaseGroup=phaseGroup, nmodes=0))
    t_ref2, m_ref2 = ref(full["gvel"] if phaseGroup == 1 else full["pvel"])
    got2 = S.likelihood(pending=True, want_arrays=True, **kw)
    assert np.array_equal(got2["phase_time"], t_ref2)
    assert got2["like"] == m_ref2["like"] and got2["misfit"] == m_ref2["misfit"]
    with pytest.raises(mct.MctError):
        S.likelihood(**kw)                      # the current model is not addressable while a proposal is pending
    S.reject()
    got3 = S.likelihood(**kw)
    assert got3["like"] == m_ref["like"]
    # host-time entry point (fm2d's times) and the zero-noise condition
    h = mct.surf_misfit(t_ref, ttime, raystat, sigdep=sigdep, srdist=dist, **kw)
    assert h["like"] == m_ref["like"] and h["rc"] == 0
    if sigdep == 0:
        tt0 = ttime.copy()
        tt0[0, 1, np.argmax(raystat[0, 0] == 1)] = 0.0
        assert mct.surf_misfit(t_ref, tt0, raystat)["rc"] == mct.MCT_E_ZERO_NOISE
        assert orc.surf_misfit(t_ref, tt0, raystat)["rc"] == 6
    S.close()


def test_session_invalid_initial_model_and_stat_accumulation(mct):
    """a session whose first model check_model rejects must not expose raw memory; stat_rti's sums."""
    grid = synth.make_grid(10, 9, 12)
    freqs = synth.freqs(4)
    S = mct.Session(grid, freqs, disp_opts())
    pts, par = synth.generate_model(grid, 12, 3)
    bad = par.copy()
    top = np.argmin(pts[:, 2])
    bad[top, 1] = 9.0                                   # fastest cell on top: vs(k) < vs(1) below it
    r = S.set_model(pts, bad)
    assert r["model_invalid"] == 1
    pv, gv, ie = S.get_maps()
    assert (pv == 100.0).all() and (gv == 100.0).all() and (ie == 0).all()
    with pytest.raises(mct.MctError):
        S.group_times(np.zeros((2, 2)), np.array([0, 2] + [2] * (len(freqs) - 1), np.int64), 1)
    r = S.set_model(pts, par)
    assert r["model_invalid"] == 0
    acc = [np.zeros(grid.shape) for _ in range(4)]
    for it in range(3):
        p2 = pts.copy()
        p2[it] += 0.3
        S.set_model(p2, par, want_maps=False)
        S.stat_accumulate()
        vp, vs, rho, sid = S.get_model()
        acc[0] += vs; acc[1] += vs ** 2; acc[2] += vp; acc[3] += vp ** 2
    got, n = S.stat_get()
    assert n == 3
    for a, b in zip(got, acc):
        assert np.array_equal(a, b)
    S.close()


def test_concurrent_host_threads(mct):
    """the library has one context per process; entry points now serialise on a lock, so several host
    threads (ctypes releases the GIL during the calls) may drive it at once -- results equal the single-threaded ones."""
    import threading
    grid = synth.make_grid(12, 10, 20)
    freqs = synth.freqs(5)
    opts = disp_opts(phaseGroup=1)
    models = [synth.generate_model(grid, 20 + 5 * t, 900 + t) for t in range(4)]
    ref = [mct.forward_eval(p, a, grid, freqs, opts, want_model=True) for p, a in models]
    out = [None] * 4
    errs = []

    def work(t):
        try:
            for _ in range(6):
                out[t] = mct.forward_eval(*models[t], grid, freqs, opts, want_model=True)
                S = mct.Session(grid, freqs, opts)
                S.set_model(*models[t], want_maps=False)
                S.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(4):
        for k in ("pvel", "gvel", "ierr", "sites_id", "vs"):
            assert np.array_equal(out[t][k], ref[t][k]), (t, k)


def test_shutdown_and_reinit(mct):
    """mct_shutdown frees every buffer; calls then fail with MCT_E_NOINIT (no fallback), and a new mct_init brings the
    library back with identical results.  A session outliving the shutdown can still be destroyed."""
    grid = synth.make_grid(8, 7, 12)
    freqs = synth.freqs(4)
    pts, par = synth.generate_model(grid, 15, 2)
    opts = disp_opts()
    a = mct.forward_eval(pts, par, grid, freqs, opts)
    S = mct.Session(grid, freqs, opts)
    S.set_model(pts, par)
    mct.shutdown()
    with pytest.raises(mct.MctError) as e:
        mct.forward_eval(pts, par, grid, freqs, opts)
    assert e.value.code == mct.MCT_E_NOINIT
    S.close()
    mct.init(0)
    b = mct.forward_eval(pts, par, grid, freqs, opts)
    assert np.array_equal(a["pvel"], b["pvel"]) and np.array_equal(a["ierr"], b["ierr"])


@pytest.mark.parametrize("sigdep", [0, 1])
def test_session_likelihood_curved_rays_fm2d(mct, sigdep):
    """surf_likelihood for phase-velocity data with curved rays (likelihood_surf.F90:244-336,356-404) entirely on the
    resident maps: like%vel assembled on the device, every (period, source) marched by the fast-marching kernel,
    like%srdist = like%phaseTime, sigma and the Gaussian sums -- for the current model and for a pending proposal --
    bit-identical to the composition of the restatements (assemble_vel -> fm2d -> misfit)."""
    grid = synth.make_grid(31, 27, 20)
    freqs = synth.freqs(4)
    np_ = len(freqs)
    S = mct.Session(grid, freqs, disp_opts(raylov=1, phaseGroup=0, nmodes=0))
    pts, par = synth.generate_model(grid, 40, 21)
    r0 = S.set_model(pts, par)
    rng = np.random.default_rng(8)
    nsrc, nrc = 3, 5
    src = rng.uniform(-4.5, 4.5, (nsrc, 2))
    rcv = rng.uniform(-4.5, 4.5, (nrc, 2))
    nrr = nsrc * nrc
    raystat = np.zeros((np_, 2, nrr), np.int32)
    raystat[:, 0, :] = rng.uniform(size=(np_, nrr)) < 0.8
    raystat[1, 0, nrc:2 * nrc] = 0          # period 2: the second source carries no data and is not marched
    ttime = np.zeros((np_, 3, nrr))
    ttime[:, 0, :] = rng.uniform(1.0, 4.0, (np_, nrr))
    ttime[:, 1, :] = rng.uniform(0.05, 0.3, (np_, nrr))
    sn0, sn1 = rng.uniform(0.01, 0.05, np_), rng.uniform(0.02, 0.1, np_)
    S.set_data(ttime, raystat, sigdep=sigdep, srdist=np.zeros((np_, nrr)) if sigdep else None)
    S.set_fm2d(src, rcv, mct.fm2d_opts(sgdic=3, sgext=5))
    kw = dict(snoise0=sn0, snoise1=sn1) if sigdep else {}

    def ref(maps):
        vel = np.zeros((grid.nx + 2, grid.ny + 2, np_))
        orc.assemble_vel(maps, np_, grid.nx, grid.ny, (1, grid.nx, 1, grid.ny), vel)
        t = np.zeros((np_, nrr))
        for m in range(np_):
            srs = raystat[m, 0].reshape(nsrc, nrc)
            err, tt, _, _ = orc.fm2d_times(src, rcv, srs, np.ascontiguousarray(vel[:, :, m]), grid.xmin, grid.ymin, grid.dx, grid.dy, sgdl=3, sgs=5)
            assert err == 0
            t[m] = np.where(srs == 1, tt, 0.0).ravel()
        return t, orc.surf_misfit(t, ttime, raystat, sigdep=sigdep, srdist=t, **kw)

    t_ref, m_ref = ref(r0["pvel"])
    got = S.likelihood_fm2d(want_arrays=True, **kw)
    assert np.array_equal(got["phase_time"], t_ref), np.abs(got["phase_time"] - t_ref).max()
    assert np.array_equal(got["sigma"], m_ref["sigma"])
    for k in ("like", "misfit", "unweighted_misfit"):
        assert got[k] == m_ref[k], k
    # a pending proposal: evaluated before it is accepted or rejected
    pts2 = pts.copy()
    pts2[5] += [0.9, -0.6, 0.5]
    pr = S.propose(pts2, par, np.array([-5.0, -5.0, 0.0, 5.0, 5.0, 12.0]))
    assert pr["model_invalid"] == 0
    full = mct.forward_eval(pts2, par, grid, freqs, disp_opts(raylov=1, phaseGroup=0, nmodes=0))
    t2, m2 = ref(full["pvel"])
    got2 = S.likelihood_fm2d(pending=True, want_arrays=True, **kw)
    assert np.array_equal(got2["phase_time"], t2) and got2["like"] == m2["like"] and got2["misfit"] == m2["misfit"]
    assert not np.array_equal(t2, t_ref)
    S.reject()
    assert S.likelihood_fm2d(**kw)["like"] == m_ref["like"]
    S.close()


@pytest.mark.parametrize("sigdep", [0, 1])
def test_session_likelihood_curved_rays_group_velocity(mct, sigdep):
    """Group-velocity data (uar = 0): rays bent by the PHASE map, traced on the device, CalGroupTime through the GROUP map
    along them, like%srdist = their lengths (likelihood_surf.F90:259-353) -- against assemble_vel -> fm2d + rpaths ->
    CalGroupTime -> misfit of the restatements, for the current model and a pending proposal."""
    grid = synth.make_grid(31, 27, 20)
    freqs = synth.freqs(3)
    np_ = len(freqs)
    opts = disp_opts(raylov=1, phaseGroup=1, nmodes=0)
    S = mct.Session(grid, freqs, opts)
    pts, par = synth.generate_model(grid, 40, 21)
    r0 = S.set_model(pts, par)
    rng = np.random.default_rng(12)
    nsrc, nrc = 3, 4
    src = rng.uniform(-4.3, 4.3, (nsrc, 2))
    rcv = rng.uniform(-4.3, 4.3, (nrc, 2))
    nrr = nsrc * nrc
    raystat = np.zeros((np_, 2, nrr), np.int32)
    raystat[:, 0, :] = rng.uniform(size=(np_, nrr)) < 0.85
    raystat[:, 1, :] = np.arange(1, nrr + 1)            # the ray of pair n is kept in slot n, as MCTomo's data files have it
    ttime = np.zeros((np_, 3, nrr))
    ttime[:, 0, :] = rng.uniform(1.0, 4.0, (np_, nrr))
    ttime[:, 1, :] = rng.uniform(0.05, 0.3, (np_, nrr))
    sn0, sn1 = rng.uniform(0.01, 0.05, np_), rng.uniform(0.02, 0.1, np_)
    S.set_data(ttime, raystat, sigdep=sigdep, srdist=np.zeros((np_, nrr)) if sigdep else None)
    S.set_fm2d(src, rcv, mct.fm2d_opts())
    kw = dict(snoise0=sn0, snoise1=sn1) if sigdep else {}

    def ref(pmap, gmap):
        vel = np.zeros((grid.nx + 2, grid.ny + 2, np_))
        orc.assemble_vel(pmap, np_, grid.nx, grid.ny, (1, grid.nx, 1, grid.ny), vel)
        rp, ro, ln_all = [], [0], np.zeros((np_, nrr))
        for m in range(np_):
            srs = raystat[m, 0].reshape(nsrc, nrc)
            err, tt, npts, rpts, ln, crazy = orc.fm2d_rays(src, rcv, srs, np.ascontiguousarray(vel[:, :, m]), grid.xmin, grid.ymin, grid.dx, grid.dy,
                                                           srsv=raystat[m, 1].reshape(nsrc, nrc))
            assert err == 0 and crazy == 0
            ln_all[m] = ln
            for s_ in range(nrr):
                rp.append(rpts[s_, :npts[s_]])
                ro.append(ro[-1] + npts[s_])
        t = orc.cal_group_time(gmap, grid, np.concatenate(rp), np.array(ro, np.int64), nrr)
        return t, ln_all, orc.surf_misfit(t, ttime, raystat, sigdep=sigdep, srdist=ln_all, **kw)

    t_ref, ln_ref, m_ref = ref(r0["pvel"], r0["gvel"])
    got = S.likelihood_fm2d(want_arrays=True, **kw)
    assert np.array_equal(got["phase_time"], t_ref), np.abs(got["phase_time"] - t_ref).max()
    assert np.array_equal(got["sigma"], m_ref["sigma"])
    for k in ("like", "misfit", "unweighted_misfit"):
        assert got[k] == m_ref[k], k
    pts2 = pts.copy()
    pts2[7] += [-0.8, 0.6, 0.4]
    pr = S.propose(pts2, par, np.array([-5.0, -5.0, 0.0, 5.0, 5.0, 12.0]))
    assert pr["model_invalid"] == 0
    full = mct.forward_eval(pts2, par, grid, freqs, opts)
    t2, _, m2 = ref(full["pvel"], full["gvel"])
    got2 = S.likelihood_fm2d(pending=True, want_arrays=True, **kw)
    assert np.array_equal(got2["phase_time"], t2) and got2["like"] == m2["like"] and got2["misfit"] == m2["misfit"]
    S.accept()
    assert S.likelihood_fm2d(**kw)["like"] == m2["like"]
    S.close()
