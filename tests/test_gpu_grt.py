"""GPU parity of the generalized R/T branch (mctomo_b200/csrc/k5_grt.cuh) against oracle/grt_ref.c through the C ABI.

Bit-identical phase and group velocities, ierr and both work counters against the oracle in PORTABLE math mode (same
exp / sincos as the device); within north_star's 1e-5 km/s of the oracle with libm.  The oracle of this branch is
pinned end to end on the reference's own Fortran 90, translated mechanically (tests/test_oracle_grt.py: oracle/f90toc_love.py),
and on physics; the last test of this file holds the device to the reference-derived fixtures directly."""
import numpy as np
import pytest

import oracle_lib as orc
from mctomo_b200 import synth
from mctomo_b200.capi import disp_opts
from test_oracle_grt import FREQS, MODELS, crust

pytestmark = pytest.mark.gpu


def _batch(cols):
    offs = [0]
    for c in cols:
        offs.append(offs[-1] + len(c[0]))
    a = [np.concatenate([c[k] for c in cols]) for k in range(4)]
    return a, offs


def _random_lvl_columns(rng, n, water=False):
    cols = []
    while len(cols) < n:
        nl = int(rng.integers(4, 9))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[0] - rng.uniform(0.1, 0.5)  # a layer slower than the top one: the GRT branch
        th = rng.uniform(0.5, 4.0, nl)
        th[-1] = 0
        c = crust(vs, th, water=rng.uniform(0.3, 2.0) if water else None)
        cols.append(c)
    return cols


@pytest.mark.parametrize("raylov,pg", [(1, 0), (1, 1), (0, 0), (0, 1)])
def test_grt_prelayered_columns_bit_identical(mct, raylov, pg):
    rng = np.random.default_rng(100 + 2 * raylov + pg)
    cols = [MODELS[k] for k in sorted(MODELS)] + _random_lvl_columns(rng, 21)
    # ordinary columns in between: they must keep going through surfdisp96
    vs = np.array([2.5, 3.0, 3.5, 4.2]); cols.insert(2, crust(vs, [1.0, 2.0, 3.0, 0.0]))
    (th, vp, vs_, rho), offs = _batch(cols)
    opts = disp_opts(raylov=raylov, phaseGroup=pg, nmodes=0)
    mct.set_grt(False)
    ph0, gr0, ie0, rc0 = mct.surfmodes_batch(th, vp, vs_, rho, offs, FREQS, opts)
    assert rc0 == mct.MCT_E_GRT_NEEDED and (ie0 == 2).sum() == len(cols) - 1
    mct.reset_stats()
    mct.set_grt(True, orc.GRT_PAR_LIKELIHOOD)
    try:
        ph, gr, ie, rc = mct.surfmodes_batch(th, vp, vs_, rho, offs, FREQS, opts)
        st = mct.grt_stats()
    finally:
        mct.set_grt(False)
    assert rc == 0 and st["columns"] == len(cols) - 1
    tot = np.zeros(2, np.int64)
    nfail = 0
    for c, col in enumerate(cols):
        if c == 2:
            assert ie[c] == 0 and np.array_equal(ph[c], ph0[c]) and np.array_equal(gr[c], gr0[c])
            continue
        ierr, p, g, cnt = orc.grt_modes(*col, FREQS, modetype=raylov, phaseGroup=pg, dc=opts.dphase, par=orc.GRT_PAR_LIKELIHOOD,
                                        math_mode=orc.PORTABLE, preset=opts.preset)
        tot += cnt
        nfail += ierr
        assert ie[c] == ierr, (c, ie[c], ierr)
        assert np.array_equal(ph[c], p), (c, ph[c], p)
        if pg:
            assert np.array_equal(gr[c], g), (c, gr[c], g)
        if ierr == 0:
            _, pl, gl, _ = orc.grt_modes(*col, FREQS, modetype=raylov, phaseGroup=pg, dc=opts.dphase, par=orc.GRT_PAR_LIKELIHOOD,
                                         math_mode=orc.LIBM, preset=opts.preset)
            assert np.abs(pl - ph[c]).max() < 1e-5
    assert st["secfun"] == tot[0] and st["interface_steps"] == tot[1], (st, tot)
    assert nfail < len(cols) // 2


def test_grt_stoneley_under_water(mct):
    rng = np.random.default_rng(7)
    cols = _random_lvl_columns(rng, 12, water=True)
    (th, vp, vs_, rho), offs = _batch(cols)
    opts = disp_opts(raylov=1, phaseGroup=1, nmodes=0)
    mct.set_grt(True, orc.GRT_PAR_MODELLING)
    try:
        ph, gr, ie, rc = mct.surfmodes_batch(th, vp, vs_, rho, offs, FREQS[:7], opts)
    finally:
        mct.set_grt(False)
    for c, col in enumerate(cols):
        ierr, p, g, _ = orc.grt_modes(*col, FREQS[:7], modetype=1, phaseGroup=1, dc=opts.dphase, par=orc.GRT_PAR_MODELLING,
                                      math_mode=orc.PORTABLE, preset=opts.preset)
        assert ie[c] == ierr and np.array_equal(ph[c], p) and np.array_equal(gr[c], g), c


def test_grt_on_model_columns_as_program_modelling_calls_it(mct):
    """Columns of a gridded model (layered on the device in float64), no check_model: forward_modelling.f90:393-429."""
    grid = synth.make_grid(6, 5, 24)
    pts, par = synth.generate_model(grid, 40, 11)
    vp, vs, rho, sid = [np.zeros(grid.shape) for _ in range(3)] + [np.zeros(grid.shape, np.int32)]
    orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
    rng = np.random.default_rng(5)
    lvl = [(1, 2), (3, 0), (4, 4), (0, 1)]
    for (i, j) in lvl:  # a slow zone below the top cell
        k0 = int(rng.integers(6, 12))
        vs[i, j, k0:k0 + 4] = vs[i, j, 0] * 0.8
        vp[i, j, k0:k0 + 4] = 1.73 * vs[i, j, k0]
    freqs = FREQS[:8]
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    win = (1, grid.nx, 1, grid.ny)
    mct.set_grt(True, orc.GRT_PAR_MODELLING)
    try:
        pv, gv, ie, inval, rc = mct.surf_dispersion(vp, vs, rho, grid, win, freqs, opts, check=False)
    finally:
        mct.set_grt(False)
    po, go, io, cnt, nun = orc.surf_dispersion(vp, vs, rho, grid, win, freqs)
    assert nun == len(lvl) and rc == 0
    for i in range(grid.nx):
        for j in range(grid.ny):
            if (i, j) in lvl:
                n, (th, a, b, r) = orc.convert_column(vp[i, j], vs[i, j], rho[i, j], grid.dz)
                ierr, p, g, _ = orc.grt_modes(th, a, b, r, freqs, modetype=1, phaseGroup=0, dc=opts.dphase, par=orc.GRT_PAR_MODELLING,
                                              math_mode=orc.PORTABLE, preset=opts.preset)
                assert ie[i, j] == ierr and np.array_equal(pv[i, j], p), (i, j)
            else:
                assert ie[i, j] == io[i, j] and np.array_equal(pv[i, j], po[i, j])


def test_grt_more_columns_than_one_scratch_chunk(mct):
    """1500 low-velocity columns (the kernel's scratch holds 1024 at a time): three distinct stacks repeated; every copy
    must equal the oracle's answer for its stack."""
    base = [MODELS[k] for k in sorted(MODELS)]
    cols = [base[c % 3] for c in range(1500)]
    (th, vp, vs_, rho), offs = _batch(cols)
    opts = disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    fr = FREQS[:4]
    mct.set_grt(True, orc.GRT_PAR_LIKELIHOOD)
    try:
        ph, gr, ie, rc = mct.surfmodes_batch(th, vp, vs_, rho, offs, fr, opts)
        st = mct.grt_stats()
    finally:
        mct.set_grt(False)
    assert rc == 0 and st["columns"] == 1500
    want = [orc.grt_modes(*b, fr, modetype=1, phaseGroup=0, dc=opts.dphase, par=orc.GRT_PAR_LIKELIHOOD, math_mode=orc.PORTABLE) for b in base]
    for c in range(1500):
        assert ie[c] == want[c % 3][0] and np.array_equal(ph[c], want[c % 3][1]), c


@pytest.mark.parametrize("raylov", [1, 0])
def test_grt_model_columns_under_water_phase_and_group(mct, raylov):
    """waterDepth > 0: the device layers the column with the water layer on top (Rayleigh: the Stoneley search; Love: the
    fluid is skipped by index), phase + group, the modelling variant's constants."""
    grid = synth.make_grid(4, 4, 30, waterDepth=0.8)
    pts, par = synth.generate_model(grid, 30, 3)
    vp, vs, rho, sid = [np.zeros(grid.shape) for _ in range(3)] + [np.zeros(grid.shape, np.int32)]
    orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
    for (i, j, k0) in [(0, 0, 8), (2, 1, 12), (3, 3, 15)]:
        vs[i, j, k0:k0 + 5] = vs[i, j, 0] * 0.8
    vp[:] = 1.73 * vs
    rho[:] = 1.74 * vp ** 0.25
    freqs = FREQS[:6]
    eps5, pre = float(np.float32(1e-5)), 1000.0
    opts = disp_opts(raylov=raylov, phaseGroup=1, nmodes=0, variant="modelling")
    win = (1, grid.nx, 1, grid.ny)
    mct.set_grt(True, orc.GRT_PAR_MODELLING)
    try:
        pv, gv, ie, inval, rc = mct.surf_dispersion(vp, vs, rho, grid, win, freqs, opts, check=False)
    finally:
        mct.set_grt(False)
    nlvl = 0
    for i in range(grid.nx):
        for j in range(grid.ny):
            n, (th, a, b, r) = orc.convert_column(vp[i, j], vs[i, j], rho[i, j], grid.dz, waterDepth=grid.waterDepth, layer_eps=eps5, water_thresh=0.0)
            if orc.L().orc_nlvls1(orc.f64(a).ctypes.data, orc.f64(b).ctypes.data, n, raylov) == 0:
                continue
            nlvl += 1
            ierr, p, g, _ = orc.grt_modes(th, a, b, r, freqs, modetype=raylov, phaseGroup=1, dc=opts.dphase, par=orc.GRT_PAR_MODELLING,
                                          math_mode=orc.PORTABLE, preset=pre)
            assert ie[i, j] == ierr and np.array_equal(pv[i, j], p) and np.array_equal(gv[i, j], g), (i, j, ierr)
    assert nlvl >= 3


def test_grt_many_layers_and_sixty_frequencies(mct):
    """The branch's limits: a 160-layer gradient stack with two slow zones, NP = 60 frequencies (surfdisp96's limit, which
    the dispersion entry points share); and a batch where the low-velocity column is the only one."""
    n = 160
    vs = np.linspace(2.6, 4.6, n)
    vs[40] = 2.3                     # ONE layer each: the reference's predicate wants a strict local minimum of vp
    vs[90] = 2.45                    # (convert_to_layer merges equal cells, so plateaus do not occur in model columns)
    th = np.full(n, 0.12); th[40] = 1.2; th[90] = 0.7; th[-1] = 0
    col = crust(vs, th)
    assert orc.L().orc_nlvls1(orc.f64(col[1]).ctypes.data, orc.f64(col[2]).ctypes.data, n, 1) == 2
    fr = 1.0 / np.linspace(0.6, 12.0, 60)
    for raylov in (1, 0):
        opts = disp_opts(raylov=raylov, phaseGroup=0, nmodes=0)
        mct.set_grt(True, orc.GRT_PAR_LIKELIHOOD)
        try:
            ph, gr, ie, rc = mct.surfmodes_batch(*col, [0, n], fr, opts)
            assert mct.grt_stats()["columns"] == 1
        finally:
            mct.set_grt(False)
        ierr, p, g, cnt = orc.grt_modes(*col, fr, modetype=raylov, phaseGroup=0, dc=opts.dphase, par=orc.GRT_PAR_LIKELIHOOD,
                                        math_mode=orc.PORTABLE, preset=opts.preset)
        assert ie[0] == ierr and np.array_equal(ph[0], p), raylov
        assert (p < 99).sum() >= 10


def test_grt_device_against_the_reference_derived_fixtures_directly(mct):
    """The device against numbers that come from the reference's own statements (tests/golden/grt_{love,rayleigh}_modes_ref.npz:
    whole columns through the mechanically translated setup_grt / C_Interval / FundaMode / secular functions / bisecim, libm
    math) -- without the restatement in between.  The device computes with the portable sin / cos / exp, so the gate is the
    north-star tolerance of 1e-5 km/s; what is observed is some 1e-15."""
    import os
    from test_oracle_grt import love_fixture_columns
    gold = os.path.join(os.path.dirname(__file__), "golden")
    worst = 0.0
    for modetype, name in ((0, "grt_love_modes_ref.npz"), (1, "grt_rayleigh_modes_ref.npz")):
        g = np.load(os.path.join(gold, name))["phase"]
        cols = list(love_fixture_columns())
        if modetype == 1:
            cols.append((*crust([3.2, 3.6, 2.9, 3.8, 4.5], [2.0, 3.0, 4.0, 6.0, 0.0], water=0.6), orc.GRT_PAR_MODELLING))
        assert len(cols) == len(g)
        for k, (th, vp, vs, rho, par) in enumerate(cols):
            (t, p, s, r), offs = _batch([(th, vp, vs, rho)])
            mct.set_grt(True, par)
            try:
                ph, gr, ie, rc = mct.surfmodes_batch(t, p, s, r, offs, FREQS, disp_opts(raylov=modetype, phaseGroup=0, nmodes=0))
            finally:
                mct.set_grt(False)
            assert rc == 0 and ie[0] == 0, (modetype, k, rc, ie)
            d = float(np.abs(ph[0] - g[k]).max())
            assert d <= 1e-5, (modetype, k, d)
            worst = max(worst, d)
    assert worst < 1e-9          # (tighter than required: report a drift long before it matters)
