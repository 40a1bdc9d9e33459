"""tools/soak_widened.py -- randomized soak of the widened path against the oracle (bit equality): fast-marching travel times /
rays over random grids, dicings, refinements, stencil orders, rough velocity maps, sources on edges; generalized R/T over
random low-velocity stacks (Rayleigh, Love, water, both parameter sets, phase and group).  Prints counts; exits 1 on a mismatch."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi
import oracle_lib as orc
from test_oracle_grt import crust

capi.init(0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
rng = np.random.default_rng(seed)
t_end = time.time() + budget
bad = 0
n_fm = n_ray = n_grt = n_stale = 0
while time.time() < t_end - budget / 2:      # ---- fm2d
    nx, ny = int(rng.integers(12, 90)), int(rng.integers(12, 90))
    gdx, gdz = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    sgdl, sgs = int(rng.integers(2, 6)), int(rng.integers(2, 9))
    fom, asgr = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    dx, dy = float(rng.uniform(0.05, 0.3)), float(rng.uniform(0.05, 0.3))
    x0, y0 = float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))
    nmaps = int(rng.integers(1, 4))
    vel = np.zeros((nmaps, nx + 2, ny + 2))
    for m in range(nmaps):
        v = rng.uniform(2.0, 4.5) + rng.uniform(0.0, 0.6) * rng.standard_normal((nx, ny))      # rough: cell-to-cell jumps
        v = np.clip(v, 1.2, 6.0)
        vel[m, 1:-1, 1:-1] = v
        vel[m, 0, :] = vel[m, 1, :]; vel[m, -1, :] = vel[m, -2, :]; vel[m, :, 0] = vel[m, :, 1]; vel[m, :, -1] = vel[m, :, -2]
    nsrc, nrc = int(rng.integers(1, 6)), int(rng.integers(1, 9))
    Lx, Ly = (nx - 1) * dx, (ny - 1) * dy
    src = np.column_stack([x0 + rng.uniform(0, Lx, nsrc), y0 + rng.uniform(0, Ly, nsrc)])
    if rng.uniform() < 0.5: src[0] = [x0, y0 + Ly]                                                 # a corner
    rcv = np.column_stack([x0 + rng.uniform(0.001 * Lx, 0.98 * Lx, nrc), y0 + rng.uniform(0.001 * Ly, 0.98 * Ly, nrc)])
    srs = (rng.uniform(size=(nmaps, nsrc, nrc)) < 0.8).astype(np.int32)
    o = capi.fm2d_opts(gridx=gdx, gridy=gdz, sgref=asgr, sgdic=sgdl, sgext=sgs, order=fom, band=1.0)
    kw = dict(gdx=gdx, gdz=gdz, asgr=asgr, sgdl=sgdl, sgs=sgs, fom=fom, snb=1.0)
    try:
        tt, field = capi.fm2d_times(src, rcv, srs, vel, x0, y0, dx, dy, o, want_field=True)
        rays = capi.fm2d_rays(src, rcv, srs, vel, x0, y0, dx, dy, o)
    except capi.MctError as e:
        print("fm2d error", e); bad += 1; continue
    for m in range(nmaps):
        unreached = orc.fm2d_unreached(nsrc)
        err, to, fo, _ = orc.fm2d_times(src, rcv, srs[m], vel[m], x0, y0, dx, dy, want_field=True, **kw)
        orc.fm2d_disarm()
        marched = [i for i in range(nsrc) if i == 0 or srs[m][i].any()]
        good = [i for i in marched if unreached[i] == 0]      # the others: the reference returns the previous source's field
        n_stale += len(marched) - len(good)
        ok = err == 0 and np.array_equal(tt[m][good], to[good]) and np.array_equal(field[m][good], fo[good])
        cap = rays["pts"].shape[2]
        err2, t2, npts, pts, ln, crazy = orc.fm2d_rays(src, rcv, srs[m], vel[m], x0, y0, dx, dy, cap=cap, **kw)
        slots = [s_ for s_ in range(nsrc * nrc) if s_ // nrc in good]
        ok2 = err2 == 0 and np.array_equal(rays["npts"][m][slots], npts[slots]) and np.array_equal(rays["length"][m][slots], ln[slots]) and \
            all(np.array_equal(rays["pts"][m, s_, :npts[s_]], pts[s_, :npts[s_]]) for s_ in slots) and \
            (len(good) < len(marched) or rays["crazy"][m] == crazy)
        n_fm += 1; n_ray += int(npts.sum() > 0)
        if not (ok and ok2):
            bad += 1
            print("FM2D MISMATCH", dict(nx=nx, ny=ny, nmaps=nmaps, m=m, **kw), ok, ok2, flush=True)
while time.time() < t_end:                   # ---- generalized R/T
    cols = []
    water = rng.uniform() < 0.3
    for _ in range(24):
        nl = int(rng.integers(3, 12))
        vs = np.sort(rng.uniform(2.0, 4.8, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[0] - rng.uniform(0.05, 0.7)
        if rng.uniform() < 0.3 and nl > 5:
            k2 = int(rng.integers(1, nl - 1)); vs[k2] = min(vs[k2], vs[0] - rng.uniform(0.05, 0.4)) if abs(k2 - k) > 1 else vs[k2]
        th = rng.uniform(0.2, 5.0, nl); th[-1] = 0
        cols.append(crust(vs, th, water=rng.uniform(0.2, 3.0) if water else None))
    offs = [0]
    for c in cols: offs.append(offs[-1] + len(c[0]))
    a = [np.concatenate([c[q] for c in cols]) for q in range(4)]
    raylov, pg = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    par = orc.GRT_PAR_LIKELIHOOD if rng.uniform() < 0.5 else orc.GRT_PAR_MODELLING
    nf = int(rng.integers(2, 14))
    fr = np.sort(rng.uniform(0.08, 2.5, nf))[::-1].copy()
    opts = capi.disp_opts(raylov=raylov, phaseGroup=pg, nmodes=0)
    capi.set_grt(True, par)
    ph, gr, ie, rc = capi.surfmodes_batch(*a, offs, fr, opts)
    capi.set_grt(False)
    for c, col in enumerate(cols):
        if ie[c] == 2 or orc.L().orc_nlvls1(orc.f64(col[1]).ctypes.data, orc.f64(col[2]).ctypes.data, len(col[1]), raylov) == 0:
            continue
        ierr, p, g, _ = orc.grt_modes(*col, fr, modetype=raylov, phaseGroup=pg, dc=opts.dphase, par=par, math_mode=orc.PORTABLE, preset=opts.preset)
        n_grt += 1
        if not (ie[c] == ierr and np.array_equal(ph[c], p) and (not pg or np.array_equal(gr[c], g))):
            bad += 1
            print("GRT MISMATCH", dict(raylov=raylov, pg=pg, water=water, c=c, ie=int(ie[c]), ierr=ierr), flush=True)
print(f"soak seed {seed}: {n_fm} eikonal maps ({n_ray} with rays; {n_stale} sources in the last cell row/column excluded), {n_grt} low-velocity columns, mismatches {bad}")
sys.exit(1 if bad else 0)
