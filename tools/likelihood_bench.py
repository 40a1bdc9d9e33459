#!/usr/bin/env python
"""One rjMCMC step of example1's configuration, likelihood included: nuclei in -> (kdtree_to_grid of the box, property
maps, check_model, windowed dispersion, like%vel, fast marching of 11 periods x 8 sources, misfit) -> three doubles out,
all on the resident session (mct_session_propose + mct_session_likelihood_fm2d), next to the same sequence of the CPU
restatements on all host cores.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
from mctomo_b200 import capi, synth
from concurrent.futures import ThreadPoolExecutor

capi.init(0)
grid = synth.make_grid(101, 101, 121)
freqs = synth.example1_freqs()
np_ = len(freqs)
opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
rng = np.random.default_rng(1)
ncells = 300
pts, par = synth.generate_model(grid, ncells, 1001)
nsrc = nrc = 8                                           # examples/example1/sources.dat, receivers.dat: 8 stations each
src = rng.uniform(-4.5, 4.5, (nsrc, 2)); rcv = rng.uniform(-4.5, 4.5, (nrc, 2))
nrr = nsrc * nrc
raystat = np.zeros((np_, 2, nrr), np.int32); raystat[:, 0, :] = 1; raystat[:, 1, :] = np.arange(1, nrr + 1)
ttime = np.zeros((np_, 3, nrr)); ttime[:, 0, :] = rng.uniform(1, 4, (np_, nrr)); ttime[:, 1, :] = 0.01
S = capi.Session(grid, freqs, opts)
r0 = S.set_model(pts, par)
S.set_data(ttime, raystat, sigdep=0)
S.set_fm2d(src, rcv, capi.fm2d_opts())
pvel_cur = r0["pvel"].copy()
O = [np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32)]
orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), *O)
ncores = os.cpu_count() or 1
tg, tgp, tc, tcd, cols = [], [], [], [], []
for step in range(16):
    i = int(rng.integers(ncells))
    pts2 = pts.copy()
    pts2[i] = np.clip(pts[i] + rng.normal(0, 0.4, 3) * np.array([1.0, 1.0, 0.0]), [grid.xmin, grid.ymin, grid.zmin], [grid.xmax, grid.ymax, grid.zmax])
    new = [np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32)]
    orc.kdtree_to_grid(pts2, par, grid, grid.cover_box(), *new)
    m = (O[3] == i + 1) | (new[3] == i + 1)
    ii, jj, kk = np.nonzero(m)
    box = np.array([grid.xmin + ii.min() * grid.dx, grid.ymin + jj.min() * grid.dy, grid.zmin + kk.min() * grid.dz,
                    grid.xmin + ii.max() * grid.dx, grid.ymin + jj.max() * grid.dy, grid.zmin + kk.max() * grid.dz]) + 1e-9 * np.array([-1, -1, -1, 1, 1, 1])
    t0 = time.perf_counter()
    r = S.propose(pts2, par, box)
    t1 = time.perf_counter()
    if r["model_invalid"]:
        S.reject(); continue
    like = S.likelihood_fm2d(pending=True)
    S.accept()
    t2 = time.perf_counter()
    # CPU: the same sequence with the restatements
    win = r["window"]
    orc.kdtree_to_grid(pts2, par, grid, box, *O)
    vpo, rhoo = orc.vs2vp_rho(O[1], orc.LIBM)
    orc.check_model(O[1], grid)
    po, go, io, cnt, nun = orc.surf_dispersion(vpo, O[1], rhoo, grid, win, freqs, math_mode=orc.LIBM)
    t3 = time.perf_counter()
    pvel_cur[win[0] - 1:win[1], win[2] - 1:win[3], :] = po
    vel = np.zeros((grid.nx + 2, grid.ny + 2, np_))
    orc.assemble_vel(pvel_cur, np_, grid.nx, grid.ny, (1, grid.nx, 1, grid.ny), vel)
    def one(mm):
        return orc.fm2d_times(src, rcv, np.ones((nsrc, nrc), np.int32), np.ascontiguousarray(vel[:, :, mm]), grid.xmin, grid.ymin, grid.dx, grid.dy)[1].ravel()
    with ThreadPoolExecutor(ncores) as ex:
        t = np.array(list(ex.map(one, range(np_))))
    mis = orc.surf_misfit(t, ttime, raystat, sigdep=0)
    t4 = time.perf_counter()
    if step >= 3:
        tg.append(t2 - t0); tgp.append(t1 - t0); tc.append(t4 - t2); tcd.append(t3 - t2)
        cols.append((win[1] - win[0] + 1) * (win[3] - win[2] + 1))
    rel = abs(like["like"] - mis["like"]) / abs(mis["like"])
    assert rel < 1e-3, (like["like"], mis["like"])         # libm vs portable math in the dispersion: same likelihood to ~1e-6
    pts = pts2
med = lambda a: 1e3 * float(np.median(a))
print(json.dumps({"what": "one rjMCMC step of example1 (101x101x121, 300 cells, 11 periods, 8x8 stations), likelihood with curved rays included",
                  "steps": len(tg), "median_window_columns": float(np.median(cols)),
                  "gpu_ms": {"step": med(tg), "propose (grid + dispersion window)": med(tgp), "likelihood (like%vel + 88 eikonal problems + misfit)": med(tg) - med(tgp)},
                  "cpu_ms": {"step": med(tc), "grid + dispersion window": med(tcd), "fm2d (one period per thread) + misfit": med(tc) - med(tcd), "cores": ncores,
                             "kind": "port (oracle/, libm)"},
                  "speedup_median": float(np.median(np.array(tc) / np.array(tg)))}))
