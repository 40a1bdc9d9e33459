#!/usr/bin/env python
"""Proposal-replay latency on config C1 (example1's grid): what ONE rjMCMC step costs through the host-pointer
C ABI (sub-box kdtree_to_grid, whole-grid vs2vp/vp2rho, windowed dispersion incl. check_model) next to the CPU
restatement of the same three calls on all host cores.  Writes profiles/r2_replay_C1.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
from mctomo_b200 import capi, synth

capi.init(0)
grid = synth.make_grid(101, 101, 121)
freqs = synth.example1_freqs()
opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
opts_w = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0, check_scope=1)
WINDOWED = len(sys.argv) > 1 and sys.argv[1] == 'windowed'
SESSION = len(sys.argv) > 1 and sys.argv[1] == 'session'
rng = np.random.default_rng(1)
out = {}
for ncells in (100, 300):
    pts, par = synth.generate_model(grid, ncells, 1001)
    G = [np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32)]
    capi.kdtree_to_grid(pts, par, grid, grid.cover_box(), *G)
    O = [a.copy() for a in G]
    VPG, RHOG = capi.vs2vp_rho(G[1])
    tg, tc, cols = [], [], []
    if SESSION:
        S = capi.Session(grid, freqs, opts)
        S.set_model(pts, par, want_maps=False)
    for step in range(24):
        i = int(rng.integers(ncells))
        pts2 = pts.copy()
        pts2[i] = np.clip(pts[i] + rng.normal(0, 0.4, 3) * np.array([1.0, 1.0, 0.0]),  # horizontal moves keep vs(z) monotone -> valid models
                           [grid.xmin, grid.ymin, grid.zmin], [grid.xmax, grid.ymax, grid.zmax])
        # box: the moved cell's extent before and after (what CGAL returns), from the exact maps
        new = [np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape, np.int32)]
        orc.kdtree_to_grid(pts2, par, grid, grid.cover_box(), *new)
        m = (O[3] == i + 1) | (new[3] == i + 1)
        ii, jj, kk = np.nonzero(m)
        box = np.array([grid.xmin + ii.min() * grid.dx, grid.ymin + jj.min() * grid.dy, grid.zmin + kk.min() * grid.dz,
                        grid.xmin + ii.max() * grid.dx, grid.ymin + jj.max() * grid.dy, grid.zmin + kk.max() * grid.dz]) + 1e-9 * np.array([-1, -1, -1, 1, 1, 1])
        w = capi.box_window(grid, box)
        win = (max(w[0] - 1, 1), min(w[1] + 1, grid.nx), max(w[2] - 1, 1), min(w[3] + 1, grid.ny))
        t0 = time.perf_counter()
        if SESSION:    # nuclei in, window maps out; model resident; whole-grid check_model like the reference
            r = S.propose(pts2, par, box)
            S.accept()
            pv, gv, ie, inval, rc = r["pvel"], r["gvel"], r["ierr"], r["model_invalid"], r["rc"]
            t1 = time.perf_counter()
            assert r["window"] == win
            G[0], G[1], G[2], G[3] = S.get_model()
        else:
            capi.kdtree_to_grid(pts2, par, grid, box, *G)
        if SESSION:
            pass
        elif WINDOWED:   # only what changed travels: windowed property maps, window-scoped check_model
            capi.vs2vp_rho_window(G[1], VPG, RHOG, grid, w)
            vpg, rhog = VPG, RHOG
            pv, gv, ie, inval, rc = capi.surf_dispersion(vpg, G[1], rhog, grid, win, freqs, opts_w)
        else:
            vpg, rhog = capi.vs2vp_rho(G[1])
            pv, gv, ie, inval, rc = capi.surf_dispersion(vpg, G[1], rhog, grid, win, freqs, opts)
        if not SESSION:
            t1 = time.perf_counter()
        orc.kdtree_to_grid(pts2, par, grid, box, *O)
        vpo, rhoo = orc.vs2vp_rho(O[1], orc.LIBM)
        inv_o = orc.check_model(O[1], grid)
        po, go, io, cnt, nun = orc.surf_dispersion(vpo, O[1], rhoo, grid, win, freqs, math_mode=orc.LIBM)
        t2 = time.perf_counter()
        assert np.array_equal(G[3], O[3]) and inval == inv_o
        if not inval:
            assert np.array_equal(ie, io) and np.abs(pv - po).max() <= 1e-5
        if step >= 4 and not inval:
            tg.append(t1 - t0); tc.append(t2 - t1); cols.append((win[1] - win[0] + 1) * (win[3] - win[2] + 1))
        pts = pts2
    out[f"ncells_{ncells}"] = {"proposals": len(tg), "median_columns": float(np.median(cols)), "gpu_ms_median": 1e3 * float(np.median(tg)),
                               "cpu_ms_median": 1e3 * float(np.median(tc)), "cpu_cores": os.cpu_count(),
                               "speedup_median": float(np.median(np.array(tc) / np.array(tg)))}
    print(ncells, out[f"ncells_{ncells}"], flush=True)
out["what"] = "C1 (101x101x121, 11 periods, Rayleigh phase): move proposals; GPU = host-pointer C ABI incl. all H2D/D2H; CPU = oracle (libm) on all cores"
out["mode"] = "resident session (mct_session_propose + accept: nuclei in, window maps out, whole-grid check_model)" if SESSION else "windowed transfers (mct_vs2vp_rho_window, check_scope=1)" if WINDOWED else "reference-shaped calls (whole-grid vs2vp, whole-grid check_model)"
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "replay_C1_session.json" if SESSION else "replay_C1_windowed.json" if WINDOWED else "replay_C1.json"), "w"), indent=1)
