"""tools/sanitize_small.py -- a small pass through every dispersion kernel shape and the session path, for
compute-sanitizer (memcheck / racecheck):  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
capi.init(0)
grid = synth.make_grid(6, 5, 16)
pts, par = synth.generate_model(grid, 12, 5)
freqs = synth.freqs(3)
vp, vs, rho = np.zeros(grid.shape), np.zeros(grid.shape), np.zeros(grid.shape)
sid = np.zeros(grid.shape, np.int32)
capi.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
vp, rho = capi.vs2vp_rho(vs)
ref = None
for raylov, pg, nm in ((1, 1, 0), (0, 0, 2)):
    opts = capi.disp_opts(raylov=raylov, phaseGroup=pg, nmodes=nm)
    ref = None
    for mode, lanes in ((1, 0), (2, 128), (2, 64), (2, 32), (2, 8), (2, 2)):
        capi.set_k2_mode(mode); capi.set_k2_lanes(lanes)
        pv, gv, ie, inval, rc = capi.surf_dispersion(vp, vs, rho, grid, (1, 6, 1, 5), freqs, opts)
        if ref is None:
            ref = (pv.copy(), gv.copy(), ie.copy())
        assert np.array_equal(pv, ref[0]) and np.array_equal(gv, ref[1]) and np.array_equal(ie, ref[2]), (mode, lanes)
capi.set_k2_mode(0); capi.set_k2_lanes(0)
S = capi.Session(grid, freqs, capi.disp_opts())
S.set_model(pts, par)
pts2 = pts.copy(); pts2[3, 0] += 0.7
r = S.propose(pts2, par, grid.cover_box()); S.reject()
r = S.propose(pts2, par, grid.cover_box()); S.accept()
S.close()
print("sanitize_small: ok")
