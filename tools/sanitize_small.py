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
# round-2 additions: all three nearest-nucleus kernels on odd shapes, de-duplication on/off, straight-ray likelihood,
# stat_rti sums, the 2-D product and point location, a slab call
for mode in (1, 2, 0):
    capi.set_k1_mode(mode)
    g2 = synth.make_grid(7, 9, 13)
    p2, a2 = synth.generate_model(g2, 40, 6)
    m = [np.zeros(g2.shape), np.zeros(g2.shape), np.zeros(g2.shape), np.zeros(g2.shape, np.int32)]
    capi.kdtree_to_grid(p2, a2, g2, g2.cover_box(), *m)
    capi.kdtree_to_grid(p2, a2, g2, np.array([-1.0, -2.0, 1.3, 2.0, 3.0, 7.7]), *m)
capi.set_k1_mode(0)
for dd in (True, False):
    capi.set_dedup(dd)
    capi.forward_eval(pts[:3], par[:3], grid, freqs, capi.disp_opts(phaseGroup=1))
capi.set_dedup(True)
np_ = len(freqs)
rays = np.array([[-4.0, -3.0], [0.0, 0.0], [4.0, 3.0]])
rp = np.tile(rays, (np_ * 2, 1)); ro = np.arange(0, 3 * np_ * 2 + 1, 3).astype(np.int64)
tt = np.zeros((np_, 3, 2)); tt[:, 0] = 2.0; tt[:, 1] = 0.1
rs = np.ones((np_, 2, 2), np.int32)
S.set_rays(rp, ro, 2); S.set_data(tt, rs)
S.likelihood()
r = S.propose(pts2, par, grid.cover_box()); S.likelihood(pending=True); S.reject()
S.stat_accumulate(); S.stat_get()
S.close()
capi.voronoi_to_grid_2d(pts[:, :2], par, 9, 8, -5.0, -5.0, 1.25, 1.4)
capi.nearest_nucleus(pts, np.random.default_rng(0).uniform(-5, 5, (50, 3)))
capi.sites_locate(pts, sid, grid, np.random.default_rng(1).uniform(-5, 12, (50, 3)))
print("sanitize_small: ok")
