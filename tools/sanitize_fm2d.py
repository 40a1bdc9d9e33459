"""tools/sanitize_fm2d.py -- a small pass through the fast-marching kernel and the generalized R/T kernel for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_fm2d.py
    compute-sanitizer --tool racecheck python tools/sanitize_fm2d.py
    compute-sanitizer --tool synccheck python tools/sanitize_fm2d.py"""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi
capi.init(0)
# the fast-marching kernel (two warps per problem, named barriers): travel times, field, rays; first and mixed order
rng = np.random.default_rng(2)
velm = 3.0 + 0.3 * rng.random((2, 15, 13))
fsrc = np.array([[0.4, 0.5], [1.7, 1.1], [2.3, 0.2]]); frcv = np.array([[0.2, 1.9], [2.4, 1.5], [1.0, 0.3], [2.0, 1.8]])
fsrs = np.ones((2, 3, 4), np.int32); fsrs[1, 2] = 0
for order in (0, 1):
    for sgref in (0, 1):
        fo = capi.fm2d_opts(sgref=sgref, sgdic=2, sgext=2, order=order)
        capi.fm2d_times(fsrc, frcv, fsrs, velm, 0.0, 0.0, 0.21, 0.2, fo, want_field=True)
        capi.fm2d_rays(fsrc, frcv, fsrs, velm, 0.0, 0.0, 0.21, 0.2, fo)
# one low-velocity column through the generalized R/T kernel, Rayleigh and Love
th = np.array([2.0, 3.0, 4.0, 6.0, 0.0]); b = np.array([3.0, 2.2, 3.4, 3.8, 4.4]); a = 1.75 * b; r = 0.32 * a + 0.77
freqs = 1.0 / np.array([4.0, 8.0, 12.0])
capi.set_grt(True)
try:
    for raylov in (1, 0):
        capi.surfmodes_batch(th, a, b, r, np.array([0, 5]), freqs, capi.disp_opts(raylov=raylov, phaseGroup=1))
finally:
    capi.set_grt(False)
print("sanitize_fm2d: ok", capi.grt_stats())
