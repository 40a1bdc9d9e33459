"""tools/make_golden_fm2d_ref.py -- writes tests/golden/fm2d_travel_ref.npz: outputs of the reference's own `travel`
(fm2d/fm2d_ttime.f90, translated mechanically by oracle/f90toc.py into oracle/_ref/libfm2d_ttime_f2c.so; run
oracle/build_ref.sh first) on the seeded cases of tests/test_oracle_fm2d_vs_reference.py.  Needs /root/reference; the
fixture itself travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc                                              # noqa: E402
from test_oracle_fm2d_vs_reference import cases, run_case, times_cases, field_digest, rays_cases, ray_digests, jumps_of   # noqa: E402

SEED, N = 20261018, 40
assert orc.have_fm2d_reference(), "oracle/_ref/libfm2d_ttime_f2c.so missing: run oracle/build_ref.sh"
out = {"n": N, "seed": SEED}
for k, case in enumerate(cases(SEED, N)):
    for j, (rc, ttn, nsts, heap) in enumerate(run_case("reference", case)):
        out[f"{k}_{j}_rc"] = rc
        out[f"{k}_{j}_ttn"] = ttn
        out[f"{k}_{j}_nsts"] = nsts
        out[f"{k}_{j}_heap"] = heap
# whole calls of modrays for travel times (gridder + the source loop's body + travel + bsplrefine + srtimes, all translated):
# receiver times and a digest of every marched source's field
NT = 24
out["nt"] = NT
for k, (src, rcv, srs, vel, gox, goz, dvx, dvz, kw) in enumerate(times_cases(SEED, NT)):
    err, tt, field = orc.fm2d_times_reference(src, rcv, srs, vel, gox, goz, dvx, dvz, **kw)
    out[f"t{k}_err"] = err
    out[f"t{k}_tt"] = tt
    out[f"t{k}_digest"] = np.frombuffer(field_digest(field, srs), dtype=np.uint8)
# whole calls for group-velocity data (rpaths translated too): times, point counts, a digest of every ray's points, the crazy
# count, and which rays contain the corner jump that follows a 0/0 gradient in the Fortran
NR = 16
out["nr"] = NR
for k, (src, rcv, srs, vel, gox, goz, dvx, dvz, kw, cap) in enumerate(rays_cases(SEED, NR)):
    err, tt, npts, pts, crazy = orc.fm2d_rays_reference(src, rcv, srs, vel, gox, goz, dvx, dvz, cap=cap, **kw)
    assert err == 0
    out[f"r{k}_tt"] = tt
    out[f"r{k}_npts"] = npts
    out[f"r{k}_digest"] = ray_digests(npts, pts)
    out[f"r{k}_jump"] = jumps_of(npts, pts, 0.5 * min(dvx / kw["gdx"], dvz / kw["gdz"]))
    out[f"r{k}_crazy"] = crazy
path = os.path.join(ROOT, "tests", "golden", "fm2d_travel_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
