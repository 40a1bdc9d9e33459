"""tools/make_golden_love_ref.py -- writes tests/golden/grt_love_secfun_ref.npz: values of the reference's own Love secular
function (surfmodes/Love.f90 translated mechanically by oracle/f90toc_love.py into oracle/_ref/liblove_f2c.so; run
oracle/build_ref.sh first) on the fixed columns of tests/test_oracle_grt.py.  Needs /root/reference; the fixture travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc                                              # noqa: E402
from test_oracle_grt import love_fixture_points, FREQS as FREQS_ALL   # noqa: E402

assert orc.have_love_reference(), "oracle/_ref/liblove_f2c.so missing: run oracle/build_ref.sh"
vals = np.array([orc.grt_love_secfun_reference(th, vp, vs, rho, f, c) for th, vp, vs, rho, f, c in love_fixture_points()])
path = os.path.join(ROOT, "tests", "golden", "grt_love_secfun_ref.npz")
np.savez_compressed(path, values=vals)
print("wrote", path, vals.shape, os.path.getsize(path), "bytes")

# whole columns: what surfmodes returns (phase velocities of the fundamental Love mode at example1's frequencies), both parameter sets
from test_oracle_grt import love_fixture_columns                        # noqa: E402
out = []
for th, vp, vs, rho, par in love_fixture_columns():
    ierr, ph = orc.grt_love_modes_reference(th, vp, vs, rho, FREQS_ALL, dc=1e-3, par=par)
    assert ierr == 0
    out.append(ph)
path2 = os.path.join(ROOT, "tests", "golden", "grt_love_modes_ref.npz")
np.savez_compressed(path2, phase=np.array(out))
print("wrote", path2, np.array(out).shape)
