"""tools/make_golden_rayleigh_ref.py -- writes tests/golden/grt_rayleigh_secfun_ref.npz: values of the reference's own Rayleigh
surface secular function (startl + SecFunSurf of surfmodes/Rayleigh.f90, translated mechanically by oracle/f90toc_love.py into
oracle/_ref/librayleigh_f2c.so; run oracle/build_ref.sh first) on the fixed columns of tests/test_oracle_grt.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc                                              # noqa: E402
from test_oracle_grt import rayleigh_fixture_points                    # noqa: E402

assert orc.have_rayleigh_reference(), "oracle/_ref/librayleigh_f2c.so missing: run oracle/build_ref.sh"
vals = np.array([orc.grt_rayleigh_secfun_reference(th, vp, vs, rho, f, c)[:2] for th, vp, vs, rho, f, c in rayleigh_fixture_points()])
path = os.path.join(ROOT, "tests", "golden", "grt_rayleigh_secfun_ref.npz")
np.savez_compressed(path, values=vals)
print("wrote", path, vals.shape, os.path.getsize(path), "bytes", "NaN:", int(np.isnan(vals).sum()))

# whole columns: what surfmodes returns (phase velocities of the fundamental Rayleigh mode at example1's frequencies)
from test_oracle_grt import love_fixture_columns, FREQS, crust         # noqa: E402
out = []
cols = list(love_fixture_columns()) + [(*crust([3.2, 3.6, 2.9, 3.8, 4.5], [2.0, 3.0, 4.0, 6.0, 0.0], water=0.6), orc.GRT_PAR_MODELLING)]
for th, vp, vs, rho, par in cols:                                       # two of them with a water layer on top (Stoneley mode)
    ierr, ph = orc.grt_rayleigh_modes_reference(th, vp, vs, rho, FREQS, dc=1e-3, par=par)
    assert ierr == 0
    out.append(ph)
path2 = os.path.join(ROOT, "tests", "golden", "grt_rayleigh_modes_ref.npz")
np.savez_compressed(path2, phase=np.array(out))
print("wrote", path2, np.array(out).shape)
