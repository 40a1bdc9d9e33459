"""The restatement against the reference's surfdisp96.f compiled by a REAL Fortran compiler
(oracle/build_ref_surfdisp.sh + oracle/ref_harness/surfdisp_driver.f90).  The build image has none: this test is then
reported skipped.  Wherever gfortran is on PATH it pipes the committed fixture's inputs through the Fortran binary
and requires its outputs to equal both the fixture (made through oracle/f77toc.py) and the restatement, bit for bit."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MCT_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref", "surfdisp96_gfortran")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dispersion_ref.npz")


@pytest.mark.skipif(shutil.which(os.environ.get("FC", "gfortran")) is None, reason="no Fortran compiler on PATH")
@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "surfmodes", "surfdisp96.f")), reason="reference sources not present")
def test_compiled_fortran_equals_fixture_and_restatement():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref_surfdisp.sh")])
    assert os.path.exists(BIN)
    g = np.load(GOLD)
    lines, meta = [], []
    for k in range(int(g["n"])):
        m = g[f"{k}_model"]
        raylov, igr, nm, ie = (int(v) for v in g[f"{k}_sw"])
        fr = g[f"{k}_freqs"]
        t = 1.0 / fr
        lines.append(f"{m.shape[1]} {2 if raylov else 1} {max(nm, 1)} {igr} {len(fr)} {1 if nm > 0 else 0} {float(g[f'{k}_dph'])!r}")
        for row in m:
            lines.append(" ".join(repr(float(v)) for v in row))     # float32 values print exactly as doubles
        lines.append(" ".join(repr(float(v)) for v in t))
        meta.append((k, len(fr) * max(nm, 1), ie))
    out = subprocess.run([BIN], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split()
    pos = 0
    for k, n, ie in meta:
        ierr = int(out[pos]); pos += 1
        cp = np.array([int(x, 16) for x in out[pos:pos + n]], np.uint64).view(np.float64); pos += n
        cg = np.array([int(x, 16) for x in out[pos:pos + n]], np.uint64).view(np.float64); pos += n
        assert ierr == ie and np.array_equal(cp, g[f"{k}_cp"]) and np.array_equal(cg, g[f"{k}_cg"]), f"case {k}"
