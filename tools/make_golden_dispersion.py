#!/usr/bin/env python
"""Generates tests/golden/dispersion_oracle.npz: SELF-GENERATED regression vectors for stage 2.

The reference ships no dispersion values and no Fortran compiler exists in the build image, so these
vectors come from OUR restatement (oracle/surfdisp96_ref.c, libm math) -- they pin the oracle against
drift, they do not pin it against the reference ("parity unpinned", see the oracle's header).
Inputs taken from the reference tree: examples/example2/surf/initial_model.dat (7-layer depth/vp/vs
model) and the example1 frequency list (examples/example1/otimes.dat:2).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc  # noqa: E402
from mctomo_b200 import synth  # noqa: E402


def example2_model():
    m = np.loadtxt("/root/reference/examples/example2/surf/initial_model.dat", skiprows=1)
    depth, vp, vs = m[:, 0], m[:, 1], m[:, 2]
    thick = np.append(np.diff(depth), 0.0)
    rho = 1.74 * vp ** 0.25
    return thick, vp, vs, rho


def main():
    out = {}
    f11 = synth.example1_freqs()
    thick, vp, vs, rho = example2_model()
    out["ex2_model"] = np.stack([thick, vp, vs, rho])
    out["freqs11"] = f11
    for mt, name in ((1, "ray"), (0, "love")):
        for pg in (0, 1):
            rc, ph, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, f11, mt, pg, 0, math_mode=orc.LIBM)
            out[f"ex2_{name}_pg{pg}"] = np.concatenate([ph, gr, [ierr, rc], cnt])
        rc, ph, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, f11, mt, 1, 3, math_mode=orc.LIBM)
        out[f"ex2_{name}_mm3"] = np.concatenate([ph, gr, [ierr, rc], cnt])
    # generate_model-family columns
    grid = synth.make_grid(6, 5, 40)
    pts, par = synth.generate_model(grid, 60, 4242)
    r = orc.forward_eval(pts, par, grid, synth.freqs(20), math_mode=orc.LIBM, phaseGroup=1)
    out["gm_points"] = pts
    out["gm_params"] = par
    out["gm_pvel"] = r["pvel"]
    out["gm_gvel"] = r["gvel"]
    out["gm_ierr"] = r["ierr"]
    out["gm_counters"] = r["counters"]
    path = os.path.join(ROOT, "tests", "golden", "dispersion_oracle.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
