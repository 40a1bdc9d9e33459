#!/usr/bin/env python
"""Generates tests/golden/dispersion_ref.npz from the REFERENCE'S OWN surfdisp96.f.

Runs in the build container only: needs oracle/_ref/libsurfdisp96_f2c.so, i.e. /root/reference/surfmodes/surfdisp96.f
translated statement by statement by oracle/f77toc.py and compiled with gcc (oracle/build_ref.sh).  Each case stores the
layered model (already narrowed to float32, as surfmodes.f90:81-83 hands it over), the frequencies, the call's switches
and what surfdisp96 / surfdisp_mmodes returned (cp, cg, ierr).  The fixtures travel to the GPU box, where neither the
reference nor the translated library need exist: tests/test_oracle_vs_reference.py and tests/test_gpu_reference.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc  # noqa: E402


def random_stack(rng, water=False, nmax=14):
    n = int(rng.integers(1, nmax))
    vs = np.sort(rng.uniform(1.2, 5.2, n))
    th = rng.uniform(0.15, 3.5, n)
    th[-1] = 0.0
    vp = vs * float(np.float32(1.730))
    rho = float(np.float32(1.74)) * vp ** 0.25
    if water:
        th = np.concatenate([[rng.uniform(0.2, 3.0)], th])
        vp = np.concatenate([[1.5], vp])
        vs = np.concatenate([[0.0], vs])
        rho = np.concatenate([[1.0], rho])
    f32 = lambda a: a.astype(np.float32).astype(np.float64)  # noqa: E731
    return f32(th), f32(vp), f32(vs), f32(rho)


def main():
    assert orc.have_f2c(), "run oracle/build_ref.sh first"
    rng = np.random.default_rng(20261018)
    cases = []
    freq_sets = [1.0 / np.linspace(0.5, 10.0, 12), np.array([2.0, 1.0, 0.5, 0.333333, 0.25, 0.2, 0.166667, 0.142857, 0.125, 0.111111, 0.1]),
                 1.0 / np.linspace(1.0, 40.0, 9)]
    for it in range(260):
        water = it % 5 == 4
        th, vp, vs, rho = random_stack(rng, water)
        freqs = freq_sets[it % 3]
        for raylov, iwave in ((1, 2), (0, 1)):
            for igr in (0, 1):
                for nm in (0, 1, 3):
                    if (it + raylov + igr + nm) % 3:     # a third of the combinations per stack
                        continue
                    dph = 1e-3 if it % 7 else 5e-4
                    cp, cg, ie = orc.f2c_surfdisp(th, vp, vs, rho, freqs, iwave, igr, nm, dph)
                    cases.append((th, vp, vs, rho, freqs, raylov, igr, nm, dph, cp, cg, ie))
    out = {"n": np.array(len(cases))}
    for k, (th, vp, vs, rho, fr, raylov, igr, nm, dph, cp, cg, ie) in enumerate(cases):
        out[f"{k}_model"] = np.stack([th, vp, vs, rho]).astype(np.float32)
        out[f"{k}_freqs"] = fr
        out[f"{k}_sw"] = np.array([raylov, igr, nm, ie], np.int32)
        out[f"{k}_dph"] = np.array(dph)
        out[f"{k}_cp"] = cp
        out[f"{k}_cg"] = cg
    path = os.path.join(ROOT, "tests", "golden", "dispersion_ref.npz")
    np.savez_compressed(path, **out)
    fails = sum(int(c[-1]) for c in cases)
    print(f"wrote {path}: {len(cases)} cases ({fails} with ierr = 1), {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
