"""Pins the stage-2 oracle (oracle/surfdisp96_ref.c, a hand-written restatement) against the REFERENCE'S OWN
surfmodes/surfdisp96.f:

  * tests/golden/dispersion_ref.npz -- 1039 calls of surfdisp96 / surfdisp_mmodes whose outputs were produced by the
    reference source itself, translated statement by statement to C by oracle/f77toc.py (a generic FORTRAN 77 subset
    translator) and compiled with gcc (tools/make_golden_dispersion_ref.py).  The fixtures travel; these tests run
    everywhere, GPU box included.
  * live, where oracle/_ref/libsurfdisp96_f2c.so exists (build container, and the GPU box since oracle/_ref/ travels):
    fresh random stacks every run.
The restatement runs in LIBM math mode here (sin/cos/exp of the C library, which is what the Fortran binary calls):
every output must be BIT-IDENTICAL -- phase velocity, group velocity, ierr, for Rayleigh and Love, phase and group,
fundamental and overtones, with and without a water layer, including the calls that fail."""
import os

import numpy as np
import pytest

import oracle_lib as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dispersion_ref.npz")


def _cases():
    g = np.load(GOLD)
    for k in range(int(g["n"])):
        m = g[f"{k}_model"].astype(np.float64)
        raylov, igr, nm, ie = (int(v) for v in g[f"{k}_sw"])
        yield k, m, g[f"{k}_freqs"], raylov, igr, nm, float(g[f"{k}_dph"]), g[f"{k}_cp"], g[f"{k}_cg"], ie


def test_restatement_reproduces_the_reference_fixtures_bit_for_bit():
    n = nfail = ngroup = nwater = nover = 0
    for k, m, freqs, raylov, igr, nm, dph, cp, cg, ie in _cases():
        rc, p0, g0, e0, _ = orc.surfmodes(m[0], m[1], m[2], m[3], freqs, raylov, igr, nm, dc=dph, math_mode=orc.LIBM)
        assert rc == 0, f"case {k}: the oracle routes this stack elsewhere (rc {rc})"
        assert e0 == ie, f"case {k}: ierr {e0} vs reference {ie}"
        assert np.array_equal(p0, cp), f"case {k}: phase velocities differ in {(p0 != cp).sum()} of {cp.size} outputs"
        assert np.array_equal(g0, cg), f"case {k}: group velocities differ in {(g0 != cg).sum()} of {cg.size} outputs"
        n += 1
        nfail += ie
        ngroup += igr
        nwater += m[2][0] == 0.0
        nover += nm > 1
    assert n > 1000 and nfail > 100 and ngroup > 300 and nwater > 100 and nover > 200   # the fixture covers every branch


def test_portable_math_mode_is_within_tolerance_of_the_reference():
    """PORTABLE mode (mct_math.h: what the CUDA kernels compute with) against the reference's libm outputs: phase
    velocities within the north-star 1e-5 km/s and float32-identical but for a handful; the reference's group velocity is
    a float32 finite difference that amplifies a one-ulp change of the phase velocity a hundredfold
    (surfdisp96.f:236-239,321), so for it the gate is the mismatch RATE, reported."""
    tot = dp = dg = 0
    worst_p = worst_g = 0.0
    for k, m, freqs, raylov, igr, nm, dph, cp, cg, ie in _cases():
        rc, p0, g0, e0, _ = orc.surfmodes(m[0], m[1], m[2], m[3], freqs, raylov, igr, nm, dc=dph, math_mode=orc.PORTABLE)
        assert e0 == ie, f"case {k}: ierr differs between math modes"
        tot += cp.size
        dp += int((p0 != cp).sum())
        dg += int((g0 != cg).sum())
        worst_p = max(worst_p, float(np.abs(p0 - cp).max()))
        worst_g = max(worst_g, float(np.abs(g0 - cg).max()))
    print(f"portable vs reference(libm): {dp} of {tot} phase and {dg} group outputs differ; max |dc| {worst_p:.3g}, max |dU| {worst_g:.3g} km/s")
    assert worst_p <= 1e-5
    assert dp <= 2e-4 * tot and dg <= 2e-4 * tot and worst_g <= 5e-4


@pytest.mark.skipif(not orc.have_f2c(), reason="oracle/_ref/libsurfdisp96_f2c.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(4))
def test_restatement_against_the_translated_reference_live(seed):
    rng = np.random.default_rng(seed)
    n = 0
    for it in range(150):
        nl = int(rng.integers(1, 16))
        vs = np.sort(rng.uniform(1.0, 5.5, nl))
        th = rng.uniform(0.1, 4.0, nl)
        th[-1] = 0.0
        vp = vs * rng.uniform(1.6, 1.9)
        rho = 1.74 * vp ** 0.25
        if it % 4 == 0:
            th, vp, vs, rho = (np.concatenate([[a], b]) for a, b in ((rng.uniform(0.1, 2), th), (1.5, vp), (0.0, vs), (1.0, rho)))
        th, vp, vs, rho = (a.astype(np.float32).astype(np.float64) for a in (th, vp, vs, rho))
        freqs = 1.0 / np.sort(rng.uniform(0.3, 30.0, int(rng.integers(1, 25))))
        raylov = int(rng.integers(0, 2))
        igr, nm = int(rng.integers(0, 2)), int(rng.choice([0, 1, 2, 4]))
        dph = float(rng.choice([1e-3, 2e-3, 5e-4]))
        rc, p0, g0, e0, _ = orc.surfmodes(th, vp, vs, rho, freqs, raylov, igr, nm, dc=dph, math_mode=orc.LIBM)
        if rc != 0:
            continue
        cp, cg, ie = orc.f2c_surfdisp(th, vp, vs, rho, freqs, 2 if raylov else 1, igr, nm, dph)
        assert e0 == ie and np.array_equal(p0, cp) and np.array_equal(g0, cg), (seed, it)
        n += 1
    assert n > 100
