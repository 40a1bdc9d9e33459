"""The dispersion oracle against an independent multiprecision eigen-solver (tests/independent_modal.py).

The reference ships no golden dispersion values and no Fortran compiler exists in this image, so the
oracle's bit-level behaviour is pinned on a mechanical translation of surfdisp96.f (tests/test_oracle_vs_reference.py), not on
a gfortran binary.  Independent of any reading of the Fortran is that its answers are the roots of the elastic eigenproblem: here every phase
velocity the oracle returns -- Rayleigh and Love, fundamental and overtones, with and without a water
layer -- is required to lie within 1e-5 km/s (the north-star tolerance) of a root of a secular function
derived from the equations of motion alone and evaluated with 50 digits, the number of modes it finds is
required to equal the number of roots of that function, and its group velocity must agree with d(omega)/dk
of those roots."""
import mpmath as mp
import numpy as np
import pytest

import independent_modal as im
import oracle_lib as orc

TOL = 1e-5          # km/s, BASELINE.json north_star
PERIODS = np.array([0.5, 3.0, 15.0])


def _models():
    rng = np.random.default_rng(20261017)
    out = []
    for k in range(4):
        n = int(rng.integers(3, 7))
        vs = np.sort(rng.uniform(1.2, 4.6, n))
        vs[1:] = np.maximum(vs[1:], vs[:-1] + 0.15)        # strictly increasing: stays on the surfdisp96 branch
        vp = 1.73 * vs
        rho = 1.74 * vp ** 0.25
        thick = np.append(rng.uniform(0.4, 3.0, n - 1), 0.0)
        out.append((f"solid{k}", thick, vp, vs, rho))
    # water over solids
    thick, vp, vs, rho = out[0][1:]
    out.append(("water", np.insert(thick, 0, 0.8), np.insert(vp, 0, 1.5), np.insert(vs, 0, 0.0), np.insert(rho, 0, 1.03)))
    return out


MODELS = _models()


def _secular(kind, period, thick, vp, vs, rho):
    if kind == 1:
        return lambda c: im.rayleigh_secular(c, period, thick, vp, vs, rho)
    return lambda c: im.love_secular(c, period, thick, vs, rho)


def _true_root(F, c0, half_width=5e-5):
    fa, fb = F(c0 - half_width), F(c0 + half_width)
    assert fa * fb < 0, "no root of the independent secular function next to the oracle's value"
    return mp.findroot(F, (c0 - half_width, c0 + half_width), solver="illinois", tol=1e-12)


@pytest.mark.parametrize("name,thick,vp,vs,rho", MODELS, ids=[m[0] for m in MODELS])
@pytest.mark.parametrize("kind", [1, 0], ids=["rayleigh", "love"])
def test_fundamental_roots_are_physical(name, thick, vp, vs, rho, kind):
    for mm in (orc.LIBM, orc.PORTABLE):
        rc, ph, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, 1 / PERIODS, kind, 0, 0, math_mode=mm)
        assert rc == 0 and ierr == 0
        if mm == orc.LIBM:
            ph_libm = ph
    assert np.abs(ph - ph_libm).max() <= TOL
    for T, c in zip(PERIODS, ph_libm):
        r = _true_root(_secular(kind, T, thick, vp, vs, rho), c)
        assert abs(r - c) < TOL, (name, kind, T, c, float(r))


@pytest.mark.parametrize("kind", [1, 0], ids=["rayleigh", "love"])
def test_overtone_roots_and_mode_counts(kind):
    name, thick, vp, vs, rho = MODELS[1]
    periods = np.array([0.35, 1.5])
    nm = 4
    rc, ph, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, 1 / periods, kind, 0, nm, math_mode=orc.LIBM)
    assert rc == 0
    ph = ph.reshape(nm, len(periods))                    # mode-major (surfmodes.f90 layout)
    bmax = vs[-1]
    for j, T in enumerate(periods):
        F = _secular(kind, T, thick, vp, vs, rho)
        found = [c for c in ph[:, j] if c > 0]
        lo = 0.8 * vs[vs > 0].min()
        n_true = im.count_sign_changes(F, lo, bmax * (1 - 1e-9), 160)
        assert len(found) == min(nm, n_true), (T, found, n_true)
        assert np.all(np.diff(found) > 0)
        for c in found:
            r = _true_root(F, c)
            assert abs(r - c) < TOL, (T, c, float(r))


@pytest.mark.parametrize("kind", [1, 0], ids=["rayleigh", "love"])
def test_group_velocity_is_domega_dk(kind):
    name, thick, vp, vs, rho = MODELS[2]
    rc, _, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, 1 / PERIODS, kind, 1, 0, math_mode=orc.LIBM)
    assert rc == 0 and ierr == 0
    rc, ph, _, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, 1 / PERIODS, kind, 0, 0, math_mode=orc.LIBM)
    for T, c, u in zip(PERIODS, ph, gr):
        h = 1e-4                                           # centred difference on exact roots
        w1, w2 = 2 * mp.pi / T * (1 - h), 2 * mp.pi / T * (1 + h)
        c1 = _true_root(_secular(kind, 2 * mp.pi / w1, thick, vp, vs, rho), c, 2e-3)
        c2 = _true_root(_secular(kind, 2 * mp.pi / w2, thick, vp, vs, rho), c, 2e-3)
        u_true = (w2 - w1) / (w2 / c2 - w1 / c1)
        assert abs(u_true - u) < 2e-3, (T, u, float(u_true))   # the reference differentiates float32 roots over +-0.5 %
