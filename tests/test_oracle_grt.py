"""The oracle's restatement of the generalized R/T branch (oracle/grt_ref.c) against physics.

The reference ships no value for this branch and cannot be compiled here.  Two kinds of pins: (1) the reference's own
Fortran 90 translated mechanically to C (oracle/f90toc_love.py) -- every routine alone and whole columns end to end, bit for
bit (the second half of this file); (2) physics, independent of any reading of the Fortran: every phase velocity the restatement returns is a zero of an INDEPENDENT secular function (tests/independent_modal.py:
propagator matrices in 50-digit arithmetic, no formula shared with the R/T recursion), that Rayleigh roots of ordinary
crustal models are the lowest mode, that group velocities equal d(omega)/dk of those roots, and that the search is
deterministic in both math modes."""
import numpy as np
import pytest

import independent_modal as im
import oracle_lib as orc

FREQS = np.array([2.0, 1.0, 0.5, 0.333333, 0.25, 0.2, 0.166667, 0.142857, 0.125, 0.111111, 0.1])  # example1's, as shipped


def crust(vs, thick, water=None):
    vs = np.asarray(vs, float)
    vp = 1.73 * vs
    rho = 1.74 * vp ** 0.25
    th = np.asarray(thick, float)
    if water is not None:
        vs = np.concatenate([[0.0], vs]); vp = np.concatenate([[1.5], vp]); rho = np.concatenate([[1.0], rho]); th = np.concatenate([[water], th])
    return th, vp, vs, rho


MODELS = {
    "lvl_mid": crust([3.2, 3.6, 2.9, 3.8, 4.5], [2.0, 3.0, 4.0, 6.0, 0.0]),
    "lvl_two": crust([3.0, 2.6, 3.4, 2.8, 3.9, 4.4], [1.5, 2.0, 3.0, 2.5, 5.0, 0.0]),
    "lvl_thin": crust([2.8, 3.1, 2.5, 3.3, 3.6, 4.2], [0.8, 1.2, 0.6, 3.0, 6.0, 0.0]),
}


def _is_root(f, c, h=3e-5):
    return f(c - h) * f(c + h) < 0


@pytest.mark.parametrize("name", sorted(MODELS))
def test_rayleigh_roots_are_the_fundamental_mode(name):
    th, vp, vs, rho = MODELS[name]
    assert orc.L().orc_nlvls1(orc.f64(vp).ctypes.data, orc.f64(vs).ctypes.data, len(vp), 1) > 0  # takes the GRT branch
    ierr, ph, gr, cnt = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=1, phaseGroup=1, math_mode=orc.LIBM)
    assert ierr == 0 and cnt[0] > 0
    for f, c in zip(FREQS[::2], ph[::2]):
        F = lambda x: im.rayleigh_secular(x, 1 / f, th, vp, vs, rho)
        assert _is_root(F, c), (name, f, c)
        assert im.count_sign_changes(F, 0.6 * vs[vs > 0].min(), c - 1e-4, 60) == 0, "a lower mode exists"
    # group velocity: the reference's finite difference over dh = 0.005 Hz of its own roots
    dh = float(np.float32(0.005))
    for i in (0, 4, 10):
        ierr2, ph2, _, _ = orc.grt_modes(th, vp, vs, rho, np.array([FREQS[i] + dh]), modetype=1, phaseGroup=0, math_mode=orc.LIBM)
        u = dh / ((FREQS[i] + dh) / ph2[0] - FREQS[i] / ph[i])
        assert abs(u - gr[i]) < 2e-2 * gr[i]  # the second search starts from c0 = phase(i) with another tolerance: same mode, ~1e-3


@pytest.mark.parametrize("name", sorted(MODELS))
def test_love_roots_are_zeros_of_the_independent_secular_function(name):
    th, vp, vs, rho = MODELS[name]
    ierr, ph, gr, cnt = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=0, phaseGroup=1, math_mode=orc.LIBM)
    assert ierr == 0
    for f, c, u in zip(FREQS[::2], ph[::2], gr[::2]):
        F = lambda x: im.love_secular(x, 1 / f, th, vs, rho)
        assert _is_root(F, c), (name, f, c)
        assert 0 < u < c + 1e-9


def test_stoneley_branch_under_water():
    th, vp, vs, rho = crust([3.2, 3.6, 2.9, 3.8, 4.5], [2.0, 3.0, 4.0, 6.0, 0.0], water=1.0)
    assert orc.L().orc_nlvls1(orc.f64(vp).ctypes.data, orc.f64(vs).ctypes.data, len(vp), 1) > 0
    fr = FREQS[:6]
    ierr, ph, gr, cnt = orc.grt_modes(th, vp, vs, rho, fr, modetype=1, phaseGroup=0, math_mode=orc.LIBM)
    found = ph < 99
    assert found.any()
    for f, c in zip(fr[found], ph[found]):
        F = lambda x: im.rayleigh_secular(x, 1 / f, th, vp, vs, rho)
        assert _is_root(F, c), (f, c)


def test_math_modes_agree_and_failure_leaves_presets():
    th, vp, vs, rho = MODELS["lvl_mid"]
    a = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=1, phaseGroup=1, math_mode=orc.LIBM)
    b = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=1, phaseGroup=1, math_mode=orc.PORTABLE)
    assert a[0] == b[0] == 0
    assert np.abs(a[1] - b[1]).max() < 1e-5  # roots to tol = 1e-6..1e-5 (north_star's 1e-5 km/s)
    # the reference's own LVL test model (surfmodes/model.dat): the search loses the mode at 0.2 Hz -> ierr = 1, the
    # entries never assigned keep the caller's preset
    m = np.loadtxt(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "grt_model_dat.txt"))
    ierr, ph, gr, _ = orc.grt_modes(m[:, 0], m[:, 1], m[:, 2], m[:, 3], FREQS, modetype=1, phaseGroup=1, math_mode=orc.LIBM, preset=1000.0)
    assert ierr == 1 and (ph[:5] < 10).all() and (ph[5:] == 1000.0).all()


def _refine_root(F, c, h=2e-4, it=40):
    """bisection of the independent secular function in a bracket around the oracle's root"""
    import mpmath as mp
    a, b = mp.mpf(c - h), mp.mpf(c + h)
    fa = F(a)
    assert fa * F(b) < 0
    for _ in range(it):
        m = (a + b) / 2
        fm = F(m)
        if fa * fm < 0:
            b = m
        else:
            a, fa = m, fm
    return float((a + b) / 2)


@pytest.mark.parametrize("modetype", [1, 0])
def test_group_velocity_is_the_finite_difference_of_true_roots(modetype):
    """CalGroup (surfmodes.f90:308-320): U = dh / ((f+dh)/c(f+dh) - f/c(f)), dh = 0.005 Hz.  The same quotient formed from
    roots of the INDEPENDENT secular function (refined to 1e-15) must agree with the restatement's group velocity to the
    accuracy its own root tolerance allows (1e-6 km/s in c -> ~2e-4 relative in U at these frequencies)."""
    th, vp, vs, rho = MODELS["lvl_mid"]
    fr = FREQS[[1, 4, 8]]
    ierr, ph, gr, _ = orc.grt_modes(th, vp, vs, rho, fr, modetype=modetype, phaseGroup=1, math_mode=orc.LIBM)
    assert ierr == 0
    dh = float(np.float32(0.005))
    for f, c, u in zip(fr, ph, gr):
        if modetype == 1:
            F0 = lambda x: im.rayleigh_secular(x, 1 / f, th, vp, vs, rho)
            F1 = lambda x: im.rayleigh_secular(x, 1 / (f + dh), th, vp, vs, rho)
        else:
            F0 = lambda x: im.love_secular(x, 1 / f, th, vs, rho)
            F1 = lambda x: im.love_secular(x, 1 / (f + dh), th, vs, rho)
        c0 = _refine_root(F0, c)
        # the second search of the reference starts from c(f): its root is the same mode a little lower
        c1 = _refine_root(F1, c0 - (c0 / u - 1) * c0 * dh / f if u > 0 else c0, h=2e-3)
        assert abs(c0 - c) < 3e-6, (f, c, c0)
        u_ind = dh / ((f + dh) / c1 - f / c0)
        assert abs(u_ind - u) < 2e-3 * u, (f, u, u_ind)


def test_complex_primitives_are_the_compilers_fortran_rules(tmp_path):
    """The R/T restatement spells complex arithmetic out by hand (oracle/grt_ref.c: cmul, cdiv, g_cexp, csq).  gfortran
    expands COMPLEX*16 multiplication and division in GCC's middle end under its Fortran rules (-fcx-fortran-rules: the plain
    product, the range-reduced quotient, no NaN recovery) and calls libm's cexp / csqrt; gcc is here and shares that middle
    end, so a C file using `double _Complex` compiled with that flag IS the reference's arithmetic.  Bit-identical over random
    operands of every magnitude mix."""
    import ctypes as C
    import subprocess
    src = tmp_path / "cx.c"
    src.write_text("""
#include <complex.h>
void cx_prim(int op, double are, double aim, double bre, double bim, double* out) {
  double _Complex a = are + aim * I, b = bre + bim * I, r;
  if (op == 0) r = a * b; else if (op == 1) r = a / b; else if (op == 2) r = cexp(a);
  else { double t = are / bre; r = csqrt((double _Complex)(1 - t * t)); }
  out[0] = creal(r); out[1] = cimag(r);
}
""")
    so = tmp_path / "libcx.so"
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-fcx-fortran-rules", "-ffp-contract=off", "-o", str(so), str(src), "-lm"])
    ref = C.CDLL(str(so)).cx_prim
    got = orc.L().orc_grt_cprim
    for fn in (ref, got):
        fn.argtypes = [C.c_int] + [C.c_double] * 4 + [C.c_void_p]
        fn.restype = None
    rng = np.random.default_rng(11)
    a, b = np.zeros(2), np.zeros(2)
    n = 0
    for k in range(40000):
        op = k % 4
        sc = 10.0 ** rng.integers(-8, 9, 4) if k % 3 else np.ones(4)
        v = rng.standard_normal(4) * sc
        if op == 2:
            v[0] = np.clip(v[0], -600, 600)
            if k % 40 == 2:
                v[1] = 0.0                                   # exp of a real
        if op == 3:
            v[0], v[2] = abs(v[0]) + 1e-3, abs(v[2]) + 1e-3  # csq(c, vel): positive velocities
        if k % 50 == 1 and op < 2:
            v[rng.integers(0, 4)] = 0.0
        ref(op, *v, a.ctypes.data)
        got(op, *v, b.ctypes.data)
        assert a.tobytes() == b.tobytes() or (np.isnan(a).any() and np.isnan(b).any()), (op, v, a, b)
        n += 1
    assert n == 40000


@pytest.mark.skipif(not orc.have_love_reference(), reason="oracle/_ref/liblove_f2c.so not built (needs /root/reference)")
def test_love_secular_function_equals_the_translated_reference():
    """The Love secular function of the R/T branch (Love.f90: EinvE_L, propup_L, SecFuns_L, with csq and T_GRT of GRT.f90) is
    the one part of those 3 000 lines inside the reach of a mechanical translation (oracle/f90toc_love.py: complex arithmetic
    under gcc's Fortran rules, array sections / constructors / MATMUL scalarised).  The restatement's secfun_L must equal it
    BIT FOR BIT -- value and Imf -- over random low-velocity columns, with and without a water layer, at every frequency and
    over the whole range of trial velocities the search scans (evanescent and propagating layers, the deep-layer cut-off of
    startl)."""
    rng = np.random.default_rng(orc.live_seed("test_love_secular_function_equals_the_translated_reference"))
    n = 0
    cols = [MODELS[k] for k in sorted(MODELS)]
    for _ in range(40):
        nl = int(rng.integers(4, 14))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[k - 1] * rng.uniform(0.7, 0.95)                      # a low-velocity layer
        th = np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0)
        cols.append(crust(vs, th, water=float(rng.uniform(0.2, 3.0)) if rng.random() < 0.3 else None))
    for th, vp, vs, rho in cols:
        lo, hi = 0.8 * vs[vs > 0].min(), 1.05 * vs.max()
        for f in FREQS[::2]:
            for c in rng.uniform(lo, hi, 25):
                rc, re, im_ = orc.grt_secfun(th, vp, vs, rho, float(f), 0, float(c), math_mode=orc.LIBM)
                assert rc == 0
                v, imf = orc.grt_love_secfun_reference(th, vp, vs, rho, float(f), float(c))
                assert _bits_equal(re, v) and _bits_equal(im_, imf), (vs, th, f, c, (re, im_), (v, imf))
                n += 1
    assert n > 5000


def love_fixture_points():
    """(column, frequency, trial velocity) of the committed fixture tests/golden/grt_love_secfun_ref.npz"""
    rng = np.random.default_rng(77)
    cols = [MODELS[k] for k in sorted(MODELS)]
    cols.append(crust([3.1, 2.7, 3.5, 3.0, 4.0, 4.5], [1.0, 2.0, 2.5, 3.0, 5.0, 0.0], water=1.2))
    for th, vp, vs, rho in cols:
        lo, hi = 0.8 * vs[vs > 0].min(), 1.05 * vs.max()
        for f in FREQS:
            for c in rng.uniform(lo, hi, 12):
                yield th, vp, vs, rho, float(f), float(c)


def test_love_secular_function_reproduces_the_reference_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grt_love_secfun_ref.npz"))["values"]
    got = np.array([orc.grt_secfun(th, vp, vs, rho, f, 0, c, math_mode=orc.LIBM)[1:] for th, vp, vs, rho, f, c in love_fixture_points()])
    assert got.shape == g.shape == (4 * 11 * 12, 2)
    assert all(_bits_equal(a, b) for a, b in zip(got.ravel(), g.ravel())), f"{(got != g).sum()} of {g.size} values differ"


def _bits_equal(a, b):
    """the same double, NaN for NaN; the SIGN OF A ZERO may differ: where every layer is evanescent all arithmetic is real and the
    imaginary part of the secular function is an exact zero whose sign depends on how a real operand meets a complex one
    (the restatement scales, the compiler's expansion of the promoted product adds a 0*x term); the searches read Imf only
    through abs() and squares (util.f90:45,134)"""
    return np.float64(a).tobytes() == np.float64(b).tobytes() or (np.isnan(a) and np.isnan(b)) or (a == 0.0 and b == 0.0)


@pytest.mark.skipif(not orc.have_rayleigh_reference(), reason="oracle/_ref/librayleigh_f2c.so not built (needs /root/reference)")
def test_rayleigh_surface_secular_function_equals_the_translated_reference():
    """The Rayleigh secular function a column without a water layer is searched with -- SecFunSurf over propup, EinvE (the 4 x 4
    MATMUL whose block structure the restatement and the device exploit) and inv2, plus startl's choice of the deepest layer --
    as the reference's own Rayleigh.f90 computes it (oracle/f90toc_love.py: sections, constructors, MATMUL, RESHAPE, pointers to
    sections, array-valued functions scalarised; complex arithmetic under gcc's Fortran rules).  Bit for bit: value, Imf, ll.
    (The water-layer functions and the searches have their own tests below.)"""
    rng = np.random.default_rng(orc.live_seed("test_rayleigh_surface_secular_function_equals_the_translated_reference"))
    cols = [MODELS[k] for k in sorted(MODELS)]
    for _ in range(40):
        nl = int(rng.integers(4, 14))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[k - 1] * rng.uniform(0.7, 0.95)
        th = np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0)
        cols.append(crust(vs, th))
    n = nan = deep = 0
    for th, vp, vs, rho in cols:
        lo, hi = 0.75 * vs.min(), 1.05 * vs.max()
        for f in FREQS[::2]:
            for c in rng.uniform(lo, hi, 25):
                rc, re, im_ = orc.grt_secfun(th, vp, vs, rho, float(f), 1, float(c), math_mode=orc.LIBM)
                assert rc == 0
                v, imf, ll_ref, ll = orc.grt_rayleigh_secfun_reference(th, vp, vs, rho, float(f), float(c))
                assert ll_ref == ll, (vs, th, f, c, ll_ref, ll)
                assert _bits_equal(re, v) and _bits_equal(im_, imf), (vs, th, f, c, (re, im_), (v, imf))
                n += 1
                nan += int(np.isnan(v))
                deep += int(ll < len(th))
    assert n > 5000 and nan < n // 2 and deep > 100, (n, nan, deep)        # startl's cut-off is exercised; most values are finite


def rayleigh_fixture_points():
    rng = np.random.default_rng(78)
    cols = [MODELS[k] for k in sorted(MODELS)]
    cols.append(crust([3.1, 2.7, 3.5, 3.0, 4.0, 4.5], [1.0, 2.0, 2.5, 3.0, 5.0, 0.0]))
    for th, vp, vs, rho in cols:
        lo, hi = 0.75 * vs.min(), 1.05 * vs.max()
        for f in FREQS:
            for c in rng.uniform(lo, hi, 12):
                yield th, vp, vs, rho, float(f), float(c)


def test_rayleigh_surface_secular_function_reproduces_the_reference_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grt_rayleigh_secfun_ref.npz"))["values"]
    got = np.array([orc.grt_secfun(th, vp, vs, rho, f, 1, c, math_mode=orc.LIBM)[1:] for th, vp, vs, rho, f, c in rayleigh_fixture_points()])
    assert got.shape == g.shape == (4 * 11 * 12, 2)
    assert all(_bits_equal(a, b) for a, b in zip(got.ravel(), g.ravel())), f"{(got != g).sum()} of {g.size} values differ"


@pytest.mark.skipif(not (orc.have_rayleigh_reference() and orc.have_love_reference()), reason="oracle/_ref translations not built (needs /root/reference)")
@pytest.mark.parametrize("modetype", [1, 0])
def test_root_refinement_equals_the_translated_reference(modetype):
    """bisecim (util.f90:90-167: bisection whose inverse-interpolation estimate enters the stopping test and the final choice)
    as the reference's own statements, driving the reference's own secular functions, against the restatement's -- the same
    brackets a scan would hand over (neighbouring trial velocities with a sign change), the same smin / tol range the callers
    set.  Root, iq, and the end values: bit for bit."""
    rng = np.random.default_rng(orc.live_seed("test_root_refinement_equals_the_translated_reference"))
    cols = [MODELS[k] for k in sorted(MODELS)]
    for _ in range(12):
        nl = int(rng.integers(4, 12))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[k - 1] * rng.uniform(0.7, 0.95)
        cols.append(crust(vs, np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0)))
    n = found = rejected = 0
    for th, vp, vs, rho in cols:
        for f in FREQS[1::3]:
            grid = np.linspace(0.8 * vs.min(), 0.999 * vs.max(), 60)
            vals = np.array([orc.grt_secfun(th, vp, vs, rho, float(f), modetype, float(c), math_mode=orc.LIBM)[1] for c in grid])
            for i in np.nonzero(vals[:-1] * vals[1:] < 0)[0][:4]:
                smin, tol = float(10.0 ** rng.uniform(-5, -2)), float(10.0 ** rng.uniform(-7, -4))
                a = orc.grt_bisecim("port", th, vp, vs, rho, float(f), modetype, float(grid[i]), float(grid[i + 1]), smin, tol)
                b = orc.grt_bisecim("reference", th, vp, vs, rho, float(f), modetype, float(grid[i]), float(grid[i + 1]), smin, tol)
                assert a[0] == b[0] and all(_bits_equal(x, y) for x, y in zip(a[1:], b[1:])), (vs, f, grid[i], a, b)
                n += 1
                found += int(a[0] == 0)
                rejected += int(a[0] == -1)
    assert n > 100 and found > 50, (n, found, rejected)


@pytest.mark.skipif(not (orc.have_rayleigh_reference() and orc.have_love_reference()), reason="oracle/_ref translations not built (needs /root/reference)")
@pytest.mark.parametrize("modetype", [1, 0])
def test_trial_velocity_lists_equal_the_translated_reference(modetype):
    """C_Interval / C_Interval_L (with N_cf / N_cf_L and util.f90's sort) -- the list of trial phase velocities every frequency's
    scan walks, and im1 -- as the reference's own statements build them, against the restatement's: every entry bit for bit,
    at low frequencies (the dense fill) and high ones (the N_cf-driven subdivision, the layer resonances, two midpoint passes)."""
    rng = np.random.default_rng(orc.live_seed("test_trial_velocity_lists_equal_the_translated_reference"))
    cols = [MODELS[k] for k in sorted(MODELS)]
    for _ in range(10):
        nl = int(rng.integers(4, 12))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[k - 1] * rng.uniform(0.7, 0.95)
        cols.append(crust(vs, np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0)))
    n = dense = fine = 0
    for th, vp, vs, rho in cols:
        for f in FREQS:
            tol = float(10.0 ** rng.uniform(-6, -4))
            a, ia, over = orc.grt_cinterval("port", th, vp, vs, rho, float(f), modetype, tol)
            if over:
                continue                      # the Fortran overruns vvv(20000) there: undefined
            b, ib, _ = orc.grt_cinterval("reference", th, vp, vs, rho, float(f), modetype, tol)
            assert len(a) == len(b) and ia == ib and a.tobytes() == b.tobytes(), (vs, f, len(a), len(b), ia, ib)
            n += 1
            dense += int(f < 0.12)
            fine += int(f >= 0.12)
    assert n > 100 and dense > 10 and fine > 50, (n, dense, fine)


@pytest.mark.skipif(not orc.have_rayleigh_reference(), reason="oracle/_ref/librayleigh_f2c.so not built (needs /root/reference)")
@pytest.mark.parametrize("modetype", [1, 0])
def test_setup_grt_equals_the_translated_reference(modetype):
    """setup_grt (surfmodes.f90:320-450) as the reference's own statements on a T_GRT initialised as init_grt does: the scaled
    rigidities, the sorted velocity list, the low-velocity layers and their order, ifs / nlvl1 / nlvls1 / lvlast / L1, vsy / vs1 /
    vsm / vss1 -- everything the searches read -- against the restatement's, with and without a water layer, one and several
    low-velocity zones."""
    rng = np.random.default_rng(orc.live_seed("test_setup_grt_equals_the_translated_reference"))
    cols = [MODELS[k] for k in sorted(MODELS)]
    for k in range(60):
        nl = int(rng.integers(3, 16))
        vs = np.sort(rng.uniform(2.2, 4.6, nl))
        for _ in range(int(rng.integers(0, 4))):
            j = int(rng.integers(1, nl - 1)) if nl > 2 else 1
            vs[j] = vs[j - 1] * rng.uniform(0.6, 0.97)
        th = np.append(rng.uniform(0.3, 6.0, nl - 1), 0.0)
        cols.append(crust(vs, th, water=float(rng.uniform(0.2, 3.0)) if k % 3 == 0 else None))
    n = water = multi = 0
    for th, vp, vs, rho in cols:
        a = orc.grt_setup("port", th, vp, vs, rho, modetype)
        b = orc.grt_setup("reference", th, vp, vs, rho, modetype)
        assert a[0] == 0 and b[0] == 0
        assert a[1].tobytes() == b[1].tobytes(), ("mu", vs)
        assert a[2].tobytes() == b[2].tobytes(), ("v", vs, a[2], b[2])
        nlv = int(a[4][1])
        assert np.array_equal(a[3][:nlv + 1], b[3][:nlv + 1]), ("lvls", vs, a[3], b[3])
        assert np.array_equal(a[4], b[4]), ("ints", vs, a[4], b[4])
        assert a[5].tobytes() == b[5].tobytes(), ("dbl", vs, a[5], b[5])
        n += 1
        water += int(a[4][0] > 0)
        multi += int(nlv > 1)
    assert n > 60 and water > 10 and multi > 5, (n, water, multi)


@pytest.mark.skipif(not orc.have_love_reference(), reason="oracle/_ref/liblove_f2c.so not built (needs /root/reference)")
def test_love_columns_end_to_end_equal_the_translated_reference():
    """A whole Love column with a low-velocity layer -- what `surfmodes` returns for it -- through the reference's own
    statements: setup_grt, C_Interval_L, FundaMode with its internal `check`, startl, SecFuns_L and bisecim are all translated;
    the driver adds init_grt's allocations, the frequency loop of LoveModes and the five calls of SearchLove's allmodes = 0
    path.  orc_grt_modes (libm math mode, what every GPU test of the branch is held to in portable mode) must return the same
    phase velocities bit for bit and the same ierr, with both parameter sets the reference's callers use."""
    rng = np.random.default_rng(orc.live_seed("test_love_columns_end_to_end_equal_the_translated_reference"))
    cols = [MODELS[k] for k in sorted(MODELS)]
    for k in range(60):
        nl = int(rng.integers(4, 12))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        j = int(rng.integers(1, nl - 1))
        vs[j] = vs[j - 1] * rng.uniform(0.7, 0.95)
        cols.append(crust(vs, np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0), water=float(rng.uniform(0.3, 2.0)) if k % 4 == 0 else None))
    n = fail = 0
    for k, (th, vp, vs, rho) in enumerate(cols):
        par = orc.GRT_PAR_LIKELIHOOD if k % 2 else orc.GRT_PAR_MODELLING
        e0, p0, _, _ = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=0, phaseGroup=0, dc=1e-3, par=par, math_mode=orc.LIBM)
        e1, p1 = orc.grt_love_modes_reference(th, vp, vs, rho, FREQS, dc=1e-3, par=par)
        if e1 == -2:
            continue                          # no low-velocity layer by the reference's predicate: surfdisp96's column
        assert e0 == e1, (vs, e0, e1)
        m = len(FREQS) if e0 == 0 else 0
        assert p0[:m].tobytes() == p1[:m].tobytes(), (vs, th, p0, p1)
        n += 1
        fail += int(e0 == 1)
    assert n >= 25, (n, fail)


def love_fixture_columns():
    for k, name in enumerate(sorted(MODELS)):
        yield (*MODELS[name], orc.GRT_PAR_LIKELIHOOD if k % 2 else orc.GRT_PAR_MODELLING)
    yield (*crust([3.1, 2.7, 3.5, 3.0, 4.0, 4.5], [1.0, 2.0, 2.5, 3.0, 5.0, 0.0], water=1.2), orc.GRT_PAR_LIKELIHOOD)
    yield (*crust([2.9, 3.3, 2.6, 3.6, 3.1, 4.1, 4.6], [0.7, 1.4, 2.2, 1.8, 3.5, 6.0, 0.0]), orc.GRT_PAR_MODELLING)


def test_love_columns_reproduce_the_reference_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grt_love_modes_ref.npz"))["phase"]
    for k, (th, vp, vs, rho, par) in enumerate(love_fixture_columns()):
        ierr, ph, _, _ = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=0, phaseGroup=0, dc=1e-3, par=par, math_mode=orc.LIBM)
        assert ierr == 0 and ph.tobytes() == g[k].tobytes(), (k, ph, g[k])
    assert g.shape == (5, len(FREQS)) and (g < 5).all() and (g > 2).all()


@pytest.mark.skipif(not orc.have_rayleigh_reference(), reason="oracle/_ref/librayleigh_f2c.so not built (needs /root/reference)")
def test_rayleigh_columns_end_to_end_equal_the_translated_reference():
    """A whole Rayleigh column with a low-velocity layer -- what `surfmodes` returns for it -- through the reference's own
    statements: setup_grt, C_Interval, FundaMode (an internal procedure of SearchRayleigh, with CR0_Finder and its internal
    Rayhomo), startl, SecFunSurf and bisecim; with a water layer on top St_Finder (with getSt), StMode, SecFunSt, Stoneley,
    propdn_f, EinvE_f and det3 -- all translated; the driver adds init_grt's allocations, the frequency loop of RayleighModes
    and the calls of SearchRayleigh's allmodes = 0 path.  orc_grt_modes (libm math
    mode) must return the same phase velocities bit for bit and the same ierr, with both parameter sets of the callers."""
    rng = np.random.default_rng(orc.live_seed("test_rayleigh_columns_end_to_end_equal_the_translated_reference"))
    cols = [MODELS[k] for k in sorted(MODELS)]
    for k in range(60):
        nl = int(rng.integers(4, 12))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        j = int(rng.integers(1, nl - 1))
        vs[j] = vs[j - 1] * rng.uniform(0.7, 0.95)
        cols.append(crust(vs, np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0), water=float(rng.uniform(0.3, 2.5)) if k % 3 == 0 else None))
    n = fail = water = 0
    for k, (th, vp, vs, rho) in enumerate(cols):
        par = orc.GRT_PAR_LIKELIHOOD if k % 2 else orc.GRT_PAR_MODELLING
        e0, p0, _, _ = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=1, phaseGroup=0, dc=1e-3, par=par, math_mode=orc.LIBM)
        e1, p1 = orc.grt_rayleigh_modes_reference(th, vp, vs, rho, FREQS, dc=1e-3, par=par)
        if e1 == -2:
            continue                          # no low-velocity layer by the reference's predicate: surfdisp96's column
        assert e0 == e1, (vs, e0, e1)
        m = len(FREQS) if e0 == 0 else 0
        assert p0[:m].tobytes() == p1[:m].tobytes(), (vs, th, p0, p1)
        n += 1
        fail += int(e0 == 1)
        water += int(vs[0] == 0)
    assert n >= 25 and water >= 6, (n, fail, water)


def test_rayleigh_columns_reproduce_the_reference_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grt_rayleigh_modes_ref.npz"))["phase"]
    cols = list(love_fixture_columns()) + [(*crust([3.2, 3.6, 2.9, 3.8, 4.5], [2.0, 3.0, 4.0, 6.0, 0.0], water=0.6), orc.GRT_PAR_MODELLING)]
    for k, (th, vp, vs, rho, par) in enumerate(cols):
        ierr, ph, _, _ = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=1, phaseGroup=0, dc=1e-3, par=par, math_mode=orc.LIBM)
        assert ierr == 0 and ph.tobytes() == g[k].tobytes(), (k, ph, g[k])
    assert g.shape == (6, len(FREQS)) and (g < 5).all() and (g > 1).all()


@pytest.mark.skipif(not orc.have_rayleigh_reference(), reason="oracle/_ref/librayleigh_f2c.so not built (needs /root/reference)")
def test_stoneley_secular_function_equals_the_translated_reference():
    """Columns with a water layer on top: SecFunSt over Stoneley, propdn_f, EinvE_f, propup, EinvE and det3, and startl, as the
    reference's own statements compute them -- value, Imf and ll bit for bit (the sign of an exact zero Imf aside)."""
    rng = np.random.default_rng(orc.live_seed("test_stoneley_secular_function_equals_the_translated_reference"))
    cols = [crust([3.1, 2.7, 3.5, 3.0, 4.0, 4.5], [1.0, 2.0, 2.5, 3.0, 5.0, 0.0], water=1.2)]
    for _ in range(30):
        nl = int(rng.integers(4, 12))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        k = int(rng.integers(1, nl - 1))
        vs[k] = vs[k - 1] * rng.uniform(0.7, 0.95)
        cols.append(crust(vs, np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0), water=float(rng.uniform(0.2, 3.0))))
    n = 0
    for th, vp, vs, rho in cols:
        for f in FREQS[::2]:
            for c in rng.uniform(0.7, 1.05 * vs.max(), 20):
                rc, re, im_ = orc.grt_secfun(th, vp, vs, rho, float(f), 1, float(c), math_mode=orc.LIBM)
                assert rc == 0
                v, imf, ll_ref, ll = orc.grt_stoneley_secfun_reference(th, vp, vs, rho, float(f), float(c))
                assert ll_ref == ll and _bits_equal(re, v) and _bits_equal(im_, imf), (vs, th, f, c, (re, im_), (v, imf))
                n += 1
    assert n > 3000


@pytest.mark.skipif(not (orc.have_rayleigh_reference() and orc.have_love_reference()), reason="oracle/_ref translations not built (needs /root/reference)")
@pytest.mark.parametrize("modetype", [1, 0])
def test_group_velocities_end_to_end_equal_the_translated_reference(modetype):
    """paras%phaseGroup = 1: every frequency is searched a second time at freq + 0.005 Hz from the phase velocity just found and
    CalGroup (surfmodes.f90:296-312, translated too) forms the quotient.  Phase AND group velocities of whole columns, bit for
    bit, with and without a water layer."""
    rng = np.random.default_rng(orc.live_seed(f"group{modetype}"))
    cols = [MODELS[k] for k in sorted(MODELS)] + [crust([3.1, 2.7, 3.5, 3.0, 4.0, 4.5], [1.0, 2.0, 2.5, 3.0, 5.0, 0.0], water=1.2)]
    for k in range(24):
        nl = int(rng.integers(4, 11))
        vs = np.sort(rng.uniform(2.4, 4.6, nl))
        j = int(rng.integers(1, nl - 1))
        vs[j] = vs[j - 1] * rng.uniform(0.7, 0.95)
        cols.append(crust(vs, np.append(rng.uniform(0.5, 6.0, nl - 1), 0.0), water=float(rng.uniform(0.3, 2.5)) if k % 3 == 0 else None))
    ref = orc.grt_rayleigh_modes_reference if modetype == 1 else orc.grt_love_modes_reference
    n = 0
    for k, (th, vp, vs, rho) in enumerate(cols):
        par = orc.GRT_PAR_LIKELIHOOD if k % 2 else orc.GRT_PAR_MODELLING
        e0, p0, g0, _ = orc.grt_modes(th, vp, vs, rho, FREQS, modetype=modetype, phaseGroup=1, dc=1e-3, par=par, math_mode=orc.LIBM)
        e1, p1, g1 = ref(th, vp, vs, rho, FREQS, dc=1e-3, par=par, group=True)
        if e1 == -2:
            continue
        assert e0 == e1, (vs, e0, e1)
        if e0 == 0:
            assert p0.tobytes() == p1.tobytes() and g0.tobytes() == g1.tobytes(), (vs, th, g0, g1)
            assert np.isfinite(g0).all() and (g0 >= 0).all()        # (CalGroup returns 0 where its quotient is not positive)
        n += 1
    assert n >= 12
