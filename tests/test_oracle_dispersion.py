"""Known-answer and regression tests of the dispersion oracle (oracle/surfdisp96_ref.c).

The reference has no tests or golden values for this path; the oracle is pinned on the reference's own surfdisp96.f,
translated mechanically to C, in tests/test_oracle_vs_reference.py.  Here: physics (half-space Rayleigh velocity, Love cut-off, monotone dispersion), internal consistency
(libm vs portable math, surfdisp96 vs surfdisp_mmodes on the fundamental) and drift (self-generated
vectors in tests/golden/dispersion_oracle.npz)."""
import os

import numpy as np
import pytest
from scipy.optimize import brentq

import oracle_lib as orc
from mctomo_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dispersion_oracle.npz")
F20 = synth.freqs(20)


def rayleigh_halfspace(vs, vp):
    """Root of the Rayleigh function gtsolh iterates on (surfdisp96.f:711-732)."""
    g = vs / vp

    def f(k):
        return (2 - k * k) ** 2 - 4 * np.sqrt(1 - (g * k) ** 2) * np.sqrt(1 - k * k)
    return vs * brentq(f, 0.5, 0.999)


@pytest.mark.parametrize("vs", [1.0, 3.0, 4.5])
def test_halfspace_rayleigh_velocity(vs):
    vp = 1.73 * vs
    cr = rayleigh_halfspace(vs, vp)
    assert abs(cr / vs - 0.919255) < 2e-5
    for mm in (orc.LIBM, orc.PORTABLE):
        rc, ph, gr, ierr, cnt = orc.surfmodes([0.0], [vp], [vs], [2.5], F20, 1, 1, 0, math_mode=mm)
        assert rc == 0 and ierr == 0
        assert np.abs(ph - cr).max() < 5e-6 * cr       # nevill stops at 1e-6 relative
        assert np.abs(gr - cr).max() < 2e-3             # non-dispersive: U = c (float32 finite difference)


def test_halfspace_love_not_found():
    """No Love wave exists in a half-space: surfdisp96 reports ierr = 1 and zeroes cg (surfdisp96.f:333-376)."""
    rc, ph, gr, ierr, cnt = orc.surfmodes([0.0], [5.0], [3.0], [2.5], F20, 0, 0, 0)
    assert rc == 0 and ierr == 1
    assert (ph == 100.0).all() and (gr == 0.0).all()


def test_layer_over_halfspace_love_limits():
    """Love fundamental: c -> vs2 at long period, c -> vs1 at short period, monotone in between."""
    T = np.geomspace(0.05, 100.0, 40)
    rc, ph, gr, ierr, cnt = orc.surfmodes([2.0, 0.0], [3.5, 7.0], [2.0, 4.0], [2.3, 3.0], 1 / T, 0, 0, 0)
    assert rc == 0 and ierr == 0
    assert np.all(np.diff(ph) >= 0)
    assert abs(ph[0] - 2.0) < 1e-2 and abs(ph[-1] - 4.0) < 1e-2 and ph[-1] < 4.0


def test_portable_math_agrees_with_libm():
    g = np.load(GOLD)
    grid = synth.make_grid(6, 5, 40)
    a = orc.forward_eval(g["gm_points"], g["gm_params"], grid, F20, math_mode=orc.LIBM, phaseGroup=1)
    b = orc.forward_eval(g["gm_points"], g["gm_params"], grid, F20, math_mode=orc.PORTABLE, phaseGroup=1)
    assert np.array_equal(a["ierr"], b["ierr"])
    assert np.abs(a["pvel"] - b["pvel"]).max() <= 1e-5 and np.abs(a["gvel"] - b["gvel"]).max() <= 1e-5
    assert np.array_equal(a["counters"], b["counters"])


def test_regression_vectors():
    g = np.load(GOLD)
    thick, vp, vs, rho = g["ex2_model"]
    f = g["freqs11"]
    for mt, name in ((1, "ray"), (0, "love")):
        for pg in (0, 1):
            rc, ph, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, f, mt, pg, 0, math_mode=orc.LIBM)
            assert np.array_equal(np.concatenate([ph, gr, [ierr, rc], cnt]), g[f"ex2_{name}_pg{pg}"])
        rc, ph, gr, ierr, cnt = orc.surfmodes(thick, vp, vs, rho, f, mt, 1, 3, math_mode=orc.LIBM)
        assert np.array_equal(np.concatenate([ph, gr, [ierr, rc], cnt]), g[f"ex2_{name}_mm3"])
    grid = synth.make_grid(6, 5, 40)
    r = orc.forward_eval(g["gm_points"], g["gm_params"], grid, F20, math_mode=orc.LIBM, phaseGroup=1)
    assert np.array_equal(r["pvel"], g["gm_pvel"]) and np.array_equal(r["gvel"], g["gm_gvel"])
    assert np.array_equal(r["counters"], g["gm_counters"])


def test_mmodes_fundamental_equals_surfdisp96_phase():
    thick = [1.0, 2.0, 4.0, 0.0]
    vs = np.array([2.0, 2.6, 3.3, 4.2])
    vp = 1.73 * vs
    rho = 1.74 * vp ** 0.25
    _, p0, _, e0, _ = orc.surfmodes(thick, vp, vs, rho, F20, 1, 0, 0)
    _, p1, _, e1, _ = orc.surfmodes(thick, vp, vs, rho, F20, 1, 0, 2)
    assert e0 == 0 and np.array_equal(p0, p1[:20])
    found = p1[20:] > 0
    assert found[:3].all()                       # the first overtone exists at short periods ...
    assert np.all(np.diff(found.astype(int)) <= 0)  # ... and once lost it stays lost (ift, surfdisp96.f:566,688)
    assert np.all(p1[20:][found] > p1[:20][found])  # no mode crossing


def test_lvl_model_takes_grt_branch():
    """surfmodes/model.dat (the reference's own LVL example) must be routed away from surfdisp96."""
    m = np.array([[0.1, 2.8343793669911984, 2.5825142775621530, 2.1667060638969216],
                  [1.7, 2.8547849595267989, 1.5121378245836730, 2.2617395930051476],
                  [1.0, 4.6579939060979507, 2.2201301658572934, 2.5562244215165064],
                  [2.1, 4.1008595181156924, 2.3704390016781254, 2.4760988227494232],
                  [0.0, 4.5746158111015758, 2.9196517156147563, 2.5447077199113832]])
    rc, ph, gr, ierr, cnt = orc.surfmodes(m[:, 0], m[:, 1], m[:, 2], m[:, 3], F20, 1, 0, 0)
    assert rc == 2


def test_convert_to_layer_rules():
    vs = np.array([2.0, 2.0, 2.0 + 5e-11, 3.0, 3.0, 4.0])
    vp = 1.73 * vs
    rho = np.arange(6.0) + 1
    n, (th, al, be, rk) = orc.convert_column(vp, vs, rho, dz=0.5)
    # the 5e-11 wiggle is below EPS = 1e-10f: three runs, the last one is the half-space with thick 0
    assert n == 3 and np.allclose(th, [1.5, 1.0, 0.0]) and np.array_equal(be, [2.0, 3.0, 4.0])
    assert rk[0] == 1 and rk[1] == 4 and rk[2] == 6        # first node of a run; bottom cell for the half-space
    n, (th, al, be, rk) = orc.convert_column(vp, vs, rho, dz=0.5, waterDepth=0.7, scaling=2.0)
    assert n == 4 and be[0] == 0 and al[0] == 1.5 and th[0] == 0.7 and np.allclose(th[1:], [0.75, 0.5, 0.0])
    # modelling variant: EPS = 1e-5f merges 3.0 and 3.0 + 5e-6
    vs2 = np.array([2.0, 3.0, 3.0 + 5e-6, 4.0])
    n, _ = orc.convert_column(1.73 * vs2, vs2, vs2, 0.5, layer_eps=float(np.float32(1e-5)), water_thresh=0.0)
    assert n == 3
    n, _ = orc.convert_column(1.73 * vs2, vs2, vs2, 0.5)
    assert n == 4


def test_check_model_and_property_maps():
    grid = synth.make_grid(3, 4, 5)
    vs = np.tile(np.linspace(2, 3, 5), (3, 4, 1))
    assert orc.check_model(vs, grid) == 0
    vs[1, 2, 3] = 1.9
    assert orc.check_model(vs, grid) == 1
    vp, rho = orc.vs2vp_rho(vs, orc.LIBM)
    assert np.array_equal(vp, vs * float(np.float32(1.73)))
    assert np.allclose(rho, float(np.float32(1.74)) * vp ** 0.25, rtol=1e-15)
    vp2, rho2 = orc.vs2vp_rho(vs, orc.PORTABLE)
    assert np.array_equal(vp, vp2) and np.abs(rho - rho2).max() <= np.spacing(rho.max())
