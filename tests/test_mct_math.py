"""Accuracy of the portable math functions (mctomo_b200/csrc/mct_math.h) against mpmath."""
import ctypes as C

import mpmath as mp
import numpy as np

import oracle_lib as orc

mp.mp.prec = 200


def _ulps(got, exact):
    return float(abs(mp.mpf(got) - exact) / mp.mpf(float(np.spacing(abs(float(exact))))))


def test_exp_accuracy_and_exact_points():
    L = orc.L()
    rng = np.random.default_rng(1)
    xs = np.concatenate([-rng.uniform(0, 60, 3000), rng.uniform(-700, 700, 1000), -rng.uniform(0, 1e-3, 300)])
    worst = max(_ulps(L.orc_mct_exp(float(x)), mp.exp(mp.mpf(float(x)))) for x in xs)
    assert worst < 1.0
    assert L.orc_mct_exp(0.0) == 1.0 and L.orc_mct_exp(-0.0) == 1.0   # a0 = exp(-0) must be exactly 1
    assert L.orc_mct_exp(-800.0) == 0.0


def test_sincos_accuracy():
    L = orc.L()
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(0, 100, 3000), rng.uniform(0, 1e4, 1000), rng.uniform(0, 1, 500), -rng.uniform(0, 50, 300)])
    s, c = C.c_double(), C.c_double()
    ws = wc = 0.0
    for x in xs:
        L.orc_mct_sincos(float(x), C.byref(s), C.byref(c))
        ws = max(ws, _ulps(s.value, mp.sin(mp.mpf(float(x)))))
        wc = max(wc, _ulps(c.value, mp.cos(mp.mpf(float(x)))))
    assert ws < 1.6 and wc < 1.6
    L.orc_mct_sincos(0.0, C.byref(s), C.byref(c))
    assert s.value == 0.0 and c.value == 1.0


def test_pow025_accuracy():
    L = orc.L()
    xs = np.random.default_rng(3).uniform(0.5, 12, 3000)
    worst = max(_ulps(L.orc_mct_pow025(float(x)), mp.mpf(float(x)) ** mp.mpf(0.25)) for x in xs)
    assert worst < 0.51


def test_log_accuracy():
    L = orc.L()
    L.orc_mct_log.restype = C.c_double
    L.orc_mct_log.argtypes = [C.c_double]
    rng = np.random.default_rng(4)
    xs = np.concatenate([rng.uniform(1e-3, 10, 3000), np.exp(rng.uniform(-700, 700, 1000)), rng.uniform(0.99, 1.01, 500),
                         [1e-310, 5e-324, 2.0, 0.5, np.sqrt(2.0)]])
    worst = max(_ulps(L.orc_mct_log(float(x)), mp.log(mp.mpf(float(x)))) for x in xs)
    assert worst < 1.0
    assert L.orc_mct_log(1.0) == 0.0
