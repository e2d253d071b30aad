"""The branch-free fp64 log / exp / lgamma of treatppmx_b200/csrc/fastmath.cuh.

CPU part: the same formulas restated in numpy (float32 reciprocal seed + two Newton steps in
place of rcp.approx.ftz.f64) against scipy — pins the algorithm and its stated error bounds.
GPU part: the device functions themselves through tppmx_probe_math."""
import numpy as np
import pytest
from scipy.special import gammaln

LN2_HI, LN2_LO = 0.693147180369123816490, 1.90821492927058770002e-10


def _rcp(a):
    r = (1.0 / a.astype(np.float32)).astype(np.float64)
    for _ in range(2):
        r = r + r * (1.0 - a * r)
    return r


def np_fast_log(y):
    m, e = np.frexp(y)          # m in [0.5, 1)
    m, e = m * 2.0, e - 1       # m in [1, 2)
    big = m > 1.4142135623730951
    m = np.where(big, m * 0.5, m)
    e = e + big
    num, den = m - 1.0, m + 1.0
    r = _rcp(den)
    f = num * r
    f = f + (num - f * den) * r
    f2 = f * f
    p = np.full_like(f, 1.0 / 19.0)
    for c in (17, 15, 13, 11, 9, 7, 5, 3):
        p = p * f2 + 1.0 / c
    t = f * f2 * p
    return e * LN2_HI + (e * LN2_LO + 2.0 * (f + t))


def np_fast_exp(x):
    n = np.rint(x * 1.4426950408889634)
    r = (x - n * LN2_HI) - n * LN2_LO
    r2 = r * r
    fact = [1.0]
    for k in range(1, 14):
        fact.append(fact[-1] * k)
    pa = np.full_like(r, 1.0 / fact[12])
    for k in (10, 8, 6, 4, 2, 0):
        pa = pa * r2 + 1.0 / fact[k]
    pb = np.full_like(r, 1.0 / fact[13])
    for k in (11, 9, 7, 5, 3, 1):
        pb = pb * r2 + 1.0 / fact[k]
    return np.where(x < -700.0, 0.0, np.ldexp(pb * r + pa, n.astype(np.int64)))


def np_lgamma_pos(x):
    small = x < 8.0
    p = x * (x + 1.0) * (x + 2.0) * (x + 3.0) * (x + 4.0) * (x + 5.0) * (x + 6.0) * (x + 7.0)
    y = np.where(small, x + 8.0, x)
    r = _rcp(y)
    r2 = r * r
    s = np.full_like(r, 1.0 / 156.0)
    for c in (-691.0 / 360360.0, 1.0 / 1188.0, -1.0 / 1680.0, 1.0 / 1260.0, -1.0 / 360.0, 1.0 / 12.0):
        s = s * r2 + c
    lg = (y - 0.5) * np_fast_log(y) - y + 0.91893853320467274178 + s * r
    return lg - np_fast_log(np.where(small, p, 1.0))


def _inputs():
    rng = np.random.default_rng(3)
    xl = np.exp(rng.uniform(-30, 30, 200000))
    xe = -rng.uniform(0, 700, 200000)
    xg = np.concatenate([np.exp(rng.uniform(-12, 12, 200000)), [0.5, 1.0, 1.5, 2.0, 3.0, 7.999999, 8.0, 100.0]])
    return xl, xe, xg


def test_numpy_restatement_meets_the_stated_bounds():
    xl, xe, xg = _inputs()
    assert np.max(np.abs(np_fast_log(xl) - np.log(xl)) / np.maximum(np.abs(np.log(xl)), 1e-3)) < 1e-15
    assert np.max(np.abs(np_fast_exp(xe) - np.exp(xe)) / np.exp(xe)) < 5e-16
    ref = gammaln(xg)
    assert np.max(np.abs(np_lgamma_pos(xg) - ref) / (1.0 + np.abs(ref))) < 2e-14


@pytest.mark.gpu
def test_device_fastmath_against_scipy():
    from treatppmx_b200.engine import Context
    ctx = Context(0)
    xl, xe, xg = _inputs()
    got = ctx.probe_math("log", xl)
    assert np.max(np.abs(got - np.log(xl)) / np.maximum(np.abs(np.log(xl)), 1e-3)) < 1e-15
    got = ctx.probe_math("exp", xe)
    assert np.max(np.abs(got - np.exp(xe)) / np.exp(xe)) < 5e-16
    assert ctx.probe_math("exp", np.array([-701.0, -1e4, 0.0]))[:2].tolist() == [0.0, 0.0]
    got = ctx.probe_math("lgamma", xg)
    ref = gammaln(xg)
    assert np.max(np.abs(got - ref) / (1.0 + np.abs(ref))) < 2e-14
    ctx.close()
