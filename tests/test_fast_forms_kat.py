"""CPU known-answer tests of the ALGEBRA behind the fast sweep kernel's special instantiations
(treatppmx_b200/csrc/sweep_fast.cuh), each identity restated in numpy and checked against the oracle's own
similarity functions (which are pinned against the reference's sources, tests/test_oracle_vs_reference.py):

 (i)   categorical covariates (CAT): gsimcatDM(counts + e_c) - gsimcatDM(counts) = log(n_c + w) - log(n + C w),
       and the auxiliary-cluster value log(w) - log(C w)                                   src/utils_ct.cpp:468-495
 (ii)  double dipper (DDIP): sum_p gsimconNN(DD = 1) from a = sum t^2, st = sum t, with u = 2 t - c0             :382-405
 (iii) calibration 1 (CAL1): lgN / lgY themselves from (P A0[n], H[n], a, sum sumx2)                               :382-405
 (iv)  NNIG row evaluation (NNIG): gsimconNNIG from the per-n table row with quotients from tabulated reciprocals  :427-459
The GPU tests (tests/test_gpu_fast_probs.py) check the kernels themselves; these keep the derivations honest without a GPU."""
import numpy as np
import pytest

import oracle_binding as ob

L = ob.lib()
RNG = np.random.default_rng(11)


def _nn_tab(m0, s20, v2, nmax):
    """(A0, H, B0, H2) per n: treatppmx_b200/csrc/tppmx.cu:build_nn_tab"""
    rows = []
    sd0 = np.sqrt(s20)
    ld2 = -(0.918938533204672741780329736406 + 0.5 * (m0 / sd0) ** 2 + np.log(sd0))
    for n in range(nmax + 2):
        s2s = 1 / (n / v2 + 1 / s20)
        s2ss = 1 / (n / v2 + 1 / s2s)
        c1 = -0.5 * n * np.log(2 * np.pi * v2)
        rows.append((c1 + ld2 + 0.918938533204672741780329736406 + 0.5 * np.log(s2s), 0.5 * s2s,
                     c1 - 0.5 * np.log(s2s) + 0.5 * np.log(s2ss), 0.5 * s2ss))
    return np.array(rows)


@pytest.mark.parametrize("C", [2, 3, 5])
def test_categorical_with_minus_without_is_two_logs(C):
    w = 0.1
    dirw = np.full(C, w)
    for _ in range(20):
        counts = RNG.integers(0, 7, size=C).astype(np.float64)
        if counts.sum() == 0:
            counts[0] = 1
        c = int(RNG.integers(0, C))
        lgN = L.oracle_gsimcatDM(counts.ctypes.data_as(ob.c_dp), dirw.ctypes.data_as(ob.c_dp), C, 0, 1)
        plus = counts.copy()
        plus[c] += 1
        lgY = L.oracle_gsimcatDM(plus.ctypes.data_as(ob.c_dp), dirw.ctypes.data_as(ob.c_dp), C, 0, 1)
        assert lgY - lgN == pytest.approx(np.log(counts[c] + w) - np.log(counts.sum() + C * w), abs=2e-13)
    one = np.zeros(C)
    one[1] = 1
    aux = L.oracle_gsimcatDM(one.ctypes.data_as(ob.c_dp), dirw.ctypes.data_as(ob.c_dp), C, 0, 1)
    assert aux == pytest.approx(np.log(w) - np.log(C * w), abs=2e-13)


@pytest.mark.parametrize("n,P", [(1, 3), (4, 10), (23, 7)])
def test_double_dipper_difference_from_a_b_st(n, P):
    m0, s20, v2 = 0.3, 10.0, 0.5
    tab = _nn_tab(m0, s20, v2, n + 1)
    iv2, c0 = 1 / v2, m0 / s20
    X = RNG.normal(0.2, 1.1, size=(n, P))
    x = RNG.normal(0.0, 1.0, size=P)
    S, SS = X.sum(0), (X * X).sum(0)
    want = sum(L.oracle_gsimconNN(m0, v2, s20, S[p] + x[p], SS[p] + x[p] * x[p], n + 1, 1, 1)
               - L.oracle_gsimconNN(m0, v2, s20, S[p], SS[p], n, 1, 1) for p in range(P))
    t = S * iv2 + c0
    a, b, st, qx, sx = (t * t).sum(), (t * x).sum(), t.sum(), (x * x).sum(), x.sum()
    tY = a + iv2 * (2 * b + iv2 * qx)
    au = 4 * a - 4 * c0 * st + P * c0 * c0
    bu = 2 * b - c0 * sx
    uY = au + 4 * iv2 * (bu + iv2 * qx)
    A0n, Hn, B0n, H2n = tab[n]
    A0y, Hy, B0y, H2y = tab[n + 1]
    got = P * (B0y - B0n) - 0.5 * iv2 * qx - (Hy * tY - Hn * a) + (H2y * uY - H2n * au)
    assert got == pytest.approx(want, rel=1e-11, abs=1e-10)


@pytest.mark.parametrize("n,P", [(1, 4), (6, 10), (31, 5)])
def test_calibration1_needs_lgN_and_lgY_themselves(n, P):
    m0, s20, v2 = -0.1, 4.0, 0.5
    tab = _nn_tab(m0, s20, v2, n + 1)
    iv2, c0 = 1 / v2, m0 / s20
    X = RNG.normal(size=(n, P))
    x = RNG.normal(size=P)
    S, SS = X.sum(0), (X * X).sum(0)
    lgN = sum(L.oracle_gsimconNN(m0, v2, s20, S[p], SS[p], n, 0, 1) for p in range(P))
    lgY = sum(L.oracle_gsimconNN(m0, v2, s20, S[p] + x[p], SS[p] + x[p] ** 2, n + 1, 0, 1) for p in range(P))
    t = S * iv2 + c0
    a, b, qx, ss = (t * t).sum(), (t * x).sum(), (x * x).sum(), SS.sum()
    tY = a + iv2 * (2 * b + iv2 * qx)
    assert P * tab[n][0] - 0.5 * iv2 * ss + tab[n][1] * a == pytest.approx(lgN, rel=1e-12, abs=1e-10)
    assert P * tab[n + 1][0] - 0.5 * iv2 * (ss + qx) + tab[n + 1][1] * tY == pytest.approx(lgY, rel=1e-12, abs=1e-10)


def _div_by_tab(a, d, r):
    """q = a r; q += (a - q d) r with the residual formed exactly (an fma on the device; exact rational arithmetic here)"""
    from fractions import Fraction
    q = a * r
    rem = float(Fraction(a) - Fraction(q) * Fraction(d))   # exact a - q d, rounded once
    return float(Fraction(rem) * Fraction(r) + Fraction(q))  # fma(rem, r, q): one rounding


@pytest.mark.parametrize("n", [1, 2, 9, 40, 76])
def test_nnig_row_evaluation_reproduces_the_oracle_bit_for_bit_up_to_the_log(n):
    """gsim_nnig_row (sim.cuh) with libm's log in place of the device's table-driven one: every other operation is the
    reference's, in its order, so the result must agree with oracle_gsimconNNIG to the rounding of that one log."""
    m0, k0, nu0, s20 = 0.1, 1.0, 2.0, 10.0
    mu, v = 10.0, 0.1
    a0, b0 = 0.5 * nu0, 0.5 * nu0 * s20
    sd0 = np.sqrt(v / k0)
    t0 = (mu - m0) / sd0
    ld2 = -(0.918938533204672741780329736406 + 0.5 * t0 * t0 + np.log(sd0)) + (a0 * np.log(b0) - float(__import__("math").lgamma(a0)) - (a0 + 1) * np.log(v) - (b0 / v))
    import math
    a0s, kn = a0 + 0.5 * n, k0 + n
    sd = math.sqrt(v / kn)
    row = [-0.5 * n * math.log(2 * math.pi * v), float(n), 1 / float(n), kn, 1.0 / kn, 0.5 * n * k0, sd, 1.0 / sd, math.log(sd), a0s,
           math.lgamma(a0s), (a0s + 1) * math.log(v), n * mu * mu, ld2]
    for _ in range(25):
        x = RNG.normal(0.3, 1.3, size=n)
        S, SS = float(x.sum()), float((x * x).sum())
        xbar = S * row[2]
        nx = row[1] * xbar
        m0s = _div_by_tab(k0 * m0 + nx, row[3], row[4])
        assert m0s == (k0 * m0 + nx) / row[3]                      # the corrected quotient IS the correctly rounded one
        dx = xbar - m0
        t3 = _div_by_tab(row[5] * dx * dx, row[3], row[4])
        assert t3 == (row[5] * dx * dx) / row[3]
        b0s = b0 + 0.5 * (SS - nx * xbar) + t3
        ld1 = row[0] - 5.0 * ((SS - 2.0 * S * mu) + row[12])
        t = _div_by_tab(mu - m0s, row[6], row[7])
        assert t == (mu - m0s) / row[6]
        dnl = -((0.918938533204672741780329736406 + 0.5 * t * t) + row[8])
        q10 = _div_by_tab(b0s, 0.1, 10.0)
        assert q10 == b0s / 0.1
        ig = ((row[9] * math.log(b0s) - row[10]) - row[11]) - q10
        got = (ld1 + row[13]) - (dnl + ig)
        want = L.oracle_gsimconNNIG(m0, k0, nu0, s20, S, SS, n, 0, 1)
        assert got == pytest.approx(want, rel=0, abs=4e-13 + 4e-16 * abs(want))
