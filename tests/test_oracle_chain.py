"""Known-answer tests of the oracle's per-iteration updates (oracle/ppmx_chain.c) and of the
incomplete-gamma pair that stands in for R::pgamma / R::qgamma (treatppmx_b200/csrc/incgamma.h).
The reference has no tests for this code; these pins are what makes the oracle trustworthy."""
import ctypes as C

import numpy as np
import pytest
from scipy import special, stats

import oracle_binding as ob
from cases import build_case
from treatppmx_b200._abi import c_dp


def test_gammainc_pair_matches_scipy():
    L = ob.lib()
    rng = np.random.default_rng(1)
    for a in (0.5, 1.0, 2.5, 2.7, 3.5, 8.5, 13.0, 40.0, 64.5):
        xs = np.concatenate([rng.gamma(a, size=200), [1e-8, 1e-3, a, 3 * a + 10, 80.0]])
        got = np.array([L.oracle_gammainc(a, float(x)) for x in xs])
        np.testing.assert_allclose(got, special.gammainc(a, xs), rtol=2e-13, atol=1e-300)
        ps = np.concatenate([rng.uniform(0, 0.999, 200), [1e-10, 1e-4, 0.5, 1 - 1e-5]])
        inv = np.array([L.oracle_gammaincinv(a, float(p)) for p in ps])
        np.testing.assert_allclose(inv, special.gammaincinv(a, ps), rtol=1e-9)  # conditioning ~ 1/(1-p)
        np.testing.assert_allclose(special.gammainc(a, inv), ps, rtol=1e-11, atol=1e-15)


def test_dmvnrm_is_mvn_logpdf():
    L = ob.lib()
    rng = np.random.default_rng(2)
    for n in (1, 3, 5):
        A = rng.normal(size=(n, n))
        S = np.asfortranarray(A @ A.T + np.eye(n))
        x, mu = rng.normal(size=n), rng.normal(size=n)
        got = L.oracle_dmvnrm_log(x.ctypes.data_as(c_dp), mu.ctypes.data_as(c_dp), S.ctypes.data_as(c_dp), n)
        assert abs(got - stats.multivariate_normal(mu, S).logpdf(x)) < 1e-11


def test_bartlett_wishart_moments():
    """E[W] = df S; Var[W_ij] = df (S_ij^2 + S_ii S_jj)."""
    L = ob.lib()
    n, df = 3, 9.5
    A = np.random.default_rng(3).normal(size=(n, n))
    S = np.asfortranarray(A @ A.T + np.eye(n))
    W = np.zeros((n, n), order="F")
    acc, acc2, N = np.zeros((n, n)), np.zeros((n, n)), 20000
    for l in range(N):
        L.oracle_wishart(77, 0, l, S.ctypes.data_as(c_dp), df, n, W.ctypes.data_as(c_dp))
        acc += W
        acc2 += W * W
    mean, var = acc / N, acc2 / N - (acc / N) ** 2
    np.testing.assert_allclose(mean, df * S, rtol=0.03, atol=0.05 * df)
    np.testing.assert_allclose(var, df * (S ** 2 + np.outer(np.diag(S), np.diag(S))), rtol=0.08)
    assert np.allclose(W, W.T)


def _chain(name="default_ngg", n=60, **kw):
    from cases import CASES
    model, dims, shared, dsets, labels, nclu, betas = build_case(name)
    pb = ob.Problem(model, dims, shared, dsets[0])
    st = pb.init_state(3, 0, labels, nclu, betas[0])
    return pb, st, (model, dims, shared, dsets, labels, nclu, betas)


def test_update_keeps_loggamma_consistent():
    """After any number of full iterations loggamma_i == eta*[label_i] + beta'z_i (the invariant
    every MH step of the reference maintains incrementally)."""
    pb, st, (model, dims, shared, dsets, labels, nclu, betas) = _chain()
    acc = pb.new_acc()
    pb.iterate(st, 3, 0, 0, 30, acc)
    z, treat = dsets[0].z, dsets[0].treat
    pos = np.zeros(dims.nobs, dtype=int)
    for t in range(dims.T):
        pos[treat == t] = np.arange((treat == t).sum())
    for i in range(dims.nobs):
        t = treat[i]
        lab = st.labels[t, pos[i]]
        for k in range(dims.D):
            want = st.eta_star[k, lab - 1, t] + sum(st.beta[k + q * dims.D] * z[i, q] for q in range(dims.Q))
            assert abs(st.loggamma[i, k] - want) < 1e-9
    assert acc[0].sum() > 0 and acc[1].sum() > 0          # some eta and beta proposals were accepted
    assert (st.JJ > 0).all() and (st.ss > 0).all() and st.tau[0] > 0
    assert np.isin(st.sigma, shared.grid[0]).all()
    for t in range(dims.T):                                 # Sigma stays symmetric positive definite
        S = st.Sigma[:, :, t]
        assert np.allclose(S, S.T) and (np.linalg.eigvalsh(S) > 0).all()


def test_jj_ss_conditional_moments():
    """JJ_ik | . ~ Gamma(y_ik + exp(loggamma_ik), scale 1/(ss_i + 1))  (:1045)."""
    pb, st, (model, dims, *_rest) = _chain()
    base = st.copy()
    y = _rest[1][0].y
    N = 400
    acc = np.zeros_like(st.JJ)
    for l in range(N):
        s2 = base.copy()
        # only the JJ/ss block: disable everything else
        m2 = type(model).from_buffer_copy(model)
        pb2 = ob.Problem(m2, dims, pb.shared, pb.data)
        pb2.model.upd_hier, pb2.model.noprog = 0, 1
        pb2.c.model = pb2.model
        pb2.model.mhtune_eta = 0.0
        pb2.c.model = pb2.model
        pb2.update_iter(s2, 5, 0, l)
        acc += s2.JJ
    want = (y + np.exp(base.loggamma)) / (base.ss[:, None] + 1.0)
    rel = np.abs(acc / N - want) / want
    assert np.median(rel) < 0.06 and rel.max() < 0.5
