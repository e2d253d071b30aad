"""The prior-only sampler (reference prior_ppmx_core, src/prior_ppmx.cpp) on the device:
(1) lock-step with the oracle's likelihood-free sweep, (2) the distribution-level known answer
of SURVEY §8c(iv): with PPMx = 0, DP cohesion and CC = 1 the sweep is an exact CRP(alpha) Gibbs
sampler, so the number of clusters follows P(K=k) = |s(n,k)| alpha^k Gamma(alpha)/Gamma(alpha+n)."""
import numpy as np
import pytest
from scipy.special import gammaln

import oracle_binding as ob
from treatppmx_b200 import datagen
from treatppmx_b200._abi import default_model

pytestmark = pytest.mark.gpu


def _crp_count_law(n, alpha):
    ls = np.full((n + 1, n + 1), -np.inf)      # log |s(m, k)|, unsigned Stirling numbers of the first kind
    ls[0, 0] = 0.0
    for m in range(1, n + 1):
        for k in range(1, m + 1):
            a = ls[m - 1, k - 1]
            b = ls[m - 1, k] + np.log(m - 1) if m > 1 else -np.inf
            ls[m, k] = np.logaddexp(a, b)
    k = np.arange(1, n + 1)
    return np.exp(ls[n, 1:] + k * np.log(alpha) + gammaln(alpha) - gammaln(alpha + n))


def test_device_sweep_is_crp_gibbs_sampler():
    from treatppmx_b200.engine import Batch, Context
    n, alpha, nchains, burn, keep = 12, 1.3, 512, 60, 120
    dims_kw, dsets, ex = datagen.genmech_synthetic(n=n, npred=0, P=2, Q=2, T=1, seed=5)
    model = default_model(PPMx=0, cohesion=1, CC=1, nolik=1, alpha=alpha)
    dims = datagen.make_dims(G=100, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=1)
    labels, nclu = datagen.initial_partition(dsets[0].treat, 1, dims.ntmax, nclu_init=3)
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0] * nchains)
    b.init(11, labels, nclu, np.zeros(dims.Q * dims.D))
    b.sweep(11, 0, burn)
    counts = np.zeros(n + 1)
    for s in range(keep):
        b.sweep(11, burn + s, 1)
        _, K = b.get_partitions()
        counts += np.bincount(K[:, 0], minlength=n + 1)
    emp = counts[1:] / counts.sum()
    law = _crp_count_law(n, alpha)
    ek = (alpha / (alpha + np.arange(n))).sum()
    assert (emp * np.arange(1, n + 1)).sum() == pytest.approx(ek, abs=0.03)
    assert np.abs(emp - law).max() < 0.008
    b.close()
    ctx.close()


@pytest.mark.parametrize("cohesion", [1, 2])
def test_prior_ppmx_core_matches_oracle(cohesion):
    from treatppmx_b200.engine import Context
    from treatppmx_b200.prior import prior_ppmx_core
    n, P, iters, burn, thin, seed = 30, 3, 40, 10, 2, 17
    rng = np.random.default_rng(2)
    x = rng.standard_normal((n, P))
    sigma, kappa = 0.2, 1.5
    V = np.zeros((n, n))
    for r in (n - 2, n - 1):
        V[r, : r + 1] = np.exp(datagen.ngg_logV(r + 1, r + 1, sigma, kappa))
    curr = (np.arange(n) % 4) + 1
    sp = (0.0, 10.0, 0.5, 1.0, 2.0, 0.0, 0.1)
    ctx = Context(0)
    got = prior_ppmx_core(ctx, iters, burn, thin, n, 1, P, 0, 1.0, sigma, V, cohesion, 1, 1, 1, 0, 1, x.ravel(), sp, curr,
                          np.bincount(curr)[1:], 4, seed=seed)
    ctx.close()
    # the same run from the oracle's sweep
    model = default_model(PPMx=1, cohesion=cohesion, CC=1, nolik=1, alpha=1.0, similparam=sp)
    dims = datagen.make_dims(G=1, kcap=0, nobs=n, T=1, D=1, Q=0, P=P, ncat=0, maxC=0, npred=0, ntmax=n)
    logV = None
    if cohesion == 2:
        logV = np.full((1, 2, n, 1), -np.inf)
        with np.errstate(divide="ignore"):
            logV[0, 0, :, 0], logV[0, 1, :, 0] = np.log(V[n - 2]), np.log(V[n - 1])
    from treatppmx_b200._abi import HostDataset, HostShared
    shared = HostShared(catvec=np.zeros(1, np.int32), grid=np.array([[sigma], [0.0]]), logV=logV)
    ds = HostDataset(treat=np.zeros(n, np.int32), y=np.ones((n, 1)), z=np.zeros((n, 0)), xcon=x,
                     xcat=np.zeros((n, 1), np.int32), zpred=np.zeros((1, 0)), xconp=np.zeros((1, P)),
                     xcatp=np.zeros((1, 1), np.int32))
    pb = ob.Problem(model, dims, shared, ds)
    st = pb.init_state(seed, 0, curr.reshape(1, n).astype(np.int32), np.array([4], np.int32), np.zeros(0))
    want_K, want_nj = [], []
    for l in range(iters):
        pb.sweep(st, seed, 0, l, 1)
        if l > burn - 1 and (l + 1) % thin == 0:
            want_K.append(int(st.nclu[0]))
            row = np.zeros(n)
            row[: st.nclu[0]] = st.nj[0, : st.nclu[0]]
            want_nj.append(row)
    assert got["nclu"].tolist() == want_K
    np.testing.assert_array_equal(got["njout"], np.array(want_nj).T)
