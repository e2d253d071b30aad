"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs (SURVEY.md §8d correctness gates).

 L1  per-step conditional allocation probabilities, identical frozen state: <= 1e-10 relative
     on every probability > 1e-290 and <= 1e-9 absolute on the raw log-weights
 L2  lock-step sweeps under the shared Philox contract: labels / counts / K bit-identical
 +   predictive step: pipred <= 1e-10 relative, drawn cluster and ypred identical
"""
import numpy as np
import pytest

import oracle_binding as ob
from cases import CASES, build_case, occupied_view
from treatppmx_b200 import _lib

pytestmark = pytest.mark.gpu

REL_TOL = 1e-10     # stated tolerance of north_star for fp64 conditional probabilities
LOGW_ATOL = 1e-9    # raw log-weights are O(1e2..1e3); 1e-9 absolute ~ 1e-12 relative


@pytest.fixture(scope="module")
def ctx():
    from treatppmx_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _mk(ctx, name, n_chains=1, n_datasets=1, kcap=0, seed=11):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=n_datasets, kcap=kcap)
    cds = [c % n_datasets for c in range(n_chains)]
    batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    return batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds


def _assert_state_equal(g, o, dims, exact_float=False):
    assert (g.nclu == o.nclu).all()
    nts = [int((np.asarray(o.labels[t]) > 0).sum()) for t in range(dims.T)]
    assert (g.labels == o.labels).all()
    gv, ov = occupied_view(g, dims), occupied_view(o, dims)
    for k in gv:
        if k.startswith(("nj", "nclu")):
            assert (gv[k] == ov[k]).all(), k
        elif k.startswith("sumx"):
            # same operations in the same order -> identical bits
            assert (gv[k] == ov[k]).all(), k
        else:
            np.testing.assert_allclose(gv[k], ov[k], rtol=1e-12, atol=1e-13, err_msg=k)


@pytest.mark.parametrize("name", list(CASES))
def test_init_matches_oracle(ctx, name):
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, name, n_chains=2)
    batch.init(5, labels, nclu, betas[0])
    for c in range(2):
        o = pbs[0].init_state(5, c, labels, nclu, betas[0])
        g = batch.get_state(c)
        _assert_state_equal(g, o, dims)
        np.testing.assert_allclose(g.ss, o.ss, rtol=1e-13)
        np.testing.assert_allclose(g.eta_star, o.eta_star, rtol=0, atol=0)
        np.testing.assert_allclose(g.eta_empty, o.eta_empty, rtol=0, atol=0)
        np.testing.assert_allclose(g.loggamma, o.loggamma, rtol=0, atol=0)
        assert (g.JJ == o.JJ).all()


def _check_scores(batch, pb, st, dims, model, chain, picks):
    worst = 0.0
    for (arm, it) in picks:
        lw, pr, K = batch.score_patient(chain, arm, it, seed=3, iter_=9)
        lw2, pr2, K2 = pb.score_patient(st, arm, it, seed=3, chain_id=chain, iter_=9)
        assert K == K2
        np.testing.assert_allclose(lw, lw2, rtol=0, atol=LOGW_ATOL)
        m = pr2 > 1e-290
        rel = np.abs(pr[m] - pr2[m]) / pr2[m]
        worst = max(worst, rel.max())
        assert rel.max() <= REL_TOL, (arm, it, rel.max())
        assert abs(pr.sum() - 1) < 1e-12
    return worst


@pytest.mark.parametrize("name", list(CASES))
def test_L1_score_patient(ctx, name):
    """Frozen states taken after 0, 3 and 25 oracle sweeps; every patient of every arm."""
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, name)
    pb = pbs[0]
    st = pb.init_state(21, 0, labels, nclu, betas[0])
    n_a = np.bincount(dsets[0].treat, minlength=dims.T)
    done = 0
    for nsw in (0, 3, 22):
        pb.sweep(st, 21, 0, done, nsw)
        done += nsw
        batch.set_state(0, st)
        picks = [(t, it) for t in range(dims.T) for it in range(0, n_a[t], 1 if nsw else 3)]
        _check_scores(batch, pb, st, dims, model, 0, picks)
        # scoring must not change the chain
        g = batch.get_state(0)
        _assert_state_equal(g, st, dims)


def test_L1_edge_partitions(ctx):
    """K_t = 1 (everyone together), K_t = n_t (all singletons: >32 clusters, multi-chunk path),
    and a singleton in the middle of the label range (swap-with-last path)."""
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, "default_ngg")
    pb = pbs[0]
    n_a = np.bincount(dsets[0].treat, minlength=dims.T)
    one = np.zeros_like(labels)
    alls = np.zeros_like(labels)
    for t in range(dims.T):
        one[t, : n_a[t]] = 1
        alls[t, : n_a[t]] = np.arange(1, n_a[t] + 1)
    for lab, ncl in ((one, np.ones(dims.T, np.int32)), (alls, n_a.astype(np.int32))):
        st = pb.init_state(2, 0, lab, ncl, betas[0])
        batch.set_state(0, st)
        picks = [(t, it) for t in range(dims.T) for it in (0, 1, n_a[t] // 2, n_a[t] - 1)]
        _check_scores(batch, pb, st, dims, model, 0, picks)
    mid = labels.copy()
    ncl = nclu.copy()
    for t in range(dims.T):  # patient 7 becomes a singleton with label 2 (others shifted)
        mid[t, : n_a[t]] = np.where(mid[t, : n_a[t]] >= 2, mid[t, : n_a[t]] + 1, mid[t, : n_a[t]])
        mid[t, 7] = 2
        ncl[t] += 1
    st = pb.init_state(2, 0, mid, ncl, betas[0])
    batch.set_state(0, st)
    _check_scores(batch, pb, st, dims, model, 0, [(t, 7) for t in range(dims.T)])


@pytest.mark.parametrize("name", list(CASES))
def test_L2_lockstep_sweeps(ctx, name):
    """Device and oracle run the same chains under the same Philox streams; integer
    bookkeeping must be bit-identical, sufficient statistics too (same operation order)."""
    nchains, nsw = 3, 60
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, name, n_chains=nchains, n_datasets=2)
    batch.init(77, labels, nclu, betas[0])
    sts = [pbs[cds[c]].init_state(77, c, labels, nclu, betas[0]) for c in range(nchains)]
    it0 = 0
    for chunk in (1, 9, nsw - 10):
        batch.sweep(77, it0, chunk)
        for c in range(nchains):
            pbs[cds[c]].sweep(sts[c], 77, c, it0, chunk)
            _assert_state_equal(batch.get_state(c), sts[c], dims)
        it0 += chunk
    g = batch.get_state(0)
    np.testing.assert_allclose(g.loggamma, sts[0].loggamma, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(g.eta_empty, sts[0].eta_empty, rtol=1e-12, atol=1e-13)


def test_L2_long_lockstep_1000_sweeps(ctx):
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, "default_ngg", n_chains=2)
    batch.init(123, labels, nclu, betas[0])
    batch.sweep(123, 0, 1000)
    for c in range(2):
        st = pbs[0].init_state(123, c, labels, nclu, betas[0])
        pbs[0].sweep(st, 123, c, 0, 1000)
        _assert_state_equal(batch.get_state(c), st, dims)


@pytest.mark.parametrize("name", ["default_ngg", "dp_T3", "calib1", "calib2", "categorical", "categorical_T3", "cat_calib2",
                                  "cat_dd_calib1", "nnig", "double_dipper", "ppm_only"])
def test_predict_matches_oracle(ctx, name):
    nchains = 2
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, name, n_chains=nchains)
    batch.init(9, labels, nclu, betas[0])
    batch.sweep(9, 0, 12)
    rng = np.random.default_rng(0)
    for c in range(nchains):
        st = batch.get_state(c)
        # a non-trivial hierarchy so the auxiliary-pick branch (:1369) is exercised
        for t in range(dims.T):
            A = rng.normal(size=(dims.D, dims.D))
            st.Sigma[:, :, t] = A @ A.T + np.eye(dims.D)
            st.Theta[:, t] = rng.normal(size=dims.D)
        batch.set_state(c, st)
    pi, yp, pc = batch.predict(9, 40)
    for c in range(nchains):
        st = batch.get_state(c)
        pi2, yp2, pc2 = pbs[0].predict(st, 9, c, 40)
        assert (pc[c] == pc2).all()
        np.testing.assert_allclose(pi[c], pi2, rtol=REL_TOL, atol=0)
        assert (yp[c] == yp2).all()
        np.testing.assert_allclose(pi[c].sum(axis=1), 1.0, atol=1e-12)


def test_predict_split_between_register_and_staged_kernels(ctx):
    """An arm with more than 16 clusters is scored by predict_fast_kernel, the others by
    predict_reg_kernel, in the same call: both against the oracle."""
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, "default_ngg", n_chains=2)
    n_a = np.bincount(dsets[0].treat, minlength=dims.T)
    lab = np.zeros_like(labels)
    ncl = np.zeros_like(nclu)
    for t, k in enumerate([20, 3][: dims.T]):
        lab[t, : n_a[t]] = (np.arange(n_a[t]) % k) + 1
        ncl[t] = k
    batch.init(21, lab, ncl, betas[0])
    pi, yp, pc = batch.predict(21, 7)
    for c in range(2):
        st = batch.get_state(c)
        assert st.nclu.tolist() == [20, 3]
        pi2, yp2, pc2 = pbs[0].predict(st, 21, c, 7)
        assert (pc[c] == pc2).all()
        np.testing.assert_allclose(pi[c], pi2, rtol=REL_TOL, atol=0)
        assert (yp[c] == yp2).all()


def test_capacity_error_is_reported(ctx):
    batch, pbs, (model, dims, shared, dsets, labels, nclu, betas), cds = _mk(ctx, "dp_T3", kcap=5)
    batch.init(1, labels, nclu, betas[0])
    with pytest.raises(_lib.TppmxError) as ei:
        batch.sweep(1, 0, 200)
    assert ei.value.code == 3


def test_bad_arguments(ctx):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg")
    model.cohesion = 7
    with pytest.raises(_lib.TppmxError):
        Batch(ctx, model, dims, shared, dsets, chain_dataset=[0])
    model.cohesion = 2
    with pytest.raises(_lib.TppmxError):
        Batch(ctx, model, dims, shared, dsets, chain_dataset=[3])
