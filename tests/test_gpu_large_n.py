"""BASELINE config 4 shape class (large n, P = 50 predictive biomarkers, NGG cohesion with the
compact log V table, cluster capacity below max n_t): parity against the oracle at a size the
oracle finishes in seconds, and size-independent invariants at the full n = 20 000."""
import numpy as np
import pytest

import oracle_binding as ob
from treatppmx_b200 import datagen
from treatppmx_b200._abi import default_model

pytestmark = pytest.mark.gpu


def _case(n, P, T, kcap, n_datasets=1, seed=20240401, npred=0, **model_kw):
    model = default_model(cohesion=2, **model_kw)
    dims_kw, dsets, ex = datagen.genmech_synthetic(n=n, npred=npred, P=P, Q=2, T=T, n_datasets=n_datasets, seed=seed,
                                                   mixture=8)
    dims = datagen.make_dims(G=100, kcap=kcap, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=2, kmax=kcap)
    labels, nclu = datagen.initial_partition(dsets[0].treat, dims.T, dims.ntmax, nclu_init=5)
    return model, dims, shared, dsets, labels, nclu, datagen.init_beta(dsets[0])


def _invariants(st, dims, ds):
    """labels in 1..K, counts = label histogram, sumx/sumx2 = sums over the members."""
    for t in range(dims.T):
        idx = np.nonzero(ds.treat == t)[0]
        lab = st.labels[t, : len(idx)]
        K = int(st.nclu[t])
        assert lab.min() >= 1 and lab.max() <= K
        assert (np.bincount(lab, minlength=K + 1)[1:] == st.nj[t, :K]).all()
        X = ds.xcon[idx]
        for j in range(K):
            m = lab == j + 1
            np.testing.assert_allclose(st.sumx[t, j * dims.P:(j + 1) * dims.P], X[m].sum(0), rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(st.sumx2[t, j * dims.P:(j + 1) * dims.P], (X[m] ** 2).sum(0), rtol=1e-9, atol=1e-9)


def test_config4_shape_matches_oracle():
    from treatppmx_b200.engine import Batch, Context
    model, dims, shared, dsets, labels, nclu, beta0 = _case(n=1500, P=50, T=3, kcap=96)
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    b.init(5, labels, nclu, beta0)
    pb = ob.Problem(model, dims, shared, dsets[0])
    sts = [pb.init_state(5, c, labels, nclu, beta0) for c in range(2)]
    for l in range(3):
        b.run(5, l, 1)
        for c in range(2):
            pb.iterate(sts[c], 5, c, l, 1)
            g = b.get_state(c)
            assert (g.labels == sts[c].labels).all() and (g.nclu == sts[c].nclu).all(), (l, c)
            assert (g.idxs == sts[c].idxs).all()
            np.testing.assert_allclose(g.beta, sts[c].beta, rtol=1e-9)
            np.testing.assert_allclose(g.JJ, sts[c].JJ, rtol=1e-9)
    _invariants(b.get_state(0), dims, dsets[0])
    # per-step conditional probabilities at this shape
    st = sts[0]
    for (t, it) in [(0, 0), (1, 17), (2, 499)]:
        lw, pr, K = b.score_patient(0, t, it, seed=5, iter_=9)
        lw2, pr2, K2 = pb.score_patient(st, t, it, seed=5, chain_id=0, iter_=9)
        assert K == K2
        m = pr2 > 1e-290
        np.testing.assert_allclose(pr[m], pr2[m], rtol=1e-10)
    b.close()
    ctx.close()


def test_config4_full_size_lockstep_with_oracle():
    """BASELINE config 4 at FULL size (n = 20 000 patients, 50 biomarkers, 3 arms, NGG, compact V
    table): two chains in lock-step with the oracle for 3 full iterations (sweep + every update) --
    labels, counts, K, grid index bit-identical, sufficient statistics bit-identical, reals <= 1e-9;
    per-step probabilities at three frozen positions.  (One oracle sweep takes ~0.3 s here.)"""
    from treatppmx_b200.engine import Batch, Context
    model, dims, shared, dsets, labels, nclu, beta0 = _case(n=20000, P=50, T=3, kcap=128)
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    b.init(8, labels, nclu, beta0)
    pb = ob.Problem(model, dims, shared, dsets[0])
    sts = [pb.init_state(8, c, labels, nclu, beta0) for c in range(2)]
    for l in range(3):
        b.run(8, l, 1)
        assert b.last_sweep_info()[0] == "fast"
        for c in range(2):
            pb.iterate(sts[c], 8, c, l, 1)
            g = b.get_state(c)
            assert (g.nclu == sts[c].nclu).all(), (l, c, g.nclu, sts[c].nclu)
            assert (g.labels == sts[c].labels).all(), (l, c)
            assert (g.nj == sts[c].nj).all() and (g.idxs == sts[c].idxs).all()
            for t in range(dims.T):
                K = int(g.nclu[t])
                assert (g.sumx[t, :K * dims.P] == sts[c].sumx[t, :K * dims.P]).all()
                assert (g.sumx2[t, :K * dims.P] == sts[c].sumx2[t, :K * dims.P]).all()
                np.testing.assert_allclose(g.eta_star[:, :K, t], sts[c].eta_star[:, :K, t], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(g.beta, sts[c].beta, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(g.JJ, sts[c].JJ, rtol=1e-9)
            np.testing.assert_allclose(g.sigma, sts[c].sigma, rtol=0)
    _invariants(b.get_state(1), dims, dsets[0])
    for (t, it) in [(0, 0), (1, 3333), (2, 6665)]:
        lw, pr, K = b.score_patient(1, t, it, seed=8, iter_=9)
        lw2, pr2, K2 = pb.score_patient(sts[1], t, it, seed=8, chain_id=1, iter_=9)
        assert K == K2
        m = pr2 > 1e-290
        np.testing.assert_allclose(pr[m], pr2[m], rtol=1e-10)
    b.close()
    ctx.close()


@pytest.mark.parametrize("staged", [0, 24])
def test_big_arms_more_than_32_clusters_lockstep(staged):
    """Big-arm mode (n_t in the thousands) driven ABOVE 32 clusters per arm: a tighter within-cluster
    variance (similparam v = 0.3) takes K_t from 5 to ~70-80 within two sweeps (the K_t = 50-300
    regime SURVEY section 7.3 expects at config 4).  Whole iterations through tppmx_batch_run in lock-step
    with the oracle: labels / counts / K bit-identical while K crosses 32 mid-sweep and stays above.
    staged = 0: the long-arm kernel picks its staged capacity per launch from the largest K seen so far
    (the jump from 5 to ~70 clusters inside one sweep outgrows the first choice: hand-over, then a larger
    capacity); staged = 24: a fixed small capacity, hand-over in every sweep."""
    from treatppmx_b200.engine import Batch, Context
    model, dims, shared, dsets, labels, nclu, beta0 = _case(
        n=4500, P=50, T=3, kcap=200, similparam=(0.0, 10.0, 0.3, 1.0, 2.0, 0.0, 0.1))
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    if staged:   # a fixed small staged capacity: every unit outgrows it mid-sweep and is finished by the generic kernel
        b.set_sweep_impl("fast", staged_clusters=staged)
    b.init(5, labels, nclu, beta0)
    pb = ob.Problem(model, dims, shared, dsets[0])
    sts = [pb.init_state(5, c, labels, nclu, beta0) for c in range(2)]
    kmax = 0
    for l, chunk in ((0, 1), (1, 1), (2, 2)):
        b.run(5, l, chunk)
        assert b.last_sweep_info()[0] == "fast"
        for c in range(2):
            pb.iterate(sts[c], 5, c, l, chunk)
            g = b.get_state(c)
            assert (g.nclu == sts[c].nclu).all(), (l, c, g.nclu, sts[c].nclu)
            assert (g.labels == sts[c].labels).all(), (l, c)
            assert (g.nj == sts[c].nj).all() and (g.idxs == sts[c].idxs).all()
            for t in range(dims.T):
                K = int(g.nclu[t])
                assert (g.sumx[t, :K * dims.P] == sts[c].sumx[t, :K * dims.P]).all()
            np.testing.assert_allclose(g.beta, sts[c].beta, rtol=1e-9, atol=1e-12)
            kmax = max(kmax, int(g.nclu.max()))
    assert kmax > 32, "case did not reach more than 32 clusters"
    _invariants(b.get_state(0), dims, dsets[0])
    # per-step probabilities with K > 32 candidates
    for (t, it) in [(0, 0), (1, 700), (2, 1499)]:
        lw, pr, K = b.score_patient(0, t, it, seed=5, iter_=9)
        lw2, pr2, K2 = pb.score_patient(sts[0], t, it, seed=5, chain_id=0, iter_=9)
        assert K == K2 and K > 32
        m = pr2 > 1e-290
        np.testing.assert_allclose(pr[m], pr2[m], rtol=1e-10)
    b.close()
    ctx.close()


@pytest.mark.parametrize("nclu_init", [5, 20])
def test_predictive_10k_new_patients_matches_oracle(nclu_init):
    """BASELINE config 5's predictive step: 10 000 new patients x 3 arms against the oracle
    (src/dm_ppmx_ct.cpp:1080-1385): pipred <= 1e-10 relative, drawn cluster and ypred identical.
    nclu_init = 5 runs the register kernel (<= 16 clusters per arm), 20 the staged one."""
    from treatppmx_b200.engine import Batch, Context
    model = default_model(cohesion=2)
    dims_kw, dsets, ex = datagen.genmech_synthetic(n=152, npred=10000, P=10, Q=2, T=3, n_datasets=2, seed=20240501)
    dims = datagen.make_dims(G=100, kcap=0, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=2)
    labels, nclu = datagen.initial_partition(dsets[0].treat, dims.T, dims.ntmax, nclu_init=nclu_init)
    beta0 = datagen.init_beta(dsets[0])
    ctx = Context(0)
    cds = [0, 1, 1]
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.init(21, labels, nclu, beta0)
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    sts = [pbs[cds[c]].init_state(21, c, labels, nclu, beta0) for c in range(3)]
    nit = 0 if nclu_init > 16 else 4      # the first sweep merges the 20 starting clusters: score them as they are
    if nit:
        b.run(21, 0, nit)
    for c in range(3):
        pbs[cds[c]].iterate(sts[c], 21, c, 0, nit)
    pi, yp, pc = b.predict(21, nit)
    saw_aux = False
    for c in range(3):
        g = b.get_state(c)
        assert (g.labels == sts[c].labels).all() and (g.nclu == sts[c].nclu).all()
        if nclu_init > 16:
            assert g.nclu.max() > 16
        pi2, yp2, pc2 = pbs[cds[c]].predict(sts[c], 21, c, nit)
        assert (pc[c] == pc2).all()
        assert (yp[c] == yp2).all()
        np.testing.assert_allclose(pi[c], pi2, rtol=1e-10, atol=0)
        saw_aux |= bool((pc2 > sts[c].nclu[None, :]).any())
    assert saw_aux, "no new patient picked an auxiliary cluster"
    b.close()
    ctx.close()


def test_config4_full_size_invariants():
    """n = 20 000 patients, 50 biomarkers, 3 arms, NGG, 4 chains: two full iterations keep every
    bookkeeping invariant."""
    from treatppmx_b200.engine import Batch, Context
    model, dims, shared, dsets, labels, nclu, beta0 = _case(n=20000, P=50, T=3, kcap=128)
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0] * 4)
    b.init(8, labels, nclu, beta0)
    b.run(8, 0, 2)
    for c in (0, 3):
        _invariants(b.get_state(c), dims, dsets[0])
    lab, ncl = b.get_partitions()
    assert (ncl >= 1).all() and (ncl <= 128).all()
    assert not (lab[0] == lab[3]).all()          # different chains, different partitions
    b.close()
    ctx.close()
