"""BASELINE config 4 shape class (large n, P = 50 predictive biomarkers, NGG cohesion with the
compact log V table, cluster capacity below max n_t): parity against the oracle at a size the
oracle finishes in seconds, and size-independent invariants at the full n = 20 000."""
import numpy as np
import pytest

import oracle_binding as ob
from treatppmx_b200 import datagen
from treatppmx_b200._abi import default_model

pytestmark = pytest.mark.gpu


def _case(n, P, T, kcap, n_datasets=1, seed=20240401, npred=0):
    model = default_model(cohesion=2)
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


def test_config4_full_size_invariants():
    """n = 20 000 patients, 50 biomarkers, 3 arms, NGG, 4 chains: two full iterations keep every
    bookkeeping invariant (the oracle would need minutes here; the invariants do not)."""
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
