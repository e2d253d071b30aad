"""Edge cases of the path through the C ABI: ragged / tiny / empty arms, no prognostic covariates,
no predictive covariates in the score (PPMx = 0), the largest supported outcome / auxiliary sizes,
no new patients, a single patient per arm."""
import numpy as np
import pytest

import oracle_binding as ob
from cases import build_case
from treatppmx_b200 import _lib, datagen
from treatppmx_b200._abi import HostDataset, default_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from treatppmx_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _custom(n, T, P, Q, D, npred, treat, seed=5, **mkw):
    """A dataset with an arbitrary arm assignment and outcome dimension."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n + max(npred, 1), max(P, 1)))[:, :P] if P > 0 else np.zeros((n + max(npred, 1), 0))
    Z = rng.standard_normal((n + max(npred, 1), max(Q, 1)))[:, :Q] if Q > 0 else np.zeros((n + max(npred, 1), 0))
    y = np.zeros((n, D))
    y[np.arange(n), rng.integers(0, D, n)] = 1.0
    ds = HostDataset(treat=np.asarray(treat, np.int32), y=y, z=Z[:n], xcon=X[:n], xcat=np.zeros((n, 1), np.int32),
                     zpred=Z[n:n + max(npred, 1)], xconp=X[n:n + max(npred, 1)], xcatp=np.zeros((max(npred, 1), 1), np.int32))
    model = default_model(**mkw)
    n_a = np.bincount(ds.treat, minlength=T)
    dims = datagen.make_dims(G=100, kcap=0, nobs=n, T=T, D=D, Q=Q, P=P, ncat=0, maxC=0, npred=npred, ntmax=int(max(n_a.max(), 1)))
    shared = datagen.make_shared(dims, np.maximum(n_a, 1), cohesion=model.cohesion)
    labels, nclu = datagen.initial_partition(ds.treat, T, dims.ntmax, nclu_init=3)
    nclu = np.maximum(nclu, 1).astype(np.int32)
    beta0 = np.full(Q * D, 0.01)
    return model, dims, shared, ds, labels, nclu, beta0


def _lockstep(ctx, case, n_iter=8, nchains=2, full=True):
    from treatppmx_b200.engine import Batch
    model, dims, shared, ds, labels, nclu, beta0 = case
    b = Batch(ctx, model, dims, shared, [ds], chain_dataset=[0] * nchains)
    b.init(3, labels, nclu, beta0)
    pb = ob.Problem(model, dims, shared, ds)
    for c in range(nchains):
        st = pb.init_state(3, c, labels, nclu, beta0)
        if c == 0:
            if full:
                b.run(3, 0, n_iter)
            else:
                b.sweep(3, 0, n_iter)
        if full:
            pb.iterate(st, 3, c, 0, n_iter)
        else:
            pb.sweep(st, 3, c, 0, n_iter)
        g = b.get_state(c)
        assert (g.labels == st.labels).all() and (g.nclu == st.nclu).all()
        for t in range(dims.T):
            K = int(st.nclu[t])
            assert (g.nj[t, :K] == st.nj[t, :K]).all()
        np.testing.assert_allclose(g.JJ, st.JJ, rtol=1e-9)
        np.testing.assert_allclose(g.beta, st.beta, rtol=1e-9, atol=1e-14)
    return b, pb


def test_ragged_arms(ctx):
    treat = [0] * 37 + [1] * 5 + [2] * 1 + [0] * 3 + [1] * 2          # arm sizes 40 / 7 / 1
    _lockstep(ctx, _custom(48, 3, 4, 2, 3, 6, treat))


@pytest.mark.parametrize("mkw", [dict(calibration=1), dict(calibration=1, cohesion=1), dict(consim=2), dict(consim=2, cohesion=1),
                                 dict(similarity=2, cohesion=1), dict(PPMx=0, cohesion=1)],
                         ids=["calib1-ngg", "calib1-dp", "nnig-ngg", "nnig-dp", "double-dipper-dp", "ppm-only"])
def test_ragged_and_single_patient_arms_on_the_special_instantiations(ctx, mkw):
    """The fast kernel's instantiations for calibration 1, the NNIG similarity, the double dipper and PPMx = 0 on arms of
    40 / 7 / 1 patients (a one-patient arm is scored against auxiliary clusters only: the occupied-cluster normaliser
    of calibration 1 is then empty) and on two one-patient arms."""
    treat = [0] * 37 + [1] * 5 + [2] * 1 + [0] * 3 + [1] * 2
    b, pb = _lockstep(ctx, _custom(48, 3, 4, 2, 3, 6, treat, **mkw))
    assert b.last_sweep_info()[0] == "fast"
    b, pb = _lockstep(ctx, _custom(2, 2, 3, 1, 3, 2, [0, 1], **dict(mkw, cohesion=1)))
    assert b.last_sweep_info()[0] == "fast"


def test_single_patient_arms_and_tiny_data(ctx):
    _lockstep(ctx, _custom(2, 2, 3, 1, 3, 2, [0, 1], cohesion=1))
    _lockstep(ctx, _custom(3, 1, 2, 2, 3, 1, [0, 0, 0], cohesion=1))


def test_no_prognostic_covariates(ctx):
    """Q = 0 with noprog = 1: beta is frozen (src/dm_ppmx_ct.cpp:1009)."""
    _lockstep(ctx, _custom(40, 2, 5, 0, 3, 4, np.arange(40) % 2, noprog=1))


def test_outcome_dimension_8_and_8_auxiliary_clusters(ctx):
    _lockstep(ctx, _custom(45, 2, 3, 2, 8, 5, np.arange(45) % 2, CC=8, cohesion=1))
    _lockstep(ctx, _custom(45, 2, 3, 3, 5, 5, np.arange(45) % 2, CC=8))


def test_sixteen_prognostic_covariates(ctx):
    _lockstep(ctx, _custom(50, 2, 3, 16, 3, 4, np.arange(50) % 2, cohesion=1), n_iter=5)


def test_ppm_without_covariates_and_no_new_patients(ctx):
    b, pb = _lockstep(ctx, _custom(40, 2, 0, 2, 3, 0, np.arange(40) % 2, PPMx=0, cohesion=1))
    with pytest.raises(_lib.TppmxError):
        b.predict(1, 0)                      # npred == 0


def test_many_covariates_use_the_wide_row_path(ctx):
    """P = 40 > 16 lanes: the fast kernel's multi-register covariate row (XR = 4)."""
    case = _custom(60, 2, 40, 2, 3, 4, np.arange(60) % 2)
    b, pb = _lockstep(ctx, case, n_iter=6, full=False)
    assert b.last_sweep_info()[0] == "fast"


def test_empty_arm_is_rejected_cleanly(ctx):
    """An arm without patients: ncluster_curr would be 0, which the reference cannot represent either."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, ds, labels, nclu, beta0 = _custom(20, 3, 3, 2, 3, 2, np.arange(20) % 2, cohesion=1)
    nclu[2] = 0
    b = Batch(ctx, model, dims, shared, [ds], chain_dataset=[0])
    with pytest.raises(_lib.TppmxError):
        b.init(3, labels, nclu, beta0)


def test_set_state_rejects_out_of_range_indices():
    """labels index eta / sumx / nj with label - 1 and idxs indexes the V table on the device: a bad
    state is refused on the host instead of reading out of bounds."""
    from treatppmx_b200 import _lib
    from treatppmx_b200.engine import Batch, Context
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg")
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0])
    b.init(1, labels, nclu, betas[0])
    good = b.get_state(0)
    for field, idx, val in (("labels", (0, 3), 99), ("labels", (1, 0), 0), ("idxs", (0,), dims.G), ("idxs", (0,), -1),
                            ("nj", (0, 1), -2)):
        st = good.copy()
        getattr(st, field)[idx] = val
        with pytest.raises(_lib.TppmxError) as ei:
            b.set_state(0, st)
        assert ei.value.code == 1, field
    b.set_state(0, good)          # and the untouched state still goes through
    b.sweep(1, 0, 2)
    b.close()
    ctx.close()
