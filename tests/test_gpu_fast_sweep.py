"""The fast sweep kernel (treatppmx_b200/csrc/sweep_fast.cuh) against the CPU oracle and against
the generic kernel: same Philox streams => labels / counts / K bit-identical, sufficient
statistics bit-identical (same operations in the same order), hand-over to the generic kernel
exact at any (sweep, patient) position."""
import numpy as np
import pytest

import oracle_binding as ob
from cases import build_case, occupied_view
from treatppmx_b200 import _lib, datagen
from treatppmx_b200._abi import default_model

pytestmark = pytest.mark.gpu

FAST_CASES = ["default_ngg", "dp_T3", "calib2", "calib2_sqrt", "Q3_D3", "prior_only", "ppm_only", "noreuse", "noreuse_T3"]


@pytest.fixture(scope="module")
def ctx():
    from treatppmx_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _assert_same(g, o, dims, what=""):
    assert (g.nclu == o.nclu).all(), what
    assert (g.labels == o.labels).all(), what
    gv, ov = occupied_view(g, dims), occupied_view(o, dims)
    for k in gv:
        if k.startswith(("nj", "nclu", "sumx")):
            assert (gv[k] == ov[k]).all(), (what, k)
        else:
            np.testing.assert_allclose(gv[k], ov[k], rtol=1e-12, atol=1e-13, err_msg=f"{what} {k}")


# calibration 1 (similarities normalised over the clusters, src/dm_ppmx_ct.cpp:717-788) has its own instantiation
# of the fast kernel, for 16 | 32 lanes per unit
CAL1_CASES = [("calib1", 16), ("calib1", 32), ("calib1_dp_T3", 16), ("calib1_dp_T3", 32)]
# so has the Normal-Normal-Inverse-Gamma similarity (consim = 2, src/utils_ct.cpp:427-459)
CAL1_CASES += [("nnig", 16), ("nnig", 32), ("nnig_ngg_T3", 16), ("nnig_ngg_T3", 32)]
# and categorical covariates beside the continuous ones (Dirichlet-multinomial similarity, src/utils_ct.cpp:468-495)
CAL1_CASES += [(n, l) for n in ("categorical", "categorical_T3", "cat_calib2") for l in (16, 32)]
# and the double-dipper Normal-Normal similarity (similarity = 2)
CAL1_CASES += [(n, l) for n in ("double_dipper", "dd_dp_T3") for l in (16, 32)]


@pytest.mark.parametrize("name,lpu", [(n, l) for n in FAST_CASES for l in (8, 16, 32)] + CAL1_CASES)
def test_fast_lockstep_with_oracle(ctx, name, lpu):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=2)
    nchains = 5
    cds = [c % 2 for c in range(nchains)]
    batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    batch.set_sweep_impl("fast", lpu)
    batch.init(31, labels, nclu, betas[0])
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    sts = [pbs[cds[c]].init_state(31, c, labels, nclu, betas[0]) for c in range(nchains)]
    it0 = 0
    for chunk in (1, 2, 37):
        batch.sweep(31, it0, chunk)
        assert batch.last_sweep_info()[0] == "fast"
        for c in range(nchains):
            pbs[cds[c]].sweep(sts[c], 31, c, it0, chunk)
            g = batch.get_state(c)
            _assert_same(g, sts[c], dims, f"{name} lpu={lpu} chain={c} after {it0 + chunk}")
            np.testing.assert_allclose(g.loggamma, sts[c].loggamma, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(g.eta_empty, sts[c].eta_empty, rtol=1e-12, atol=1e-13)
        it0 += chunk


def test_fast_equals_generic_many_chains(ctx):
    """64 chains x 100 sweeps: the two kernels must leave identical partitions."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg", n_datasets=4)
    cds = [c % 4 for c in range(64)]
    out = {}
    for impl in ("generic", "fast"):
        batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
        batch.set_sweep_impl(impl)
        batch.init(5, labels, nclu, betas[0])
        batch.sweep(5, 0, 100)
        assert batch.last_sweep_info()[0] == impl
        out[impl] = batch.get_partitions()
        st = batch.get_state(63)
        out[impl + "_sumx"] = st.sumx.copy()
        batch.close()
    assert (out["generic"][0] == out["fast"][0]).all()
    assert (out["generic"][1] == out["fast"][1]).all()
    assert (out["generic_sumx"] == out["fast_sumx"]).all()


@pytest.mark.parametrize("staged,lpu", [(3, 16), (4, 8), (6, 16), (5, 32)])
def test_handover_to_generic_is_exact(ctx, staged, lpu):
    """A tiny staged capacity makes units outgrow the fast kernel all the time, at arbitrary
    (sweep, patient) positions; the generic kernel must finish them exactly."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("dp_T3", n_datasets=2)
    nchains = 6
    cds = [c % 2 for c in range(nchains)]
    batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    batch.set_sweep_impl("fast", lpu, staged_clusters=staged)
    batch.init(8, labels, nclu, betas[0])
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    sts = [pbs[cds[c]].init_state(8, c, labels, nclu, betas[0]) for c in range(nchains)]
    handovers, fast_only, it0 = 0, 0, 0
    for chunk in (3, 10, 1, 1, 1, 1, 1, 1, 17):
        batch.sweep(8, it0, chunk)
        impl, nh = batch.last_sweep_info()
        assert impl == "fast"
        handovers += nh
        fast_only += (nh < nchains * dims.T)
        for c in range(nchains):
            pbs[cds[c]].sweep(sts[c], 8, c, it0, chunk)
            _assert_same(batch.get_state(c), sts[c], dims, f"staged={staged} chain={c} after {it0 + chunk}")
        it0 += chunk
    assert handovers > 0, "case did not exercise the hand-over"
    assert fast_only > 0, "case never let the fast kernel finish a unit"


def test_small_kcap_capacity_error_and_handover(ctx):
    """kcap below what the chains need: the fast kernel hands over, the generic kernel raises
    TPPMX_ERR_CAPACITY exactly as it does on its own."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("dp_T3", kcap=5)
    batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    batch.set_sweep_impl("fast")
    batch.init(1, labels, nclu, betas[0])
    with pytest.raises(_lib.TppmxError) as ei:
        batch.sweep(1, 0, 200)
    assert ei.value.code == 3


def test_calibration1_handover_beyond_the_lanes(ctx):
    """Calibration 1 normalises over all K + CC candidates in one pass of the unit's lanes, so its staged capacity
    is lanes - CC; a partition driven beyond that hands over to the generic kernel, exactly."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("calib1_dp_T3")
    labels, nclu = datagen.initial_partition(dsets[0].treat, dims.T, dims.ntmax, nclu_init=15)   # starts beyond 16 - CC
    batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0, 0])
    batch.set_sweep_impl("fast", 16)
    batch.init(5, labels, nclu, betas[0])
    pb = ob.Problem(model, dims, shared, dsets[0])
    sts = [pb.init_state(5, c, labels, nclu, betas[0]) for c in range(3)]
    hand = 0
    for l in range(6):
        batch.sweep(5, l, 1)
        hand += batch.last_sweep_info()[1]
        for c in range(3):
            pb.sweep(sts[c], 5, c, l, 1)
            _assert_same(batch.get_state(c), sts[c], dims, f"calib1 hand-over chain {c} sweep {l}")
    assert hand > 0


def test_fast_rejected_for_ineligible_model(ctx):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = build_case("nnig_dd")   # NNIG double dipper: generic kernel only
    batch = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0])
    batch.set_sweep_impl("fast")
    batch.init(1, labels, nclu, betas[0])
    with pytest.raises(_lib.TppmxError):
        batch.sweep(1, 0, 1)
    batch.set_sweep_impl("auto")
    batch.sweep(1, 0, 1)
    assert batch.last_sweep_info()[0] == "generic"
