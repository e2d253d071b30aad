"""The in-library multi-device entry (include/tppmx.h: tppmx_multi_create / tppmx_multi_run_study):
replicate datasets sharded over the devices inside ONE process, summaries gathered with NCCL inside
the library.  On one device it must equal the same study run through a single batch; on two devices
(skipped when the box has one) it must equal the one-device result (co-clustering bit for bit) -- chain c of dataset
d keeps Philox stream d * chains_per_dataset + c whatever the number of devices."""
import numpy as np
import pytest

from cases import build_case

pytestmark = pytest.mark.gpu

UW = [0.0, 40.0, 100.0]


def _study():
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg", n_datasets=5)
    return model, dims, shared, dsets, labels, nclu, betas[0]


def _same(a, b):
    """co-clustering counts and chosen arms identical; the predictive means are sums of doubles over the
    chains of a dataset accumulated with atomics, so their last bit depends on the arrival order"""
    np.testing.assert_array_equal(a["psm"], b["psm"])
    np.testing.assert_array_equal(a["best_arm"], b["best_arm"])
    np.testing.assert_allclose(a["pi_mean"], b["pi_mean"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(a["utility"], b["utility"], rtol=1e-13, atol=0)


def _single_batch(model, dims, shared, dsets, labels, nclu, beta0, cpd, seed, n_iter, burn, thin):
    from treatppmx_b200.engine import Batch, Context
    ctx = Context(0)
    cds = np.repeat(np.arange(len(dsets), dtype=np.int32), cpd)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)      # chain ids 0..n-1 = d * cpd + c
    b.init(seed, labels, nclu, beta0)
    b.reset_summaries()
    b.run(seed, 0, n_iter, burn=burn, thin=thin, psm=True, predictive=True)
    out = b.summaries(UW)
    b.close()
    ctx.close()
    return out


def test_multi_one_device_equals_single_batch():
    from treatppmx_b200.multigpu import run_study_multi
    st = _study()
    want = _single_batch(*st, cpd=3, seed=9, n_iter=14, burn=4, thin=2)
    got = run_study_multi(*st[:4], 3, st[4], st[5], st[6], seed=9, n_iter=14, burn=4, thin=2, util_w=UW)
    assert got["info"]["n_devices"] == 1 and not got["info"]["used_nccl"] and got["info"]["ms_run"] > 0
    _same(got, want)
    assert got["psm"].max() == 1.0 and (np.diagonal(got["psm"][0, 0, :5, :5]) == 1.0).all()


def test_multi_two_devices_equal_one_device():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from treatppmx_b200.multigpu import run_study_multi
    st = _study()
    a = run_study_multi(*st[:4], 2, st[4], st[5], st[6], seed=4, n_iter=12, burn=2, thin=2, util_w=UW)
    b = run_study_multi(*st[:4], 2, st[4], st[5], st[6], seed=4, n_iter=12, burn=2, thin=2, util_w=UW, device_ids=[0, 1])
    assert b["info"]["n_devices"] == 2 and b["info"]["used_nccl"] and b["info"]["ms_gather"] > 0
    _same(a, b)                                                   # 5 datasets over 2 devices: blocks of 3 and 2
    c = run_study_multi(*st[:4], 2, st[4], st[5], st[6], seed=4, n_iter=12, burn=2, thin=2, util_w=UW, device_ids=[1, 0])
    _same(a, c)


def test_multi_rejects_bad_devices():
    from treatppmx_b200 import _lib
    from treatppmx_b200.multigpu import run_study_multi
    st = _study()
    with pytest.raises(_lib.TppmxError):
        run_study_multi(*st[:4], 1, st[4], st[5], st[6], seed=1, n_iter=1, device_ids=[0, 0])
    with pytest.raises(_lib.TppmxError):
        run_study_multi(*st[:4], 1, st[4], st[5], st[6], seed=1, n_iter=1, device_ids=[99])
