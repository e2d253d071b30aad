"""Device-side posterior summaries (co-clustering counts, predictive means, utility, chosen arm)
against numpy recomputations from the per-chain outputs of the same calls."""
import numpy as np
import pytest

from cases import build_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from treatppmx_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["default_ngg", "dp_T3"])
def test_psm_and_predictive_means(ctx, name):
    from treatppmx_b200.engine import Batch
    nd, cpd = 3, 4
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=nd)
    cds = np.repeat(np.arange(nd), cpd)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.init(4, labels, nclu, betas[0])
    b.sweep(4, 0, 10)
    b.reset_summaries()
    n_a = [np.bincount(d.treat, minlength=dims.T) for d in dsets]
    psm_ref = np.zeros((nd, dims.T, dims.ntmax, dims.ntmax))
    pi_ref = np.zeros((nd, dims.npred, dims.D, dims.T))
    saves = 6
    for s in range(saves):
        b.sweep(4, 10 + s, 1)
        b.accumulate(4, 10 + s, psm=True, predictive=True)
        lab, _ = b.get_partitions()
        pi, _, _ = b.predict(4, 10 + s, want=("pipred",))
        for c in range(len(cds)):
            for t in range(dims.T):
                l = lab[c, t, : n_a[cds[c]][t]]
                psm_ref[cds[c], t, : len(l), : len(l)] += (l[:, None] == l[None, :])
            pi_ref[cds[c]] += pi[c]
    psm_ref /= saves * cpd
    pi_ref /= saves * cpd
    w = np.array([0.0, 40.0, 100.0])
    out = b.summaries(w)
    np.testing.assert_allclose(out["psm"], psm_ref, rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["pi_mean"], pi_ref, rtol=1e-13, atol=1e-15)
    util = np.einsum("k,npkt->ntp", w, pi_ref)
    np.testing.assert_allclose(out["utility"], util, rtol=1e-12)
    assert (out["best_arm"] == util.argmax(axis=1)).all()
    # npc (R/countUT.R:26-53): which.max over categories under the assigned arm == observed outcome
    rng = np.random.default_rng(8)
    trt = rng.integers(0, dims.T, (nd, dims.npred))
    obs = rng.integers(0, dims.D, (nd, dims.npred))
    want_npc = [sum(int(np.argmax(out["pi_mean"][ds, pp, :, trt[ds, pp]]) == obs[ds, pp]) for pp in range(dims.npred))
                for ds in range(nd)]
    assert b.npc(trt, obs).tolist() == want_npc
    # diagonal of the co-clustering matrix is 1 for real patients
    for ds in range(nd):
        for t in range(dims.T):
            assert np.allclose(np.diag(out["psm"][ds, t])[: n_a[ds][t]], 1.0)
    # device views agree with the host copies
    import torch
    from treatppmx_b200.multigpu import device_summaries_as_torch
    dv = device_summaries_as_torch(b, w)
    np.testing.assert_array_equal(dv["psm"].cpu().numpy(), out["psm"])
    np.testing.assert_array_equal(dv["best_arm"].cpu().numpy(), out["best_arm"])
    np.testing.assert_array_equal(dv["pi_mean"].cpu().numpy().transpose(0, 3, 2, 1), out["pi_mean"])
