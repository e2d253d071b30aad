"""BASELINE config 1: ppmxct() on the bundled simupats data.  The reader of the reference's
data/simupats.rda and the genmech recipe (treatppmx_b200/simupats.py) against the committed fixture,
and posterior-level agreement of the CUDA drop-in with the oracle's chains on that data set.  (The
exact agreement of both with the REFERENCE's own output on these inputs is in
tests/test_reference_golden.py: cases simupats_ngg / simupats_dp.)"""
import os

import numpy as np
import pytest

import refrun
from cases import SIMUPATS_FIXTURE, build_case

RDA = "/root/reference/data/simupats.rda"


@pytest.mark.skipif(not os.path.exists(RDA), reason="reference checkout not present (GPU box)")
def test_rda_reader_and_recipe_reproduce_the_fixture():
    from treatppmx_b200 import simupats
    mat = simupats.read_rda_matrix(RDA)
    assert mat.shape == (152, 92)
    assert mat.min() == pytest.approx(0.1) and mat.max() == pytest.approx(6.461623898225299)
    assert mat.mean() == pytest.approx(1.129696495523364, rel=1e-14)
    f = np.load(SIMUPATS_FIXTURE)
    np.testing.assert_allclose([mat.sum(), (mat ** 2).sum()], f["raw_checksum"], rtol=1e-14)
    sim = simupats.genmech(mat, npred=10, progscen=1, nset=1, seed=121)
    np.testing.assert_allclose(sim["cov"], f["cov"], rtol=1e-12, atol=1e-12)
    assert (sim["treatment"] == f["treatment"]).all()
    np.testing.assert_allclose(sim["prob"][0], f["prob1"], atol=1e-15)
    # genmech's recipe: standardised columns, alternating arms, probabilities that sum to one
    assert np.allclose(sim["cov"].mean(0), 0, atol=1e-12) and np.allclose(sim["cov"].std(0, ddof=1), 1)
    assert np.allclose(sim["prob"][1].sum(1), 1) and set(sim["treatment"]) == {1, 2}


def test_fixture_shapes_are_config1():
    model, dims, shared, dsets, labels, nclu, betas = build_case("simupats_ngg")
    assert (dims.nobs, dims.T, dims.D, dims.Q, dims.P, dims.npred, dims.ntmax) == (124, 2, 3, 2, 10, 28, 62)
    assert (model.PPMx, model.cohesion, model.similarity, model.consim, model.calibration, model.CC, model.reuse) == \
        (1, 2, 1, 1, 0, 3, 1)      # defaults of ppmxct(), R/dm_ppmx_ct.R:89-93
    assert (nclu == 5).all() and labels.max() == 5
    assert np.isfinite(shared.logV[:, 1, :62, :]).all() and np.isfinite(shared.logV[:, 0, :61, :]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["simupats_ngg", "simupats_dp"])
def test_config1_posterior_agrees_with_oracle_chains(name):
    """North-star level 3 on the real data set: chains of the CUDA drop-in (seeds 1..8) against chains
    of the CPU oracle (seeds 11..18), ppmxct()'s default length (iter 1100, burn 100): number of
    clusters, predictive response probabilities of the 28 new patients and their chosen arm agree
    within Monte-Carlo error (two-sample z-scores over the independent chains)."""
    from treatppmx_b200.engine import Context
    from treatppmx_b200.ppmxct import dm_ppmx_ct
    iters, burn, thin = 1100, 100, 5
    case = build_case(name)
    model, dims, shared, dsets, labels, nclu, betas = case
    args, logV = refrun.reference_args_of(model, dims, shared, dsets[0], labels, nclu, betas[0], iters, burn, thin, False)
    ctx = Context(0)
    dev = [dm_ppmx_ct(ctx, **args, seed=s, logV_compact=logV) for s in range(1, 9)]
    ctx.close()
    ora = [refrun.oracle_run(case, iters, burn, thin, s) for s in range(11, 19)]

    def stats(runs, f):
        v = np.array([f(r) for r in runs])
        return v.mean(0), v.std(0, ddof=1) / np.sqrt(len(runs))

    util_w = np.array([0.0, 40.0, 100.0])
    for f, what in ((lambda r: r["nclus"].mean(1), "mean number of clusters per arm"),
                    (lambda r: r["pipred"].mean(3), "posterior mean of pipred"),
                    (lambda r: r["pi"].mean(2), "posterior mean of the in-sample pi")):
        md, sd = stats(dev, f)
        mo, so = stats(ora, f)
        z = (md - mo) / np.sqrt(sd ** 2 + so ** 2 + 1e-12)
        assert np.abs(z).max() < 6.5 and np.sqrt((z ** 2).mean()) < 1.8, (what, np.abs(z).max())
    # treatment choice: identical wherever the utility gap between the arms is clear of the Monte-Carlo
    # noise of either set of chains (gap of each chain -> mean and standard error over the chains)
    gap = lambda r: np.einsum("pdt,d->pt", r["pipred"].mean(3), util_w) @ np.array([1.0, -1.0])
    gd, sd = stats(dev, gap)
    go, so = stats(ora, gap)
    clear = np.abs(go) > 3.5 * (sd + so)     # may be few: on this data set most gaps are within chain-to-chain noise
    assert (np.sign(gd[clear]) == np.sign(go[clear])).all()
    z = (gd - go) / np.sqrt(sd ** 2 + so ** 2 + 1e-12)
    assert np.abs(z).max() < 6.5
