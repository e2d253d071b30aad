"""Point-estimate partition from the co-clustering matrix (minVI / minbinder.ext of mcclust.ext,
imported by the reference, R/treatppmx.R:30-31): the numpy restatement (oracle/point_estimate.py)
against scipy's hierarchical clustering and brute force, and the device kernel against it."""
import itertools

import numpy as np
import pytest

from oracle import point_estimate as pe


def _psm(n, k, nsamp, rng, noise=0.25):
    """a co-clustering matrix averaged over noisy copies of a k-block partition"""
    base = rng.integers(0, k, n)
    m = np.zeros((n, n))
    for _ in range(nsamp):
        l = np.where(rng.random(n) < noise, rng.integers(0, k + 2, n), base)
        m += l[:, None] == l[None, :]
    return m / nsamp


@pytest.mark.parametrize("linkage", ["average", "complete"])
def test_cuts_match_scipy_on_tie_free_matrices(linkage):
    from scipy.cluster.hierarchy import fcluster, linkage as sl
    from scipy.spatial.distance import squareform
    rng = np.random.default_rng(1)
    for n in (7, 19, 40):
        a = rng.random((n, n)) * 0.9 + 0.05
        psm = (a + a.T) / 2
        np.fill_diagonal(psm, 1.0)
        Z = sl(squareform(1.0 - psm, checks=False), method=linkage)
        for k, cl in pe.cuts(psm, linkage):
            want = fcluster(Z, t=k, criterion="maxclust")
            assert len(set(cl)) == k
            assert (pe.first_appearance(cl) == pe.first_appearance(want)).all(), (n, k)


def test_losses_against_direct_formulas_and_brute_force():
    rng = np.random.default_rng(2)
    n = 7
    psm = _psm(n, 2, 40, rng)
    # all partitions of 7 items via restricted growth strings
    def rgs(n):
        def rec(prefix, m):
            if len(prefix) == n:
                yield prefix
                return
            for v in range(m + 2):
                yield from rec(prefix + [v], max(m, v))
        yield from rec([0], 0)
    best_vi = min(pe.vi_lb(np.array(p), psm) for p in rgs(n))
    best_b = min(pe.binder(np.array(p), psm) for p in rgs(n))
    lab, k, v = pe.point_estimate(psm, "VI", "average", max_k=n)
    assert v >= best_vi - 1e-12 and v <= best_vi + 0.35        # the tree search is a heuristic: close to the optimum
    lab, k, v = pe.point_estimate(psm, "binder", "average", max_k=n)
    assert v >= best_b - 1e-12 and v <= best_b + 1.0
    # a clean block matrix is recovered exactly, with loss 0
    cl = np.array([0, 0, 1, 1, 1, 2, 2])
    clean = (cl[:, None] == cl[None, :]).astype(float)
    lab, k, v = pe.point_estimate(clean, "VI", "average", max_k=7)
    assert k == 3 and (lab == pe.first_appearance(cl)).all() and abs(v) < 1e-15
    assert pe.binder(cl, clean) == 0.0 and pe.vi_lb(np.zeros(7, int), clean) > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("loss,linkage", [("VI", "average"), ("binder", "average"), ("VI", "complete"), ("binder", "complete")])
def test_device_point_estimate_matches_restatement(loss, linkage):
    from cases import build_case
    from treatppmx_b200.engine import Batch, Context
    nd, cpd = 3, 6
    model, dims, shared, dsets, labels, nclu, betas = build_case("default_ngg", n_datasets=nd)
    ctx = Context(0)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=np.repeat(np.arange(nd), cpd))
    b.init(3, labels, nclu, betas[0])
    b.reset_summaries()
    b.run(3, 0, 60, burn=20, thin=2, psm=True, predictive=False)
    out = b.summaries(want=("psm",))
    for max_k in (0, 20):
        lab, k, ls = b.point_estimate(loss, linkage, max_k)
        for ds in range(nd):
            n_a = np.bincount(dsets[ds].treat, minlength=dims.T)
            for t in range(dims.T):
                n = int(n_a[t])
                psm = out["psm"][ds, t, :n, :n]
                wl, wk, wv = pe.point_estimate(psm, loss, linkage, max_k)
                kk = int(k[ds, t])
                assert 1 <= kk <= (max_k if max_k else (n + 7) // 8)
                # the device's cut is a cut of the same tree, its loss is what the restatement computes for
                # that cut, and it is the minimum (two cuts may tie up to rounding: then either k is right)
                cut = dict(pe.cuts(psm, linkage))[kk]
                assert (lab[ds, t, :n] == pe.first_appearance(cut)).all() and (lab[ds, t, n:] == 0).all()
                f = pe.vi_lb if loss == "VI" else pe.binder
                assert ls[ds, t] == pytest.approx(f(cut, psm), rel=1e-11, abs=1e-12)
                assert ls[ds, t] == pytest.approx(wv, rel=1e-11, abs=1e-12)
                if kk == wk:
                    assert (lab[ds, t, :n] == wl).all()
    b.close()
    ctx.close()
