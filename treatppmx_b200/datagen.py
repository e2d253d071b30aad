"""Synthetic inputs for the PPMx hot path: genmech-style datasets, the (sigma,kappa) grid,
NGG V-weights and the initial partition.  Host-side numpy only (inputs, not compute path).

Follows the recipe of the reference's generator (R/gendata.R:134-254 `genmech`,
:30-47 `genoutcome`) with synthetic Gaussian "genes" in place of the bundled simupats
matrix, and the marshalling of R/dm_ppmx_ct.R:190-254 (beta init, grid, V table).
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln

from ._abi import Dims, HostDataset, HostShared

# reference R/dm_ppmx_ct.R:235-240
SIGMAGRID = np.array([0.01, 0.10, 0.12, 0.14, 0.16, 0.18, 0.20, 0.23, 0.27, 0.58])
KAPPAGRID = np.array([0.01, 0.57, 0.89, 1.19, 1.51, 1.87, 2.29, 2.85, 3.73, 13.91])


def reference_grid() -> np.ndarray:
    """2 x 100 grid, row 0 = sigma, row 1 = kappa; kappa varies fastest
    (t(expand.grid(kappagrid, sigmagrid))[c(2,1),], R/dm_ppmx_ct.R:239)."""
    sig = np.repeat(SIGMAGRID, len(KAPPAGRID))
    kap = np.tile(KAPPAGRID, len(SIGMAGRID))
    return np.asfortranarray(np.stack([sig, kap], axis=0))


def ngg_logV(n: int, kmax: int, sigma: float, kappa: float, npts: int = 4097) -> np.ndarray:
    """log V_{n,k}, k=1..kmax, of the normalised generalised gamma process with Laplace
    exponent psi(u) = ((u+kappa)^sigma - kappa^sigma)/sigma:
        V_{n,k} = 1/Gamma(n) * int_0^inf u^(n-1) (u+kappa)^(k sigma - n) exp(-psi(u)) du.
    The reference takes this table from vweights::computev (un-vendored, unpinned:
    R/dm_ppmx_ct.R:247), so the table is INPUT DATA here; this generator uses the positive
    integral form (stable for n in the thousands) and is checked against the Gibbs-type
    recursion V_{n,k} = (n - k sigma) V_{n+1,k} + V_{n+1,k+1} in tests/."""
    k = np.arange(1, kmax + 1, dtype=np.float64)[:, None]

    def f(t):
        e = np.exp(t)
        return n * t + (k * sigma - n) * np.log(e + kappa) - ((e + kappa) ** sigma - kappa ** sigma) / sigma

    def fp(t):
        e = np.exp(t)
        return n + (k * sigma - n) * e / (e + kappa) - e * (e + kappa) ** (sigma - 1.0)

    lo = np.full((kmax, 1), -60.0)
    hi = np.full((kmax, 1), 60.0)
    for _ in range(70):  # f is strictly concave in t: bisection on f'
        mid = 0.5 * (lo + hi)
        pos = fp(mid) > 0
        lo = np.where(pos, mid, lo)
        hi = np.where(pos, hi, mid)
    ts = 0.5 * (lo + hi)
    e = np.exp(ts)
    fpp = ((k * sigma - n) * kappa * e / (e + kappa) ** 2
           - e * (e + kappa) ** (sigma - 2.0) * (sigma * e + kappa))
    w = 36.0 / np.sqrt(np.maximum(-fpp, 1e-12))
    w = np.minimum(w, 60.0)
    lin = np.linspace(-1.0, 1.0, npts)[None, :]
    t = ts + w * lin
    ft = f(t)
    fmax = ft.max(axis=1, keepdims=True)
    h = (w * 2.0 / (npts - 1))
    integ = np.exp(ft - fmax)
    integ[:, 0] *= 0.5
    integ[:, -1] *= 0.5
    logI = fmax[:, 0] + np.log(integ.sum(axis=1) * h[:, 0])
    return logI - gammaln(n)


def ngg_logV_table(n_a, ntmax: int, grid: np.ndarray, kmax: int | None = None) -> np.ndarray:
    """Compact table [T][2][ntmax][G] of include/tppmx.h: r=0 -> n_t-1, r=1 -> n_t.
    Entries with K > n (undefined) or K > kmax are -inf, like the zeroed infinities of the
    dense cube (R/dm_ppmx_ct.R:254)."""
    T, G = len(n_a), grid.shape[1]
    out = np.full((T, 2, ntmax, G), -np.inf)
    cache = {}
    for t, nt in enumerate(n_a):
        for r in (0, 1):
            n = int(nt) - 1 + r
            if n < 1:
                continue
            km = min(n, ntmax if kmax is None else min(kmax, ntmax))
            for g in range(G):
                key = (n, km, float(grid[0, g]), float(grid[1, g]))
                if key not in cache:
                    cache[key] = ngg_logV(n, km, grid[0, g], grid[1, g])
                out[t, r, :km, g] = cache[key]
    return out


def _spearman_fdr_beta(y: np.ndarray, Z: np.ndarray) -> np.ndarray:
    """beta initialisation of R/dm_ppmx_ct.R:190-206: Spearman correlation of each
    prognostic covariate with each compositional outcome column, kept where the
    FDR-adjusted p-value <= 0.2, plus 1e-4.  Returned in the C++ index h = k + q*D."""
    from scipy.stats import spearmanr

    D, Q = y.shape[1], Z.shape[1]
    yy = y / y.sum(axis=1, keepdims=True)
    cor = np.zeros((D, Q))
    pv = np.ones((D, Q))
    for rr in range(D):
        for cc in range(Q):
            if np.std(yy[:, rr]) == 0 or np.std(Z[:, cc]) == 0:
                continue
            res = spearmanr(Z[:, cc], yy[:, rr])
            cor[rr, cc], pv[rr, cc] = res.statistic, res.pvalue
    p = pv.flatten(order="F")
    m = len(p)
    order = np.argsort(p)
    adj = np.empty(m)
    adj[order] = np.minimum.accumulate((p[order] * m / np.arange(1, m + 1))[::-1])[::-1]
    adj = np.minimum(adj, 1.0)
    keep = (adj <= 0.2).reshape((D, Q), order="F")
    betmat = np.nan_to_num(cor * keep)
    # R: beta <- as.vector(t(betmat)) + 1e-4 -> element q + k*Q; the C++ side reads it with
    # h = k + q*D.  The flat vector is passed through unchanged.
    return betmat.T.flatten(order="F") + 1e-4


def genmech_synthetic(n: int = 152, npred: int = 28, P: int = 10, Q: int = 2, T: int = 2,
                      n_datasets: int = 1, seed: int = 20240001, ncat: int = 0, maxC: int = 3,
                      mixture: int = 0, first_replicate: int | None = None):
    """genmech-style replicate datasets (R/gendata.R:134-254): one covariate matrix, fresh
    outcomes per replicate.  Returns (dims_kwargs, [HostDataset...], extras).

    X: AR(0.5)-correlated standard normals, column-standardised (scale()); optional
    `mixture`-component Gaussian mixture in the first 5 columns (config 4).  Z: N(0,1).
    Outcome: D=3 ordinal logit on metx built from the first two principal components
    (genoutcome, :30-47), product with the prognostic-only probabilities (:187-193).
    Arms cycle 0..T-1 (trtsgn, :196).

    first_replicate: if given, replicate k (k = first_replicate + 0, 1, ...) draws its outcomes
    from its own stream keyed by (seed, k), so the ranks of a multi-GPU run can each build their
    slice of ONE simulation study (same covariates, disjoint replicates)."""
    rng = np.random.default_rng(seed)
    ntot = n + npred
    e = rng.standard_normal((ntot, P))
    X = np.empty_like(e)
    X[:, 0] = e[:, 0]
    for p in range(1, P):
        X[:, p] = 0.5 * X[:, p - 1] + np.sqrt(1 - 0.25) * e[:, p]
    if mixture > 0:
        comp = rng.integers(0, mixture, size=ntot)
        centers = rng.normal(0.0, 2.5, size=(mixture, min(5, P)))
        X[:, : min(5, P)] += centers[comp]
    X = (X - X.mean(0)) / X.std(0, ddof=1)
    Z = rng.standard_normal((ntot, Q))
    Z = (Z - Z.mean(0)) / Z.std(0, ddof=1)

    # PCA on the predictive block (prcomp)
    Xc = X - X.mean(0)
    _, _, vt = np.linalg.svd(Xc, full_matrices=False)
    pcs = Xc @ vt.T
    sc = lambda v: (v - v.mean()) / v.std(ddof=1)
    metx1 = sc(pcs[:, 0]) + sc(pcs[:, 1 % P]) + 0.5
    metx = np.sign(metx1) * np.abs(metx1) ** 0.2 * 0.45

    def probs(alpha, b1, b2, b3):
        z2 = Z[:, 0]
        z3 = Z[:, 1 % Q]
        eta1 = alpha[0] + b1[0] * metx ** 3 + b2[0] * z2 + b3[0] * z3
        eta2 = alpha[1] + b1[1] * metx ** 3 + b2[1] * z2 + b3[1] * z3
        p1 = 1 / ((np.exp(eta1) + 1) * (np.exp(eta2) + 1))
        p2 = np.exp(eta1) / ((np.exp(eta1) + 1) * (np.exp(eta2) + 1))
        p3 = np.exp(eta2) / (np.exp(eta2) + 1)
        return np.round(np.stack([p1, p2, p3], 1), 4)

    arm_par = [((-0.5, -1.0), (2.0, 2.6)), ((0.7, -1.0), (-1.0, -3.0)), ((0.0, -0.5), (0.5, -1.5))]
    prog = probs((1.0, -0.5), (0, 0), (1.0, 0.5), (0.7, 1.0))
    myprob = []
    for t in range(T):
        a, b = arm_par[t % len(arm_par)]
        pr = probs(a, b, (0, 0), (0, 0)) * prog
        myprob.append(pr / pr.sum(1, keepdims=True))

    treat_all = (np.arange(ntot) % T).astype(np.int32)
    treat = treat_all[:n]
    if ncat > 0:
        xcat_all = rng.integers(0, maxC, size=(ntot, ncat)).astype(np.int32)
    else:
        xcat_all = np.zeros((ntot, 1), dtype=np.int32)

    datasets = []
    cum = [np.cumsum(mp / mp.sum(1, keepdims=True), axis=1) for mp in myprob]
    for r in range(n_datasets):
        y = np.zeros((n, 3))
        if first_replicate is None:
            for i in range(n):
                pr = myprob[treat[i]][i]
                y[i, rng.choice(3, p=pr / pr.sum())] = 1.0
        else:
            u = np.random.default_rng([seed, first_replicate + r]).random(n)
            for i in range(n):
                y[i, min(int(np.searchsorted(cum[treat[i]][i], u[i], side="right")), 2)] = 1.0
        datasets.append(HostDataset(
            treat=treat, y=y, z=Z[:n], xcon=X[:n],
            xcat=xcat_all[:n, : max(ncat, 1)], zpred=Z[n:] if npred else np.zeros((1, Q)),
            xconp=X[n:] if npred else np.zeros((1, P)),
            xcatp=xcat_all[n:, : max(ncat, 1)] if npred else np.zeros((1, 1), np.int32)))
    n_a = np.bincount(treat, minlength=T)
    dims = dict(nobs=n, T=T, D=3, Q=Q, P=P, ncat=ncat, maxC=(maxC if ncat > 0 else 0),
                npred=npred, ntmax=int(n_a.max()))
    return dims, datasets, dict(n_a=n_a, myprob=myprob)


def make_dims(G: int, kcap: int = 0, **kw) -> Dims:
    d = Dims()
    for k, v in kw.items():
        setattr(d, k, int(v))
    d.G = int(G)
    d.kcap = int(kcap)
    return d


def make_shared(dims: Dims, n_a, cohesion: int, catvec=None, grid=None, kmax=None) -> HostShared:
    grid = reference_grid() if grid is None else grid
    catvec = np.full(max(dims.ncat, 1), dims.maxC, np.int32) if catvec is None else catvec
    logV = ngg_logV_table(n_a, dims.ntmax, grid, kmax=kmax) if cohesion == 2 else None
    return HostShared(catvec=catvec, grid=grid, logV=logV)


def initial_partition(treat: np.ndarray, T: int, ntmax: int, nclu_init: int = 5):
    """k-means-free stand-in for R/dm_ppmx_ct.R:214-225: the it-th patient of an arm starts
    in cluster (it mod nclu_init)+1.  Returns (labels T x ntmax F-order int32, nclu [T])."""
    labels = np.zeros((T, ntmax), dtype=np.int32, order="F")
    nclu = np.zeros(T, dtype=np.int32)
    for t in range(T):
        nt = int((treat == t).sum())
        labels[t, :nt] = (np.arange(nt) % nclu_init) + 1
        nclu[t] = min(nclu_init, nt)
    return labels, nclu


def init_beta(ds: HostDataset) -> np.ndarray:
    return _spearman_fdr_beta(np.asarray(ds.y), np.asarray(ds.z))
