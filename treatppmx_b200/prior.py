"""Host-side mirror of the package's second sampler, prior_ppmx_core (src/prior_ppmx.cpp:9-466;
R wrapper R/prior_ppmx.R:32-132): the same reallocation sweep with one arm, continuous
covariates only and NO likelihood term, returning the number of clusters and the cluster sizes
of every saved iteration.  On the device this is the sweep kernel with tppmx_model.nolik = 1
(SURVEY.md §8f-3); marshalling only here."""
from __future__ import annotations

import numpy as np

from ._abi import HostDataset, HostShared, default_model
from .datagen import make_dims
from .engine import Batch


def prior_ppmx_core(ctx, iter, burn, thin, nobs, PPMx, ncon, ncat, alpha, sigma, Vwm, cohesion, CC, consim,
                    similarity, calibration, coardegree, xcon, similparam, curr_cluster, card_cluster,
                    ncluster_curr, *, seed: int = 1, n_chains: int = 1) -> dict:
    """Arguments as the reference's prior_ppmx_core (xcon flattened row-major, curr_cluster 1-based,
    Vwm the n x n matrix of V_{n,k} for the given (sigma, kappa), only rows nobs-2 and nobs-1 are
    read, src/prior_ppmx.cpp:239).  `n_chains` > 1 runs independent replicas (extra leading axis).
    Returns {"nclu": [nout], "njout": [nobs, nout]} (with a leading chain axis if n_chains > 1)."""
    assert ncat == 0, "prior_ppmx_core ignores categorical covariates"
    nout = (iter - burn) // thin
    x = np.asarray(xcon, dtype=np.float64).reshape(nobs, max(ncon, 1))[:, :ncon]
    model = default_model(PPMx=PPMx, cohesion=cohesion, similarity=similarity, consim=consim, calibration=calibration,
                          coardegree=coardegree, CC=CC, nolik=1, alpha=alpha, similparam=tuple(similparam))
    dims = make_dims(G=1, kcap=0, nobs=nobs, T=1, D=1, Q=0, P=ncon, ncat=0, maxC=0, npred=0, ntmax=nobs)
    logV = None
    if cohesion == 2:
        V = np.asarray(Vwm, dtype=np.float64)
        logV = np.full((1, 2, nobs, 1), -np.inf)
        with np.errstate(divide="ignore"):
            for r in (0, 1):
                row = nobs - 2 + r
                if 0 <= row < V.shape[0]:
                    k = min(nobs, V.shape[1])
                    logV[0, r, :k, 0] = np.log(V[row, :k])
    shared = HostShared(catvec=np.zeros(1, np.int32), grid=np.array([[sigma], [0.0]]), logV=logV)
    ds = HostDataset(treat=np.zeros(nobs, np.int32), y=np.ones((nobs, 1)), z=np.zeros((nobs, 0)), xcon=x,
                     xcat=np.zeros((nobs, 1), np.int32), zpred=np.zeros((1, 0)), xconp=np.zeros((1, max(ncon, 1))),
                     xcatp=np.zeros((1, 1), np.int32))
    b = Batch(ctx, model, dims, shared, [ds], chain_dataset=[0] * n_chains)
    labels = np.asarray(curr_cluster, dtype=np.int32).reshape(1, nobs)
    b.init(seed, labels, np.array([ncluster_curr], np.int32), np.zeros(0))
    nclu = np.zeros((n_chains, nout))
    njout = np.zeros((n_chains, nobs, nout))
    ll = 0
    for l in range(iter):
        b.sweep(seed, l, 1)
        if l > burn - 1 and (l + 1) % thin == 0 and ll < nout:   # src/prior_ppmx.cpp:455
            for c in range(n_chains):
                st = b.get_state(c)
                nclu[c, ll] = st.nclu[0]
                njout[c, : st.nclu[0], ll] = st.nj[0, : st.nclu[0]]
            ll += 1
    b.close()
    if n_chains == 1:
        return {"nclu": nclu[0], "njout": njout[0]}
    return {"nclu": nclu, "njout": njout}
