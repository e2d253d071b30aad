"""Host-side mirror of the package's second sampler, prior_ppmx_core (src/prior_ppmx.cpp:9-466;
R wrapper R/prior_ppmx.R:32-132): the same reallocation sweep with one arm, continuous
covariates only and NO likelihood term, returning the number of clusters and the cluster sizes
of every saved iteration.  Computed by the CUDA library through tppmx_prior_ppmx_core (the sweep
kernel with tppmx_model.nolik = 1, one save kernel per saved iteration, no host round trip inside
the run); marshalling only here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import c_dp


class PriorArgs(C.Structure):
    _fields_ = [
        ("iter", C.c_int32), ("burn", C.c_int32), ("thin", C.c_int32), ("nobs", C.c_int32), ("PPMx", C.c_int32),
        ("ncon", C.c_int32), ("ncat", C.c_int32), ("alpha", C.c_double), ("sigma", C.c_double),
        ("Vwm", c_dp), ("Vwm_rows", C.c_int32), ("Vwm_cols", C.c_int32),
        ("cohesion", C.c_int32), ("CC", C.c_int32), ("consim", C.c_int32), ("similarity", C.c_int32),
        ("calibration", C.c_int32), ("coardegree", C.c_int32),
        ("xcon", c_dp), ("similparam", c_dp), ("curr_cluster", c_dp), ("card_cluster", c_dp),
        ("ncluster_curr", C.c_int32), ("seed", C.c_uint64), ("n_chains", C.c_int32),
    ]


class PriorOut(C.Structure):
    _fields_ = [("nclu", c_dp), ("njout", c_dp)]


def prior_ppmx_core(ctx, iter, burn, thin, nobs, PPMx, ncon, ncat, alpha, sigma, Vwm, cohesion, CC, consim,
                    similarity, calibration, coardegree, xcon, similparam, curr_cluster, card_cluster,
                    ncluster_curr, *, seed: int = 1, n_chains: int = 1) -> dict:
    """Arguments as the reference's prior_ppmx_core (xcon flattened row-major, curr_cluster 1-based,
    Vwm the matrix of V_{n,k} for the given (sigma, kappa), only rows nobs-2 and nobs-1 are
    read, src/prior_ppmx.cpp:239).  `n_chains` > 1 runs independent replicas (extra leading axis).
    Returns {"nclu": [nout], "njout": [nobs, nout]} (with a leading chain axis if n_chains > 1)."""
    nout = (iter - burn) // thin if thin > 0 else 0
    keep = []

    def dp(a, order="C"):
        if a is None:
            return c_dp()
        arr = np.require(np.asarray(a, dtype=np.float64), requirements=[order, "A"])
        keep.append(arr)
        return arr.ctypes.data_as(c_dp)

    a = PriorArgs()
    a.iter, a.burn, a.thin, a.nobs, a.PPMx, a.ncon, a.ncat = iter, burn, thin, nobs, PPMx, ncon, ncat
    a.alpha, a.sigma = alpha, sigma
    if Vwm is not None:
        V = np.asarray(Vwm, dtype=np.float64)
        a.Vwm, a.Vwm_rows, a.Vwm_cols = dp(V, "F"), V.shape[0], V.shape[1]
    a.cohesion, a.CC, a.consim, a.similarity = cohesion, CC, consim, similarity
    a.calibration, a.coardegree = calibration, coardegree
    a.xcon, a.similparam = dp(np.asarray(xcon, dtype=np.float64).ravel()), dp(similparam)
    a.curr_cluster, a.card_cluster = dp(curr_cluster), dp(card_cluster)
    a.ncluster_curr, a.seed, a.n_chains = int(ncluster_curr), seed, n_chains
    nclu = np.zeros((n_chains, max(nout, 1)))
    njout = np.zeros((n_chains, max(nout, 1), nobs))      # C-order [chain][ll][j] == col-major nobs x nout per chain
    o = PriorOut()
    o.nclu, o.njout = nclu.ctypes.data_as(c_dp), njout.ctypes.data_as(c_dp)
    L = _lib.load()
    rc = L.tppmx_prior_ppmx_core(ctx.h, C.byref(a), C.byref(o))
    if rc != 0:
        raise _lib.TppmxError(rc, ctx.err())
    nclu, njout = nclu[:, :nout], njout[:, :nout, :].transpose(0, 2, 1)
    if n_chains == 1:
        return {"nclu": nclu[0], "njout": njout[0]}
    return {"nclu": nclu, "njout": njout}
