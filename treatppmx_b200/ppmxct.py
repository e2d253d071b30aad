"""Host-side mirror of the reference's entry point for the path: dm_ppmx_ct() with the reference's 42
arguments (src/dm_ppmx_ct.cpp:10-23; R call site R/dm_ppmx_ct.R:264-277) and its 18-element
return list (:1453-1471), computed by the CUDA library through tppmx_dm_ppmx_ct.  Marshalling
only: every number comes from the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._abi import c_dp

RETURN_NAMES = ("eta", "beta", "sigma_ngg", "kappa_ngg", "eta_acc", "beta_acc", "cl_lab", "num_treat", "pi",
                "pipred", "yispred", "ypred", "clupred", "nclus", "WAIC", "lpml", "theta_prior", "sigma_prior")


class PpmxctArgs(C.Structure):
    _fields_ = [
        ("iter", C.c_int32), ("burn", C.c_int32), ("thin", C.c_int32), ("nobs", C.c_int32),
        ("treatments", c_dp), ("PPMx", C.c_int32), ("ncon", C.c_int32), ("ncat", C.c_int32),
        ("catvec", c_dp), ("alpha", C.c_double), ("grid", c_dp), ("G", C.c_int32),
        ("Vwm", c_dp), ("Vwm_rows", C.c_int32), ("Vwm_cols", C.c_int32), ("logV_compact", c_dp),
        ("cohesion", C.c_int32), ("CC", C.c_int32), ("reuse", C.c_int32), ("consim", C.c_int32),
        ("similarity", C.c_int32), ("calibration", C.c_int32), ("coardegree", C.c_int32),
        ("y", c_dp), ("D", C.c_int32), ("z", c_dp), ("Q", C.c_int32), ("zpred", c_dp), ("noprog", C.c_int32),
        ("xcon", c_dp), ("xcat", c_dp), ("xconp", c_dp), ("xcatp", c_dp), ("npred", C.c_int32),
        ("similparam", c_dp), ("hP0_mu0", c_dp), ("hP0_nu0", C.c_double), ("hP0_s0", C.c_double),
        ("hP0_Lambda0", C.c_double), ("upd_hier", C.c_int32), ("initbeta", c_dp), ("hsp", C.c_int32),
        ("mhtunepar", c_dp), ("A", C.c_int32), ("n_a", c_dp), ("curr_cluster", c_dp), ("card_cluster", c_dp),
        ("ntmax", C.c_int32), ("ncluster_curr", c_dp), ("seed", C.c_uint64),
    ]


class PpmxctOut(C.Structure):
    _fields_ = [(n, c_dp) for n in RETURN_NAMES]


def dm_ppmx_ct(ctx, iter, burn, thin, nobs, treatments, PPMx, ncon, ncat, catvec, alpha, grid, Vwm, cohesion, CC,
               reuse, consim, similarity, calibration, coardegree, y, z, zpred, noprog, xcon, xcat, xconp, xcatp,
               npred, similparam, hP0_mu0, hP0_nu0, hP0_s0, hP0_Lambda0, upd_hier, initbeta, hsp, mhtunepar, A, n_a,
               curr_cluster, card_cluster, ncluster_curr, *, seed: int = 1, logV_compact=None) -> dict:
    """Same positional arguments as the reference's dm_ppmx_ct (arrays in the reference's layouts:
    matrices column-major, xcon/xcat/xconp/xcatp flattened row-major).  `Vwm` is the dense cube
    (rows x cols x G) or None with `logV_compact` ([A][2][ntmax][G], see include/tppmx.h).
    Returns a dict with the names of the reference's list."""
    keep = []

    def dp(a, order="F"):
        if a is None:
            return c_dp()
        arr = np.require(np.asarray(a, dtype=np.float64), requirements=["F" if order == "F" else "C", "A"])
        keep.append(arr)
        return arr.ctypes.data_as(c_dp)

    y = np.asarray(y, dtype=np.float64)
    D = y.shape[1]
    z = np.zeros((nobs, 0)) if z is None else np.asarray(z, dtype=np.float64)
    Q = z.shape[1] if z.ndim == 2 else 0
    grid = np.asarray(grid, dtype=np.float64)
    G = grid.shape[1]
    cc = np.asarray(curr_cluster, dtype=np.float64)
    ntmax = cc.shape[1]
    nout = (iter - burn) // thin if thin > 0 else 0   # thin < 1 is rejected by the library
    a = PpmxctArgs()
    a.iter, a.burn, a.thin, a.nobs = iter, burn, thin, nobs
    a.treatments = dp(treatments)
    a.PPMx, a.ncon, a.ncat = PPMx, ncon, ncat
    a.catvec = dp(catvec if ncat > 0 else None)
    a.alpha, a.grid, a.G = alpha, dp(grid), G
    if Vwm is not None:
        V = np.asarray(Vwm, dtype=np.float64)
        a.Vwm, a.Vwm_rows, a.Vwm_cols = dp(V), V.shape[0], V.shape[1]
    a.logV_compact = dp(logV_compact, order="C")
    a.cohesion, a.CC, a.reuse, a.consim, a.similarity = cohesion, CC, reuse, consim, similarity
    a.calibration, a.coardegree = calibration, coardegree
    a.y, a.D, a.z, a.Q, a.zpred, a.noprog = dp(y), D, dp(z), Q, dp(zpred if npred > 0 else None), noprog
    a.xcon, a.xconp = dp(xcon, "C"), dp(xconp if npred > 0 else None, "C")
    a.xcat, a.xcatp = dp(xcat if ncat > 0 else None, "C"), dp(xcatp if ncat > 0 and npred > 0 else None, "C")
    a.npred, a.similparam, a.hP0_mu0 = npred, dp(similparam), dp(hP0_mu0)
    a.hP0_nu0, a.hP0_s0, a.hP0_Lambda0, a.upd_hier = hP0_nu0, hP0_s0, hP0_Lambda0, upd_hier
    a.initbeta, a.hsp, a.mhtunepar, a.A, a.n_a = dp(initbeta), hsp, dp(mhtunepar), A, dp(n_a)
    a.curr_cluster, a.card_cluster, a.ntmax, a.ncluster_curr = dp(cc), dp(card_cluster), ntmax, dp(ncluster_curr)
    a.seed = seed

    no = max(nout, 1)
    shapes = dict(
        eta=(D, ntmax, nout, A), beta=(Q * D, nout), sigma_ngg=(A, nout), kappa_ngg=(A, nout), eta_acc=(A, ntmax),
        beta_acc=(max(Q, 1), D), cl_lab=(A, ntmax, nout), num_treat=(A,), pi=(nobs, D, nout),
        pipred=(npred, D, A, nout), yispred=(nobs, D, nout), ypred=(npred, D, A, nout), clupred=(nout * npred, A),
        nclus=(A, nout), WAIC=(1,), lpml=(nout,), theta_prior=(D, A), sigma_prior=(D, D, A))
    res = {k: np.zeros(v if np.prod(v) > 0 else (1,), order="F") for k, v in shapes.items()}
    o = PpmxctOut()
    for k in RETURN_NAMES:
        setattr(o, k, res[k].ctypes.data_as(c_dp))
    L = _lib.load()
    rc = L.tppmx_dm_ppmx_ct(ctx.h, C.byref(a), C.byref(o))
    if rc != 0:
        raise _lib.TppmxError(rc, ctx.err())
    # eta: field (nout x A) of D x ntmax, stored as blocks (ll + nout*t): expose as [D, ntmax, nout, A]
    res["eta"] = res["eta"].reshape((D, ntmax, nout, A), order="F") if nout > 0 else res["eta"]
    res["WAIC"] = float(res["WAIC"][0])
    return res
