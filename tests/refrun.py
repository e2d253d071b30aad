"""Whole runs of the reference's entry point dm_ppmx_ct (src/dm_ppmx_ct.cpp:10-1471), three ways:

* `reference_args`  - the 42 arguments of a parity case, in the reference's layouts;
* `oracle_run`      - the run composed from the CPU oracle's pieces (optionally recording the
                      variates it consumes on a tape);
* `RefLib`          - ctypes binding of oracle/_ref/libref.so, the reference's OWN sources compiled
                      against the stand-in headers of oracle/refshim/ and driven by that tape.

TEST INFRASTRUCTURE ONLY (imported by tests/ and scripts/make_reference_golden.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle_binding as ob
from cases import build_case

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(_ROOT, "oracle", "_ref", "libref.so")
SHIM_SO = os.path.join(_ROOT, "oracle", "_ref", "libshim.so")   # the CUDA library's Rcpp shim, same driver
REF_SRC = "/root/reference/src"

RETURN_NAMES = ("eta", "beta", "sigma_ngg", "kappa_ngg", "eta_acc", "beta_acc", "cl_lab", "num_treat", "pi",
                "pipred", "yispred", "ypred", "clupred", "nclus", "WAIC", "lpml", "theta_prior", "sigma_prior")


def reference_args(name, iters, burn, thin, dense_V):
    model, dims, shared, dsets, labels, nclu, betas = build_case(name)
    args, logV = reference_args_of(model, dims, shared, dsets[0], labels, nclu, betas[0], iters, burn, thin, dense_V)
    return args, logV, (model, dims, shared, dsets, labels, nclu, betas)


def reference_args_of(model, dims, shared, ds, labels, nclu, beta0, iters, burn, thin, dense_V):
    """The reference's 42 arguments (src/dm_ppmx_ct.cpp:10-23) for one dataset of a problem."""
    n_a = np.bincount(ds.treat, minlength=dims.T).astype(float)
    Vwm = logV = None
    if model.cohesion == 2:
        if dense_V:   # the dense (max n_t + 1)^2 x G cube of R/dm_ppmx_ct.R:242-254, rows n_t-1 and n_t filled
            nmax = int(n_a.max())
            Vwm = np.zeros((nmax + 1, nmax + 1, dims.G), order="F")
            for t in range(dims.T):
                for r in (0, 1):
                    row = int(n_a[t]) - 2 + r
                    Vwm[row, : dims.ntmax, :] = np.exp(shared.logV[t, r])
        else:
            logV = shared.logV
    sp = [model.similparam[i] for i in range(7)]
    card = np.zeros_like(labels, dtype=float)   # R/dm_ppmx_ct.R:223: card_cluster[a, j] = table(cluster)[j]
    for t in range(dims.T):
        for j in range(int(nclu[t])):
            card[t, j] = np.sum(labels[t, : int(n_a[t])] == j + 1)
    args = dict(
        iter=iters, burn=burn, thin=thin, nobs=dims.nobs, treatments=ds.treat.astype(float), PPMx=model.PPMx,
        ncon=dims.P, ncat=dims.ncat, catvec=shared.catvec.astype(float), alpha=model.alpha, grid=shared.grid,
        Vwm=Vwm, cohesion=model.cohesion, CC=model.CC, reuse=model.reuse, consim=model.consim,
        similarity=model.similarity, calibration=model.calibration, coardegree=model.coardegree, y=ds.y, z=ds.z,
        zpred=ds.zpred, noprog=model.noprog, xcon=ds.xcon.ravel(), xcat=ds.xcat.astype(float).ravel(),
        xconp=ds.xconp.ravel(), xcatp=ds.xcatp.astype(float).ravel(), npred=dims.npred, similparam=sp,
        hP0_mu0=[model.hP0_mu0[k] for k in range(dims.D)], hP0_nu0=model.hP0_nu0, hP0_s0=model.hP0_s0,
        hP0_Lambda0=model.hP0_Lambda0, upd_hier=model.upd_hier, initbeta=beta0, hsp=model.hsp,
        mhtunepar=[model.mhtune_eta, model.mhtune_beta], A=dims.T, n_a=n_a, curr_cluster=labels.astype(float),
        card_cluster=card, ncluster_curr=nclu.astype(float))
    return args, logV


def oracle_run(case, iters, burn, thin, seed):
    model, dims, shared, dsets, labels, nclu, betas = case
    pb = ob.Problem(model, dims, shared, dsets[0])
    st = pb.init_state(seed, 0, labels, nclu, betas[0])
    acc = pb.new_acc()
    nout = (iters - burn) // thin
    T, D, nt, n, Q, npred = dims.T, dims.D, dims.ntmax, dims.nobs, dims.Q, dims.npred
    n_a = np.bincount(dsets[0].treat, minlength=T)
    out = dict(eta=np.zeros((D, nt, nout, T)), beta=np.zeros((Q * D, nout)), sigma_ngg=np.zeros((T, nout)),
               kappa_ngg=np.zeros((T, nout)), cl_lab=np.zeros((T, nt, nout)), pi=np.zeros((n, D, nout)),
               pipred=np.zeros((npred, D, T, nout)), yispred=np.zeros((n, D, nout)), ypred=np.zeros((npred, D, T, nout)),
               clupred=np.zeros((nout * npred, T)), nclus=np.zeros((T, nout)), lpml=np.zeros(nout))
    mnlike, mnllike, sumtot, ll = np.zeros(n), np.zeros(n), np.zeros(T), 0
    for l in range(iters):
        pb.iterate(st, seed, 0, l, 1, acc)
        sumtot += st.nclu
        if l > burn - 1 and (l + 1) % thin == 0:
            pi, isp, like = pb.insample(st, seed, 0, l)
            mnlike += like / nout
            mnllike += np.log(like) / nout
            out["lpml"][ll] = np.log(like).sum()
            pip, yp, pc = pb.predict(st, seed, 0, l)
            out["pi"][:, :, ll], out["yispred"][:, :, ll] = pi, isp
            out["pipred"][..., ll], out["ypred"][..., ll] = pip, yp
            out["clupred"][ll * npred:(ll + 1) * npred, :] = pc
            out["sigma_ngg"][:, ll], out["kappa_ngg"][:, ll], out["nclus"][:, ll] = st.sigma, st.kappa, st.nclu
            out["beta"][:, ll] = st.beta
            mymat = np.zeros((D, nt))   # :1421-1427: one scratch matrix per save, NOT cleared between arms
            for t in range(T):
                out["cl_lab"][t, : n_a[t], ll] = st.labels[t, : n_a[t]]
                mymat[:, : st.nclu[t]] = st.eta_star[:, : st.nclu[t], t]
                out["eta"][:, :, ll, t] = mymat
            ll += 1
    out["eta_acc"] = acc[0] / sumtot[:, None]
    out["beta_acc"] = acc[1].astype(float)
    out["num_treat"] = n_a.astype(float)
    out["WAIC"] = -2 * np.sum(2 * mnllike - np.log(mnlike))
    out["theta_prior"], out["sigma_prior"] = st.Theta.copy(), st.Sigma.copy()
    return out




# ---- variate tape of the oracle (oracle/philox_ref.h) --------------------------------------------
class Tape:
    """with Tape() as t: <oracle calls>; then t.kind / t.val / t.aux hold every variate the oracle
    consumed, in call order."""

    def __enter__(self):
        L = ob.lib()
        L.oracle_tape_stop.restype = C.c_long
        L.oracle_tape_start()
        return self

    def __exit__(self, *exc):
        L = ob.lib()
        n = L.oracle_tape_stop()
        self.kind, self.val, self.aux = np.zeros(n), np.zeros(n), np.zeros(n)
        dp = C.POINTER(C.c_double)
        L.oracle_tape_get(self.kind.ctypes.data_as(dp), self.val.ctypes.data_as(dp),
                          self.aux.ctypes.data_as(dp), C.c_long(n))
        return False

    def __len__(self):
        return len(self.kind)


# ---- the reference's own sources -----------------------------------------------------------------
def reference_available() -> bool:
    return os.path.exists(REF_SO) or os.path.isdir(REF_SRC)


class SeedTape:
    """Two uniforms that make the Rcpp shim draw exactly `seed` as its Philox key
    (seed = floor(u1 2^32) << 32 | floor(u2 2^32), treatppmx_shim.cpp:seed_from_r_rng)."""

    def __init__(self, seed: int):
        hi, lo = (seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF
        self.kind = np.array([1.0, 1.0])
        self.val = np.array([(hi + 0.5) / 4294967296.0, (lo + 0.5) / 4294967296.0])
        self.aux = np.zeros(2)

    def __len__(self):
        return 2


class RefLib:
    def __init__(self, so: str = REF_SO):
        """so = REF_SO: the reference's own sources; so = SHIM_SO: the CUDA library's Rcpp shim behind
        the same driver (make -C oracle shim), i.e. the reference's C++ signatures over the CUDA path."""
        if so == REF_SO and os.path.isdir(REF_SRC):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        if so == SHIM_SO and not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "shim"], stdout=subprocess.DEVNULL)
        self.L = C.CDLL(so)
        L, d, i = self.L, C.c_double, C.c_int
        dp = C.POINTER(C.c_double)
        L.ref_last_error.restype = C.c_char_p
        L.ref_tape_pos.restype = C.c_long
        L.ref_tape_set.argtypes = [dp, dp, dp, C.c_long]
        for name, args in () if so == SHIM_SO else (("ref_dinvgamma", [d, d, d, i]), ("ref_dN_IG", [d, d, d, d, d, d, i]),
                           ("ref_gsimconNN", [d, d, d, d, d, i, i, i]),
                           ("ref_gsimconNNIG", [d, d, d, d, d, d, i, i, i]), ("ref_gsimcatDM", [dp, dp, i, i, i]),
                           ("ref_calculate_gamma", [dp, i, i, dp, i, i, dp, i, i, i, i]),
                           ("ref_dmvnrm_log", [dp, dp, dp, i])):
            getattr(L, name).restype = d
            getattr(L, name).argtypes = args
        self._keep = []

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def err(self) -> str:
        return self.L.ref_last_error().decode()

    def set_tape(self, kind, val, aux=None):
        kind = np.ascontiguousarray(kind, dtype=np.float64)
        val = np.ascontiguousarray(val, dtype=np.float64)
        aux = np.zeros_like(val) if aux is None else np.ascontiguousarray(aux, dtype=np.float64)
        self._keep = [kind, val, aux]
        self.L.ref_tape_set(self._p(kind), self._p(val), self._p(aux), len(val))

    def rng_free(self, seed: int):
        """no tape on the calling thread: variates from the driver's free-running generator (timing)"""
        self.L.ref_rng_free(C.c_uint64(seed))

    def tape_pos(self) -> int:
        return self.L.ref_tape_pos()

    def prior_ppmx_core(self, iter, burn, thin, nobs, PPMx, ncon, ncat, alpha, sigma, Vwm, cohesion, CC, consim,
                        similarity, calibration, coardegree, xcon, similparam, curr_cluster, card_cluster,
                        ncluster_curr, tape) -> dict:
        """prior_ppmx_core(21 arguments) -> {"nclu": [nout], "njout": [nobs, nout]}"""
        f = lambda x: np.require(np.asarray(x, dtype=np.float64), requirements=["F", "A", "O"])
        V, xc, sp, cc, card = f(Vwm), f(np.ravel(xcon)), f(similparam), f(curr_cluster), f(card_cluster)
        nout = (iter - burn) // thin
        nclu, njout = np.zeros(nout), np.zeros((nobs, nout), order="F")
        if isinstance(tape, int):
            self.rng_free(tape)
        else:
            self.set_tape(tape.kind, tape.val, tape.aux)
        i, d = C.c_int, C.c_double
        rc = self.L.ref_prior_ppmx_core(i(iter), i(burn), i(thin), i(nobs), i(PPMx), i(ncon), i(ncat), d(alpha), d(sigma),
                                        self._p(V), i(V.shape[0]), i(V.shape[1]), i(cohesion), i(CC), i(consim),
                                        i(similarity), i(calibration), i(coardegree), self._p(xc), self._p(sp),
                                        self._p(cc), self._p(card), i(ncluster_curr), self._p(nclu), self._p(njout))
        if rc != 0:
            raise RuntimeError(f"prior_ppmx_core failed: {self.err()}")
        return {"nclu": nclu, "njout": njout}

    def dm_ppmx_ct(self, args: dict, tape, outputs: bool = True) -> dict:
        """The reference's dm_ppmx_ct(42 arguments) -> its 18-element list, with `tape` as R's RNG
        (tape = an int seeds the free-running stand-in generator instead: timing runs)."""
        a = args
        f = lambda x, order="F": np.require(np.asarray(x, dtype=np.float64), requirements=[order, "A", "O"])
        y, z = f(a["y"]), f(a["z"])
        n, D, Q, T, npred = a["nobs"], y.shape[1], z.shape[1], a["A"], a["npred"]
        grid = f(a["grid"])
        G = grid.shape[1]
        cc = f(a["curr_cluster"])
        ntmax = cc.shape[1]
        nout = (a["iter"] - a["burn"]) // a["thin"]
        V = f(a["Vwm"]) if a["Vwm"] is not None else None
        catvec = f(a["catvec"]) if a["ncat"] > 0 else np.zeros(1)   # R/dm_ppmx_ct.R:153 passes 0
        out = dict(eta=np.zeros((D, ntmax, nout, T), order="F"), beta=np.zeros((Q * D, nout), order="F"),
                   sigma_ngg=np.zeros((T, nout), order="F"), kappa_ngg=np.zeros((T, nout), order="F"),
                   eta_acc=np.zeros((T, ntmax), order="F"), beta_acc=np.zeros((Q, D), order="F"),
                   cl_lab=np.zeros((T, ntmax, nout), order="F"), num_treat=np.zeros(T),
                   pi=np.zeros((n, D, nout), order="F"), pipred=np.zeros((npred, D, T, nout), order="F"),
                   yispred=np.zeros((n, D, nout), order="F"), ypred=np.zeros((npred, D, T, nout), order="F"),
                   clupred=np.zeros((nout * npred, T), order="F"), nclus=np.zeros((T, nout), order="F"),
                   WAIC=np.zeros(1), lpml=np.zeros(nout), theta_prior=np.zeros((D, T), order="F"),
                   sigma_prior=np.zeros((D, D, T), order="F"))
        ins = [f(a["treatments"]), catvec, grid, V, y, z, f(a["zpred"]), f(a["xcon"]), f(a["xcat"]), f(a["xconp"]),
               f(a["xcatp"]), f(a["similparam"]), f(a["hP0_mu0"]), f(a["initbeta"]), f(a["mhtunepar"]), f(a["n_a"]),
               cc, f(a["card_cluster"]), f(a["ncluster_curr"])]
        P = lambda x: self._p(x) if x is not None else C.POINTER(C.c_double)()
        (treat, catvec, grid, V, y, z, zpred, xcon, xcat, xconp, xcatp, simil, mu0, initbeta, mht, n_a, cc, card,
         nclu) = ins
        if isinstance(tape, int):
            self.rng_free(tape)
        else:
            self.set_tape(tape.kind, tape.val, tape.aux)
        i, d = C.c_int, C.c_double
        rc = self.L.ref_dm_ppmx_ct(
            i(a["iter"]), i(a["burn"]), i(a["thin"]), i(n), P(treat), i(a["PPMx"]), i(a["ncon"]), i(a["ncat"]),
            P(catvec), i(len(catvec)), d(a["alpha"]), P(grid), i(G), P(V), i(V.shape[0] if V is not None else 0),
            i(V.shape[1] if V is not None else 0), i(a["cohesion"]), i(a["CC"]), i(a["reuse"]), i(a["consim"]),
            i(a["similarity"]), i(a["calibration"]), i(a["coardegree"]), P(y), i(D), P(z), i(Q), P(zpred),
            i(a["noprog"]), P(xcon), P(xcat), P(xconp), P(xcatp), i(npred), P(simil), P(mu0), d(a["hP0_nu0"]),
            d(a["hP0_s0"]), d(a["hP0_Lambda0"]), i(a["upd_hier"]), P(initbeta), i(a["hsp"]), P(mht), i(T), P(n_a),
            P(cc), P(card), P(nclu), i(ntmax),
            *[P(out[k]) for k in RETURN_NAMES])
        if rc != 0:
            raise RuntimeError(f"reference run failed after {self.tape_pos()} draws: {self.err()}")
        out["tape_used"] = self.tape_pos()
        out["WAIC"] = float(out["WAIC"][0])
        return out
