"""The drop-in entry point tppmx_dm_ppmx_ct (the reference's dm_ppmx_ct: 42 arguments in, the
18-element list out) against the same run composed from the CPU oracle's pieces: sweep
(:347-860), updates (:866-1055), in-sample fit (:1060-1075), predictive (:1080-1385), save
(:1391-1429), WAIC (:1445-1450)."""
import numpy as np
import pytest

import oracle_binding as ob
from cases import build_case
from treatppmx_b200 import datagen

pytestmark = pytest.mark.gpu


def _reference_args(name, iters, burn, thin, dense_V):
    model, dims, shared, dsets, labels, nclu, betas = build_case(name)
    ds = dsets[0]
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
    args = dict(
        iter=iters, burn=burn, thin=thin, nobs=dims.nobs, treatments=ds.treat.astype(float), PPMx=model.PPMx,
        ncon=dims.P, ncat=dims.ncat, catvec=shared.catvec.astype(float), alpha=model.alpha, grid=shared.grid,
        Vwm=Vwm, cohesion=model.cohesion, CC=model.CC, reuse=model.reuse, consim=model.consim,
        similarity=model.similarity, calibration=model.calibration, coardegree=model.coardegree, y=ds.y, z=ds.z,
        zpred=ds.zpred, noprog=model.noprog, xcon=ds.xcon.ravel(), xcat=ds.xcat.astype(float).ravel(),
        xconp=ds.xconp.ravel(), xcatp=ds.xcatp.astype(float).ravel(), npred=dims.npred, similparam=sp,
        hP0_mu0=[model.hP0_mu0[k] for k in range(dims.D)], hP0_nu0=model.hP0_nu0, hP0_s0=model.hP0_s0,
        hP0_Lambda0=model.hP0_Lambda0, upd_hier=model.upd_hier, initbeta=betas[0], hsp=model.hsp,
        mhtunepar=[model.mhtune_eta, model.mhtune_beta], A=dims.T, n_a=n_a, curr_cluster=labels.astype(float),
        card_cluster=np.zeros_like(labels, dtype=float), ncluster_curr=nclu.astype(float))
    return args, logV, (model, dims, shared, dsets, labels, nclu, betas)


def _oracle_run(case, iters, burn, thin, seed):
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
            for t in range(T):
                out["cl_lab"][t, : n_a[t], ll] = st.labels[t, : n_a[t]]
                out["eta"][:, : st.nclu[t], ll, t] = st.eta_star[:, : st.nclu[t], t]
            ll += 1
    out["eta_acc"] = acc[0] / sumtot[:, None]
    out["beta_acc"] = acc[1].astype(float)
    out["num_treat"] = n_a.astype(float)
    out["WAIC"] = -2 * np.sum(2 * mnllike - np.log(mnlike))
    out["theta_prior"], out["sigma_prior"] = st.Theta.copy(), st.Sigma.copy()
    return out


@pytest.mark.parametrize("name,dense_V", [("default_ngg", True), ("default_ngg", False), ("dp_T3", False), ("nnig", False)])
def test_dm_ppmx_ct_matches_oracle_run(name, dense_V):
    from treatppmx_b200.engine import Context
    from treatppmx_b200.ppmxct import RETURN_NAMES, dm_ppmx_ct
    iters, burn, thin, seed = 36, 12, 3, 99
    args, logV, case = _reference_args(name, iters, burn, thin, dense_V)
    ctx = Context(0)
    got = dm_ppmx_ct(ctx, **args, seed=seed, logV_compact=logV)
    ctx.close()
    want = _oracle_run(case, iters, burn, thin, seed)
    assert sorted(got) == sorted(RETURN_NAMES)
    for k in ("cl_lab", "nclus", "clupred", "yispred", "ypred", "sigma_ngg", "kappa_ngg", "num_treat", "beta_acc"):
        np.testing.assert_array_equal(got[k].reshape(want[k].shape), want[k], err_msg=k)
    for k in ("eta", "beta", "pi", "pipred", "lpml", "eta_acc", "theta_prior", "sigma_prior"):
        np.testing.assert_allclose(got[k].reshape(want[k].shape), want[k], rtol=1e-9, atol=1e-12, err_msg=k)
    assert abs(got["WAIC"] - want["WAIC"]) <= 1e-9 * abs(want["WAIC"])


def test_dm_ppmx_ct_rejects_bad_arguments():
    from treatppmx_b200 import _lib
    from treatppmx_b200.engine import Context
    from treatppmx_b200.ppmxct import dm_ppmx_ct
    args, logV, _ = _reference_args("default_ngg", 10, 2, 1, False)
    ctx = Context(0)
    with pytest.raises(_lib.TppmxError):      # cohesion 2 without any V table
        dm_ppmx_ct(ctx, **args, seed=1, logV_compact=None)
    args["thin"] = 0
    with pytest.raises(_lib.TppmxError):
        dm_ppmx_ct(ctx, **args, seed=1, logV_compact=logV)
    ctx.close()
