"""The oracle against the reference's OWN sources executed here.

oracle/_ref/libref.so is /root/reference/src/{utils_ct,dm_ppmx_ct,prior_ppmx}.cpp compiled
unmodified against the stand-in headers of oracle/refshim/ (R, Rcpp and Armadillo are absent from
this image).  Two levels:

* function by function (src/utils_ct.cpp): similarity / density / likelihood helpers on random
  inputs, <= 1e-13 relative (the Metropolis / horseshoe updates are covered by the whole runs);
* whole runs of dm_ppmx_ct (src/dm_ppmx_ct.cpp:10-1471): the oracle records every variate it
  consumes, in call order, on a tape; the reference is then run with that tape as its R RNG.  It
  must consume the tape exactly (same kinds, same gamma shapes, same count) and return the same
  18-element list: labels, counts, draws bit-identical, real-valued outputs <= 1e-12 relative.

Skipped only where neither /root/reference nor a prebuilt libref.so exists."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import refrun
from cases import CASES

pytestmark = pytest.mark.skipif(not refrun.reference_available(), reason="no reference sources and no prebuilt oracle/_ref")

EXACT = ("cl_lab", "nclus", "clupred", "yispred", "ypred", "sigma_ngg", "kappa_ngg", "num_treat", "beta_acc")
CLOSE = ("eta", "beta", "pi", "pipred", "lpml", "eta_acc", "theta_prior", "sigma_prior")


@pytest.fixture(scope="module")
def ref():
    return refrun.RefLib()


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---- the stand-in containers ---------------------------------------------------------------------------
def test_shim_linear_algebra_matches_numpy(ref):
    """The pin is only as good as the stand-in headers: inverse, Cholesky, covariance, means and a
    product through sub-views against numpy."""
    rng = np.random.default_rng(11)
    dp = C.POINTER(C.c_double)
    ref.L.ref_shim_linalg.argtypes = [dp, C.c_int, dp, C.c_int, dp, dp, dp, dp, dp]
    for n, m in ((1, 3), (3, 10), (5, 40)):
        B = rng.normal(size=(n, n))
        A = np.asfortranarray(B @ B.T + n * np.eye(n))
        X = np.asfortranarray(rng.normal(size=(m, n)))
        inv, ch, cov, prod = (np.zeros((n, n), order="F") for _ in range(4))
        mean = np.zeros(n)
        assert ref.L.ref_shim_linalg(_p(A), n, _p(X), m, _p(inv), _p(ch), _p(cov), _p(mean), _p(prod)) == 0, ref.err()
        np.testing.assert_allclose(inv, np.linalg.inv(A), rtol=1e-11)
        np.testing.assert_allclose(ch, np.linalg.cholesky(A).T, rtol=1e-12, atol=1e-14)   # upper: A = R'R
        np.testing.assert_allclose(cov, np.cov(X, rowvar=False, bias=True).reshape(n, n), rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(mean, X.mean(axis=0), rtol=1e-13)
        np.testing.assert_allclose(prod, A @ A, rtol=1e-12)


# ---- src/utils_ct.cpp, function by function ---------------------------------------------------------
def test_similarity_functions_match_reference(ref):
    rng = np.random.default_rng(1)
    O, R = ob.lib(), ref.L
    for _ in range(4000):
        n = int(rng.integers(1, 400))
        x = rng.normal(rng.normal(0, 2), rng.uniform(0.2, 3), n)
        sx, sx2 = float(x.sum()), float((x * x).sum())
        m0, s20, v2 = rng.normal(), rng.uniform(0.1, 20), rng.uniform(0.1, 5)
        k0, nu0 = rng.uniform(0.1, 5), rng.uniform(0.5, 8)
        for DD in (0, 1):
            for lg in (1, 0):
                a, b = O.oracle_gsimconNN(m0, v2, s20, sx, sx2, n, DD, lg), R.ref_gsimconNN(m0, v2, s20, sx, sx2, n, DD, lg)
                assert a == pytest.approx(b, rel=1e-13, abs=1e-300), ("NN", n, DD, lg)
                a = O.oracle_gsimconNNIG(m0, k0, nu0, s20, sx, sx2, n, DD, lg)
                b = R.ref_gsimconNNIG(m0, k0, nu0, s20, sx, sx2, n, DD, lg)
                assert a == pytest.approx(b, rel=1e-13, abs=1e-300), ("NNIG", n, DD, lg)
        y, al, be = rng.uniform(0.01, 30), rng.uniform(0.1, 20), rng.uniform(0.1, 20)
        assert O.oracle_dinvgamma(y, al, be, 1) == pytest.approx(R.ref_dinvgamma(y, al, be, 1), rel=1e-13)
        mu, s2 = rng.normal(0, 5), rng.uniform(0.05, 10)
        assert O.oracle_dN_IG(mu, s2, m0, k0, al, be, 1) == pytest.approx(R.ref_dN_IG(mu, s2, m0, k0, al, be, 1), rel=1e-13)


def test_gsimcatDM_matches_reference_incl_empty_cluster(ref):
    rng = np.random.default_rng(2)
    O, R = ob.lib(), ref.L
    for _ in range(3000):
        Cn = int(rng.integers(1, 9))
        counts = rng.integers(0, 40, Cn).astype(float)
        if rng.random() < 0.15:
            counts[:] = 0            # src/utils_ct.cpp:491: an empty cluster returns log(1)
        dirw = np.full(Cn, rng.uniform(0.05, 3))
        for DD in (0, 1):
            a = O.oracle_gsimcatDM(_p(counts), _p(dirw), Cn, DD, 1)
            b = R.ref_gsimcatDM(_p(counts), _p(dirw), Cn, DD, 1)
            assert a == pytest.approx(b, rel=1e-13, abs=1e-13)


def test_dmvnrm_and_calculate_gamma_match_reference(ref):
    rng = np.random.default_rng(3)
    O, R = ob.lib(), ref.L
    for _ in range(500):
        n = int(rng.integers(1, 6))
        A = rng.normal(size=(n, n))
        S = np.asfortranarray(A @ A.T + n * np.eye(n))
        x, m = rng.normal(size=n), rng.normal(size=n)
        assert O.oracle_dmvnrm_log(_p(x), _p(m), _p(S), n) == pytest.approx(R.ref_dmvnrm_log(_p(x), _p(m), _p(S), n), rel=1e-12)
    # calculate_gamma (src/utils_ct.cpp:497-535) incl. the "eta has one column -> treat as a vector" branch
    D, Q, nobs = 3, 2, 7
    Z = np.asfortranarray(rng.normal(size=(nobs, Q)))
    beta = rng.normal(size=Q * D)
    for ncols in (1, 5):
        eta = np.asfortranarray(rng.normal(size=(D, ncols)))
        for k in range(D):
            for i in range(nobs):
                clu = int(rng.integers(0, ncols))
                want = eta[k, clu] + sum(beta[k + q * D] * Z[i, q] for q in range(Q))
                got = R.ref_calculate_gamma(_p(eta), D, ncols, _p(Z), nobs, Q, _p(beta), clu, k, i, 1)
                assert got == pytest.approx(want, rel=1e-14)
                assert R.ref_calculate_gamma(_p(eta), D, ncols, _p(Z), nobs, Q, _p(beta), clu, k, i, 0) == pytest.approx(np.exp(want), rel=1e-14)


# ---- whole runs on the oracle's tape -------------------------------------------------------------
# (case, model overrides): the reference cannot run double-dipper NN or categorical covariates with
# NGG cohesion (quirk 8: out-of-range nj_curr index / whole-matrix gsimcatDM argument in the
# (sigma, kappa) update, src/dm_ppmx_ct.cpp:928, 942), so those cases run with DP cohesion here.
RUNS = [("default_ngg", {}), ("dp_T3", {}), ("nnig", {}), ("nnig_dd", {}), ("calib1", {}), ("calib2_sqrt", {}),
        ("calib2", {}), ("noreuse", {}), ("ppm_only", {}), ("categorical", {}), ("Q3_D3", {}),
        ("double_dipper", {"cohesion": 1}), ("cat_dd_calib1", {"cohesion": 1}),
        # the shapes of the fast kernel's special instantiations (n = 152, T = 3, P = 10): calibration 1 with DP cohesion,
        # NNIG with NGG cohesion, three categorical covariates, double dipper with coarsened similarity
        ("calib1_dp_T3", {}), ("nnig_ngg_T3", {}), ("categorical_T3", {}), ("dd_dp_T3", {}), ("cat_calib2", {}), ("noreuse_T3", {}),
        # the switches of the update part: no prognostic coefficients, frozen hierarchy, no horseshoe
        ("default_ngg", {"noprog": 1}), ("dp_T3", {"upd_hier": 0}), ("default_ngg", {"hsp": 0}),
        ("dp_T3", {"noprog": 1, "upd_hier": 0, "hsp": 0})]


def _run_both(ref, name, over, iters, burn, thin, seed):
    saved = CASES[name]
    CASES[name] = (dict(saved[0], **over), saved[1])
    try:
        args, _, case = refrun.reference_args(name, iters, burn, thin, True)
    finally:
        CASES[name] = saved
    with refrun.Tape() as tape:
        want = refrun.oracle_run(case, iters, burn, thin, seed)
    got = ref.dm_ppmx_ct(args, tape)
    return got, want, tape


def _compare(got, want, tape):
    assert got["tape_used"] == len(tape), "the reference left variates of the oracle's tape unused"
    for k in EXACT:
        np.testing.assert_array_equal(got[k].reshape(want[k].shape), want[k], err_msg=k)
    for k in CLOSE:
        np.testing.assert_allclose(got[k].reshape(want[k].shape), want[k], rtol=1e-12, atol=1e-14, err_msg=k)
    assert got["WAIC"] == pytest.approx(want["WAIC"], rel=1e-12)


@pytest.mark.parametrize("name,over", RUNS, ids=[r[0] + "".join(f"-{k}{v}" for k, v in r[1].items()) for r in RUNS])
def test_reference_run_on_the_oracle_tape(ref, name, over):
    got, want, tape = _run_both(ref, name, over, iters=40, burn=10, thin=3, seed=11)
    _compare(got, want, tape)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("name", ["default_ngg", "dp_T3", "noreuse"])
def test_reference_run_other_seeds(ref, name, seed):
    got, want, tape = _run_both(ref, name, {}, iters=60, burn=20, thin=4, seed=1000 + seed)
    _compare(got, want, tape)


def test_reference_run_long_default(ref):
    got, want, tape = _run_both(ref, "default_ngg", {}, iters=400, burn=100, thin=5, seed=5)
    _compare(got, want, tape)
    assert len(tape) > 300000


def test_tape_checker_detects_a_misordered_draw(ref):
    """The replay is a real check: swap the kinds of two neighbouring records and the run must fail."""
    args, _, case = refrun.reference_args("default_ngg", 6, 2, 1, True)
    with refrun.Tape() as tape:
        refrun.oracle_run(case, 6, 2, 1, 3)
    i = int(np.nonzero((tape.kind[:-1] == 2) & (tape.kind[1:] == 1))[0][10])
    tape.kind[[i, i + 1]] = tape.kind[[i + 1, i]]
    with pytest.raises(RuntimeError, match="asks for kind"):
        ref.dm_ppmx_ct(args, tape)
    tape.kind[[i, i + 1]] = tape.kind[[i + 1, i]]
    tape.kind, tape.val, tape.aux = tape.kind[:-5], tape.val[:-5], tape.aux[:-5]
    with pytest.raises(RuntimeError, match="tape exhausted"):
        ref.dm_ppmx_ct(args, tape)


def test_reference_rejects_double_dipper_nn_with_ngg(ref):
    """Quirk 8 as the reference behaves: the (sigma, kappa) update indexes nj_curr(tt, j*ncon+p) and runs
    out of bounds (Armadillo bounds checks are on in the package build: no ARMA_NO_DEBUG in src/Makevars)."""
    args, _, case = refrun.reference_args("double_dipper", 4, 0, 1, True)
    with refrun.Tape() as tape:
        refrun.oracle_run(case, 4, 0, 1, 3)
    with pytest.raises(RuntimeError, match="out of bounds"):
        ref.dm_ppmx_ct(args, tape)


# ---- src/prior_ppmx.cpp:9-466 ------------------------------------------------------------------------
@pytest.mark.parametrize("cohesion,similarity,consim,calibration", [(1, 1, 1, 0), (2, 1, 1, 0), (1, 2, 2, 0), (2, 1, 2, 2), (1, 1, 1, 1)])
def test_reference_prior_ppmx_core_on_the_oracle_tape(ref, cohesion, similarity, consim, calibration):
    """The package's second sampler is the same sweep without a likelihood: the reference's
    prior_ppmx_core against the oracle's sweep with nolik = 1.  Its only variates are the
    allocation uniforms (:415), so the tape is the oracle's uniforms after the prologue."""
    from treatppmx_b200 import datagen
    from treatppmx_b200._abi import HostDataset, HostShared, default_model
    n, P, iters, burn, thin, seed, CC = 30, 3, 60, 10, 2, 17, 1
    rng = np.random.default_rng(2)
    x = rng.standard_normal((n, P))
    sigma, kappa = 0.2, 1.5
    V = np.zeros((n, n), order="F")
    for r in (n - 2, n - 1):
        V[r, : r + 1] = np.exp(datagen.ngg_logV(r + 1, r + 1, sigma, kappa))
    curr = (np.arange(n) % 4) + 1
    card = np.zeros(n)
    card[:4] = np.bincount(curr)[1:]
    sp = (0.0, 10.0, 0.5, 1.0, 2.0, 0.0, 0.1)
    model = default_model(PPMx=1, cohesion=cohesion, similarity=similarity, consim=consim, calibration=calibration,
                          coardegree=1, CC=CC, nolik=1, alpha=1.0, similparam=sp)
    dims = datagen.make_dims(G=1, kcap=0, nobs=n, T=1, D=1, Q=0, P=P, ncat=0, maxC=0, npred=0, ntmax=n)
    logV = None
    if cohesion == 2:
        logV = np.full((1, 2, n, 1), -np.inf)
        with np.errstate(divide="ignore"):
            logV[0, 0, :, 0], logV[0, 1, :, 0] = np.log(V[n - 2]), np.log(V[n - 1])
    shared = HostShared(catvec=np.zeros(1, np.int32), grid=np.array([[sigma], [0.0]]), logV=logV)
    ds = HostDataset(treat=np.zeros(n, np.int32), y=np.ones((n, 1)), z=np.zeros((n, 0)), xcon=x,
                     xcat=np.zeros((n, 1), np.int32), zpred=np.zeros((1, 0)), xconp=np.zeros((1, P)),
                     xcatp=np.zeros((1, 1), np.int32))
    pb = ob.Problem(model, dims, shared, ds)
    st = pb.init_state(seed, 0, curr.reshape(1, n).astype(np.int32), np.array([4], np.int32), np.zeros(0))
    want_K, want_nj = [], []
    with refrun.Tape() as tape:
        for l in range(iters):
            pb.sweep(st, seed, 0, l, 1)
            if l > burn - 1 and (l + 1) % thin == 0:
                want_K.append(int(st.nclu[0]))
                row = np.zeros(n)
                row[: st.nclu[0]] = st.nj[0, : st.nclu[0]]
                want_nj.append(row)
    u = tape.val[tape.kind == 1]
    assert len(u) == iters * n
    ref.set_tape(np.ones(len(u)), u)
    nout = (iters - burn) // thin
    nclu, njout = np.zeros(nout), np.zeros((n, nout), order="F")
    xf, spv, cu = np.ascontiguousarray(x.ravel()), np.array(sp), curr.astype(float)
    i, d = C.c_int, C.c_double
    rc = ref.L.ref_prior_ppmx_core(i(iters), i(burn), i(thin), i(n), i(1), i(P), i(0), d(1.0), d(sigma), _p(V), i(n), i(n),
                                   i(cohesion), i(CC), i(consim), i(similarity), i(calibration), i(1), _p(xf), _p(spv),
                                   _p(cu), _p(card), i(4), _p(nclu), _p(njout))
    assert rc == 0, ref.err()
    assert ref.tape_pos() == len(u)
    assert nclu.tolist() == want_K
    np.testing.assert_array_equal(njout, np.array(want_nj).T)


def test_reference_runs_on_the_free_running_generator(ref):
    """bench.py's timing leg: no tape, R's RNG replaced by the driver's own generator.  The run must
    be a valid chain (this is not a parity check)."""
    args, _, case = refrun.reference_args("default_ngg", 30, 10, 2, True)
    out = ref.dm_ppmx_ct(args, 4242)
    dims = case[1]
    n_a = np.bincount(case[3][0].treat, minlength=dims.T)
    assert ((out["nclus"] >= 1) & (out["nclus"] <= n_a[:, None])).all()
    np.testing.assert_allclose(out["pi"].sum(axis=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(out["pipred"].sum(axis=1), 1.0, rtol=1e-12)
    assert np.isfinite(out["WAIC"]) and np.isfinite(out["lpml"]).all()
    out2 = ref.dm_ppmx_ct(args, 4242)
    np.testing.assert_array_equal(out["cl_lab"], out2["cl_lab"])      # seeded: reproducible
