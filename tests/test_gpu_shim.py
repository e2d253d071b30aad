"""The Rcpp shim (treatppmx_b200/csrc/rcpp_shim/treatppmx_shim.cpp) -- the package's two sampler
entry points with the REFERENCE's own C++ signatures (src/dm_ppmx_ct.cpp:10-23,
src/prior_ppmx.cpp:9-16) over the CUDA library -- compiled against the stand-in Rcpp / Armadillo
headers of oracle/refshim (no R in this image; `make -C oracle shim` -> oracle/_ref/libshim.so) and
driven through the same C driver that drives the reference's sources.

dm_ppmx_ct through the shim must reproduce the output of the REFERENCE's own code
(tests/golden/ref_*.npz) -- labels, counts, draws, grid picks identical, reals <= 1e-9 -- from the
reference's 42 arguments incl. the dense Vwm cube, with the Philox key drawn from (the stand-in of)
R's RNG as a real build draws it under set.seed()."""
import importlib.util
import os

import numpy as np
import pytest

import refrun
from test_reference_golden import _check

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(HERE, "golden", "make_reference_golden.py"))
mrg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mrg)


@pytest.fixture(scope="module")
def shim():
    return refrun.RefLib(refrun.SHIM_SO)


@pytest.mark.parametrize("name", sorted(mrg.REF_GOLDEN))
def test_shim_dm_ppmx_ct_reproduces_reference_output(shim, name):
    iters, burn, thin, seed = mrg.REF_GOLDEN[name]
    want = dict(np.load(mrg.path(name)))
    args, _, _ = refrun.reference_args(name, iters, burn, thin, True)       # dense Vwm cube, as ppmxct() passes it
    got = shim.dm_ppmx_ct(args, refrun.SeedTape(seed))
    assert got.pop("tape_used") == 2                                       # the two unif_rand() draws of the key
    shim.L.tppmx_shim_last_seed.restype = __import__("ctypes").c_uint64
    assert shim.L.tppmx_shim_last_seed() == seed
    _check(got, want, rtol=1e-9)


def test_shim_seed_comes_from_the_r_rng(shim):
    """set.seed() governs a run: different states of R's RNG give different chains, the same state the same."""
    iters, burn, thin, _ = mrg.REF_GOLDEN["dp_T3"]
    args, _, _ = refrun.reference_args("dp_T3", iters, burn, thin, True)
    a = shim.dm_ppmx_ct(args, 5)       # int: the driver's free-running generator seeded with 5
    b = shim.dm_ppmx_ct(args, 5)
    c = shim.dm_ppmx_ct(args, 6)
    assert (a["cl_lab"] == b["cl_lab"]).all() and (a["pipred"] == b["pipred"]).all()
    assert not (a["cl_lab"] == c["cl_lab"]).all()


def test_shim_reports_library_errors_as_exceptions(shim):
    iters, burn, thin, _ = mrg.REF_GOLDEN["dp_T3"]
    args, _, _ = refrun.reference_args("dp_T3", iters, burn, thin, True)
    args["CC"] = 99                                                        # > TPPMX_MAXCC
    with pytest.raises(RuntimeError, match="CC out of range"):
        shim.dm_ppmx_ct(args, 1)


@pytest.mark.parametrize("cohesion", [1, 2])
def test_shim_prior_ppmx_core_matches_direct_call(shim, cohesion):
    from treatppmx_b200 import datagen
    from treatppmx_b200.engine import Context
    from treatppmx_b200.prior import prior_ppmx_core
    n, P, iters, burn, thin, seed = 30, 3, 40, 10, 2, 23
    x = np.random.default_rng(4).standard_normal((n, P))
    sigma, kappa = 0.2, 1.5
    V = np.zeros((n + 1, n + 1), order="F")
    for r in (n - 2, n - 1):
        V[r, : r + 1] = np.exp(datagen.ngg_logV(r + 1, r + 1, sigma, kappa))
    curr = (np.arange(n) % 4) + 1
    sp = (0.0, 10.0, 0.5, 1.0, 2.0, 0.0, 0.1)
    a = (iters, burn, thin, n, 1, P, 0, 1.0, sigma, V, cohesion, 1, 1, 1, 0, 1, x.ravel(), sp, curr.astype(float),
         np.concatenate([np.bincount(curr)[1:], np.zeros(n - 4)]).astype(float), 4)
    got = shim.prior_ppmx_core(*a, refrun.SeedTape(seed))
    ctx = Context(0)
    want = prior_ppmx_core(ctx, *a, seed=seed)
    ctx.close()
    assert got["nclu"].tolist() == want["nclu"].tolist()
    np.testing.assert_array_equal(got["njout"], want["njout"])
    assert got["nclu"].min() >= 1 and (got["njout"].sum(0) == n).all()
