"""tests/golden/ref_<case>.npz hold the return list of the REFERENCE's own dm_ppmx_ct (its sources
compiled unmodified, tests/golden/make_reference_golden.py) on the Philox-contract variates.  The
oracle (CPU) and the CUDA drop-in tppmx_dm_ppmx_ct (GPU) must reproduce them; neither test needs
the reference or oracle/_ref at run time."""
import importlib.util
import os

import numpy as np
import pytest

import refrun

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(HERE, "golden", "make_reference_golden.py"))
mrg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mrg)

EXACT = ("cl_lab", "nclus", "clupred", "yispred", "ypred", "sigma_ngg", "kappa_ngg", "num_treat", "beta_acc")
CLOSE = ("eta", "beta", "pi", "pipred", "lpml", "eta_acc", "theta_prior", "sigma_prior")


def _check(got, want, rtol):
    for k in EXACT:
        np.testing.assert_array_equal(np.asarray(got[k], dtype=float).reshape(want[k].shape), want[k], err_msg=k)
    for k in CLOSE:
        np.testing.assert_allclose(np.asarray(got[k]).reshape(want[k].shape), want[k], rtol=rtol, atol=1e-12, err_msg=k)
    assert float(got["WAIC"]) == pytest.approx(float(want["WAIC"]), rel=rtol)


@pytest.mark.parametrize("name", sorted(mrg.REF_GOLDEN))
def test_oracle_reproduces_reference_output(name):
    iters, burn, thin, seed = mrg.REF_GOLDEN[name]
    want = dict(np.load(mrg.path(name)))
    _, _, case = refrun.reference_args(name, iters, burn, thin, False)
    _check(refrun.oracle_run(case, iters, burn, thin, seed), want, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mrg.REF_GOLDEN))
def test_device_drop_in_reproduces_reference_output(name):
    """tppmx_dm_ppmx_ct (CUDA, through the C ABI) against the reference's own output: labels, counts,
    draws and grid picks identical, real-valued outputs <= 1e-9 relative."""
    from treatppmx_b200.engine import Context
    from treatppmx_b200.ppmxct import dm_ppmx_ct
    iters, burn, thin, seed = mrg.REF_GOLDEN[name]
    want = dict(np.load(mrg.path(name)))
    args, logV, _ = refrun.reference_args(name, iters, burn, thin, False)
    ctx = Context(0)
    got = dm_ppmx_ct(ctx, **args, seed=seed, logV_compact=logV)
    ctx.close()
    _check(got, want, rtol=1e-9)
