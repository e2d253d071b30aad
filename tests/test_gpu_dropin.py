"""The drop-in entry point tppmx_dm_ppmx_ct (the reference's dm_ppmx_ct: 42 arguments in, the
18-element list out) against the same run composed from the CPU oracle's pieces: sweep
(:347-860), updates (:866-1055), in-sample fit (:1060-1075), predictive (:1080-1385), save
(:1391-1429), WAIC (:1445-1450)."""
import numpy as np
import pytest

import oracle_binding as ob
from refrun import oracle_run as _oracle_run, reference_args as _reference_args

pytestmark = pytest.mark.gpu


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


def test_dm_ppmx_ct_is_interruptible():
    from treatppmx_b200 import _lib
    from treatppmx_b200.engine import Context
    from treatppmx_b200.ppmxct import dm_ppmx_ct
    args, logV, _ = _reference_args("default_ngg", 40, 10, 2, False)
    ctx = Context(0)
    flag = np.ones(1, dtype=np.int32)
    ctx.set_interrupt_flag(flag, check_every=3)
    with pytest.raises(_lib.TppmxError) as ei:
        dm_ppmx_ct(ctx, **args, seed=1, logV_compact=logV)
    assert ei.value.code == 5 and "iteration 2" in str(ei.value)
    flag[0] = 0
    out = dm_ppmx_ct(ctx, **args, seed=1, logV_compact=logV)     # the context stays usable
    assert out["nclus"].shape[-1] == 15
    ctx.close()
