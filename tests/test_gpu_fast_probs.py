"""North-star level 1 for the PRODUCTION kernel: the per-step conditional allocation probabilities
of sweep_fast_kernel itself (closed-form Normal-Normal difference, table-driven log, branch-free
exp / lgamma, float-rounded shift, division-free draw -- treatppmx_b200/csrc/sweep_fast.cuh) against
the oracle's pweight (reference src/dm_ppmx_ct.cpp:791-802) on identical chain states, for EVERY
patient of a sweep: <= 1e-10 relative on every probability > 1e-290.

tppmx_batch_sweep_probe runs a normal sweep through the probe instantiation of the kernel, so the
chains stay in lock-step with the oracle before, during and after the probed sweep (labels / counts
bit-identical), which is what makes "identical chain state" hold at every step."""
import numpy as np
import pytest

import oracle_binding as ob
from cases import build_case
from treatppmx_b200 import datagen
from treatppmx_b200._abi import default_model

pytestmark = pytest.mark.gpu

RTOL = 1e-10        # BASELINE.json north_star: per-step probabilities within 1e-10 relative in fp64
PMIN = 1e-290


@pytest.fixture(scope="module")
def ctx():
    from treatppmx_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _compare(gp, gk, op, ok, arm_n, CC, what):
    nsteps, worst = 0, 0.0
    for t, nt in enumerate(arm_n):
        assert (gk[t, :nt] == ok[t, :nt]).all(), f"{what}: K per step differs in arm {t}"
        assert (gk[t, nt:] == -1).all()
        for it in range(int(nt)):
            k = int(ok[t, it]) + CC
            a, b = gp[t, it, :k], op[t, it, :k]
            m = b > PMIN
            rel = np.abs(a[m] - b[m]) / b[m]
            worst = max(worst, float(rel.max()))
            assert rel.max() <= RTOL, f"{what}: arm {t} step {it}: rel err {rel.max():.3e}"
            assert (a[~m] <= 1e-280).all()
            assert abs(a.sum() - 1.0) < 1e-12
            nsteps += 1
    return nsteps, worst


def _bench_workload(T=3, n=152, P=10, npred=28, nd=2):
    model = default_model(cohesion=2)
    dims_kw, dsets, ex = datagen.genmech_synthetic(n=n, npred=npred, P=P, Q=2, T=T, n_datasets=nd, seed=20240001)
    dims = datagen.make_dims(G=100, kcap=0, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=2)
    labels, nclu = datagen.initial_partition(dsets[0].treat, dims.T, dims.ntmax)
    return model, dims, shared, dsets, labels, nclu, [datagen.init_beta(d) for d in dsets]


CASES = {
    "bench_workload": _bench_workload,           # bench.py's shape: n=152, T=3, P=10, NGG, defaults
    "default_ngg": lambda: build_case("default_ngg", n_datasets=2),
    "calib2": lambda: build_case("calib2", n_datasets=2),
    "calib2_sqrt": lambda: build_case("calib2_sqrt", n_datasets=2),
    "dp_T3": lambda: build_case("dp_T3", n_datasets=2),
    "Q3_D3": lambda: build_case("Q3_D3", n_datasets=2),
    "ppm_only": lambda: build_case("ppm_only", n_datasets=2),   # PPMx = 0: cohesion x likelihood only
    "noreuse": lambda: build_case("noreuse", n_datasets=2),     # reuse = 0: auxiliary eta redrawn per patient
    "noreuse_T3": lambda: build_case("noreuse_T3", n_datasets=2),
    # calibration 1: normalised similarities, their own instantiation (16 | 32 lanes per unit)
    "calib1": lambda: build_case("calib1", n_datasets=2),
    "calib1_dp_T3": lambda: build_case("calib1_dp_T3", n_datasets=2),
    # Normal-Normal-Inverse-Gamma similarity (consim = 2): its own instantiation too
    "nnig": lambda: build_case("nnig", n_datasets=2),
    "nnig_ngg_T3": lambda: build_case("nnig_ngg_T3", n_datasets=2),
    # categorical covariates beside the continuous ones
    "categorical": lambda: build_case("categorical", n_datasets=2),
    "categorical_T3": lambda: build_case("categorical_T3", n_datasets=2),
    "cat_calib2": lambda: build_case("cat_calib2", n_datasets=2),
    # double-dipper Normal-Normal similarity
    "double_dipper": lambda: build_case("double_dipper", n_datasets=2),
    "dd_dp_T3": lambda: build_case("dd_dp_T3", n_datasets=2),
}


@pytest.mark.parametrize("name,lpu", [(n, l) for n in CASES for l in (8, 16, 32) if not (n.startswith(("calib1", "nnig", "cat", "double", "dd_")) and l == 8)
                                      and not (n in ("double_dipper", "dd_dp_T3") and l == 16)])   # their partitions outgrow 16 - CC staged clusters
def test_fast_kernel_step_probabilities(ctx, name, lpu):
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, betas = CASES[name]()
    nchains = 4
    cds = [c % 2 for c in range(nchains)]
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.set_sweep_impl("fast", lpu)
    b.init(11, labels, nclu, betas[0])
    pbs = [ob.Problem(model, dims, shared, d) for d in dsets]
    sts = [pbs[cds[c]].init_state(11, c, labels, nclu, betas[0]) for c in range(nchains)]
    CC = model.CC
    total = 0
    l = 0
    # the initial state, then states reached by full iterations (eta*, beta, JJ, sigma all moved)
    for burn in (0, 6, 3):
        if burn:
            b.run(11, l, burn)
            for c in range(nchains):
                pbs[cds[c]].iterate(sts[c], 11, c, l, burn)
            l += burn
        for chain in (0, 3):
            gp, gk = b.sweep_probe(11, l, chain)
            assert b.last_sweep_info() == ("fast", 0)
            ops = {}
            for c in range(nchains):  # every chain advanced by one sweep on the device
                if c == chain:
                    ops[c] = pbs[cds[c]].sweep_probe(sts[c], 11, c, l, gp.shape[2])
                else:
                    pbs[cds[c]].sweep(sts[c], 11, c, l, 1)
            arm_n = np.bincount(dsets[cds[chain]].treat, minlength=dims.T)
            n, worst = _compare(gp, gk, ops[chain][0], ops[chain][1], arm_n, CC, f"{name} lpu={lpu} l={l} chain={chain}")
            total += n
            for c in range(nchains):
                g = b.get_state(c)
                assert (g.labels == sts[c].labels).all() and (g.nclu == sts[c].nclu).all()
            # the rest of iteration l, so the next probe sees a properly updated state
            b.update(11, l)
            for c in range(nchains):
                pbs[cds[c]].update_iter(sts[c], 11, c, l)
            l += 1
    assert total >= 6 * dims.nobs
    b.close()


def _big_case(n, P, T, kcap, seed=20240401, mixture=8):
    model = default_model(cohesion=2)
    dims_kw, dsets, ex = datagen.genmech_synthetic(n=n, npred=0, P=P, Q=2, T=T, n_datasets=1, seed=seed, mixture=mixture)
    dims = datagen.make_dims(G=100, kcap=kcap, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=2, kmax=kcap)
    labels, nclu = datagen.initial_partition(dsets[0].treat, dims.T, dims.ntmax, nclu_init=5)
    return model, dims, shared, dsets, labels, nclu, datagen.init_beta(dsets[0])


def test_big_mode_step_probabilities(ctx):
    """BIG instantiation (arms of thousands of patients: labels / order / tables in global memory,
    one unit per warp), BASELINE config-4 shape class: n = 4500, P = 50, NGG."""
    from treatppmx_b200.engine import Batch
    model, dims, shared, dsets, labels, nclu, beta0 = _big_case(n=4500, P=50, T=3, kcap=96)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 0])
    b.init(5, labels, nclu, beta0)
    pb = ob.Problem(model, dims, shared, dsets[0])
    sts = [pb.init_state(5, c, labels, nclu, beta0) for c in range(2)]
    arm_n = np.bincount(dsets[0].treat, minlength=dims.T)
    l = 0
    for burn in (0, 2):
        if burn:
            b.run(5, l, burn)
            for c in range(2):
                pb.iterate(sts[c], 5, c, l, burn)
            l += burn
        gp, gk = b.sweep_probe(5, l, 1)
        assert b.last_sweep_info()[0] == "fast"
        pb.sweep(sts[0], 5, 0, l, 1)
        op, ok = pb.sweep_probe(sts[1], 5, 1, l, gp.shape[2])
        # steps the fast kernel ran (a unit that reaches the staged capacity hands over: K = -1 from there)
        for t, nt in enumerate(arm_n):
            ran = gk[t, :nt] >= 0
            assert ran.sum() > 0
            assert (gk[t, :nt][ran] == ok[t, :nt][ran]).all()
            for it in np.nonzero(ran)[0]:
                k = int(ok[t, it]) + model.CC
                a, o = gp[t, it, :k], op[t, it, :k]
                m = o > PMIN
                assert (np.abs(a[m] - o[m]) / o[m]).max() <= RTOL, (t, it)
        for c in range(2):
            g = b.get_state(c)
            assert (g.labels == sts[c].labels).all() and (g.nclu == sts[c].nclu).all()
        b.update(5, l)
        for c in range(2):
            pb.update_iter(sts[c], 5, c, l)
        l += 1
    b.close()
