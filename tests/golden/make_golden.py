"""Generates the committed golden vectors of tests/golden/*.npz from the CPU oracle.

    python tests/golden/make_golden.py

The reference itself (R / Rcpp) cannot be run in this image and ships no fixtures, so the
vectors are ORACLE outputs: they pin the oracle against regressions (tests/test_golden.py, CPU)
and let the GPU path be checked without building the oracle (tests/test_golden.py, -m gpu).
Inputs are regenerated from treatppmx_b200.datagen with fixed seeds; only outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import oracle_binding as ob  # noqa: E402
from cases import build_case  # noqa: E402

GOLDEN_CASES = ["default_ngg", "dp_T3", "nnig_dd", "cat_dd_calib1", "calib2_sqrt"]
SEED, SWEEPS, ITERS = 4242, 15, 12


def make(name):
    model, dims, shared, dsets, labels, nclu, betas = build_case(name)
    pb = ob.Problem(model, dims, shared, dsets[0])
    st = pb.init_state(SEED, 0, labels, nclu, betas[0])
    out = {"init_ss": st.ss.copy(), "init_eta_star": st.eta_star.copy(), "init_loggamma": st.loggamma.copy()}
    pb.sweep(st, SEED, 0, 0, SWEEPS)
    out.update(sweep_labels=st.labels.copy(), sweep_nclu=st.nclu.copy(), sweep_nj=st.nj.copy(),
               sweep_sumx=st.sumx.copy(), sweep_sumx2=st.sumx2.copy())
    n_a = np.bincount(dsets[0].treat, minlength=dims.T)
    picks = [(t, it) for t in range(dims.T) for it in (0, n_a[t] // 3, n_a[t] - 1)]
    W = dims.ntmax + model.CC
    logw, prob, Ks = np.full((len(picks), W), np.nan), np.full((len(picks), W), np.nan), []
    for r, (t, it) in enumerate(picks):
        lw, pr, K = pb.score_patient(st, t, it, seed=SEED, chain_id=0, iter_=SWEEPS)
        logw[r, : len(lw)], prob[r, : len(pr)] = lw, pr
        Ks.append(K)
    out.update(score_picks=np.array(picks), score_logw=logw, score_prob=prob, score_K=np.array(Ks))
    pi, yp, pc = pb.predict(st, SEED, 0, SWEEPS)
    out.update(pred_pi=pi, pred_y=yp, pred_class=pc)
    if not (model.cohesion == 2 and dims.ncat > 0):
        pb.iterate(st, SEED, 0, SWEEPS, ITERS)
        out.update(iter_labels=st.labels.copy(), iter_nclu=st.nclu.copy(), iter_beta=st.beta.copy(),
                   iter_sigma=st.sigma.copy(), iter_idxs=st.idxs.copy(), iter_tau=st.tau.copy(),
                   iter_Theta=st.Theta.copy(), iter_JJ=st.JJ.copy(), iter_ss=st.ss.copy())
    return out


if __name__ == "__main__":
    for name in GOLDEN_CASES:
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **make(name))
        print("wrote", name)
