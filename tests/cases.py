"""Shared definitions of the parity cases (used by the GPU tests and the golden generator)."""
from __future__ import annotations

import os

import numpy as np

from treatppmx_b200 import datagen
from treatppmx_b200._abi import default_model

# name -> (model overrides, data overrides)
CASES = {
    "default_ngg":   (dict(), dict(T=2)),
    "dp_T3":         (dict(cohesion=1, alpha=1.5), dict(T=3)),
    "double_dipper": (dict(similarity=2), dict(T=2, n=60, P=6)),
    "dd_dp_T3":      (dict(similarity=2, cohesion=1, alpha=1.5, calibration=2), dict(T=3)),
    "nnig":          (dict(consim=2, cohesion=1), dict(T=2, n=60, P=5)),
    "nnig_dd":       (dict(consim=2, similarity=2), dict(T=2, n=50, P=4)),
    "nnig_ngg_T3":   (dict(consim=2), dict(T=3)),
    "calib1":        (dict(calibration=1), dict(T=2, n=60, P=7)),
    "calib1_dp_T3":  (dict(calibration=1, cohesion=1, alpha=1.5), dict(T=3)),
    "calib2_sqrt":   (dict(calibration=2, coardegree=2), dict(T=2, n=60, P=7)),
    "calib2":        (dict(calibration=2, coardegree=1, cohesion=1), dict(T=3, n=66, P=4)),
    "noreuse":       (dict(reuse=0, CC=2), dict(T=2, n=50, P=5)),
    "noreuse_T3":    (dict(reuse=0, cohesion=1, alpha=2.0), dict(T=3)),
    "ppm_only":      (dict(PPMx=0, cohesion=1), dict(T=2, n=50, P=3)),
    "categorical":   (dict(cohesion=1), dict(T=2, n=60, P=4, ncat=2, maxC=3, npred=20)),
    "categorical_T3": (dict(cohesion=1, alpha=1.5), dict(T=3, ncat=3, maxC=4)),
    "cat_calib2":    (dict(cohesion=1, calibration=2), dict(T=2, n=60, P=4, ncat=2, maxC=3, npred=20)),
    "cat_dd_calib1": (dict(similarity=2, calibration=1), dict(T=2, n=48, P=3, ncat=3, maxC=4, npred=16)),
    "prior_only":    (dict(nolik=1, CC=1, cohesion=1), dict(T=1, n=40, P=4)),
    "Q3_D3":         (dict(), dict(T=2, n=40, P=3, Q=3)),
}


SIMUPATS_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simupats_config1.npz")
# BASELINE config 1: ppmxct() on the bundled simupats data through the genmech recipe, default
# switches; NGG cohesion with the in-repo V table, and DP cohesion (no third-party table involved)
SIMUPATS_CASES = {"simupats_ngg": dict(cohesion=2), "simupats_dp": dict(cohesion=1, alpha=1.0)}


def build_simupats_case(name: str):
    """(model, dims, shared, [dataset], labels, nclu, [beta0]) of config 1 from the committed fixture
    (tests/golden/make_simupats_fixture.py; the GPU box has no /root/reference)."""
    from treatppmx_b200 import simupats
    f = np.load(SIMUPATS_FIXTURE)
    sim = dict(cov=f["cov"], Y=f["Y"], treatment=f["treatment"])
    model, dims, shared, dsets, labels, nclu, beta0 = simupats.ppmxct_inputs(
        sim, init=(f["init_labels"], f["init_nclu"]), **SIMUPATS_CASES[name])
    return model, dims, shared, dsets, labels, nclu, [beta0]


def build_case(name: str, n_datasets: int = 1, seed: int = 20240001, kcap: int = 0):
    if name in SIMUPATS_CASES:
        return build_simupats_case(name)
    mkw, dkw = CASES[name]
    dkw = dict(dkw)
    dkw.setdefault("n", 152)
    dkw.setdefault("npred", 28)
    model = default_model(**mkw)
    dims_kw, dsets, ex = datagen.genmech_synthetic(n_datasets=n_datasets, seed=seed, **dkw)
    dims = datagen.make_dims(G=100, kcap=kcap, **dims_kw)
    shared = datagen.make_shared(dims, ex["n_a"], cohesion=model.cohesion)
    labels, nclu = datagen.initial_partition(dsets[0].treat, dims.T, dims.ntmax)
    betas = [datagen.init_beta(d) for d in dsets]
    return model, dims, shared, dsets, labels, nclu, betas


STATE_FIELDS_EXACT = ("labels", "nj", "nclu", "njc")
STATE_FIELDS_CLOSE = ("sumx", "sumx2", "eta_star", "eta_empty", "JJ", "ss", "loggamma", "beta",
                      "sigma", "kappa", "Theta", "Sigma")


def occupied_view(st, dims):
    """Mask a host state down to what is defined: labels of real patients, occupied clusters."""
    out = {}
    T = dims.T
    out["nclu"] = st.nclu.copy()
    for t in range(T):
        K = int(st.nclu[t])
        out[f"nj{t}"] = st.nj[t, :K].copy()
        out[f"eta{t}"] = st.eta_star[:, :K, t].copy()
        P = max(dims.P, 1)
        out[f"sumx{t}"] = st.sumx[t, : K * P].copy()
        out[f"sumx2{t}"] = st.sumx2[t, : K * P].copy()
        ncc = max(dims.ncat * dims.maxC, 1)
        out[f"njc{t}"] = st.njc[t, : K * ncc].copy()
    return out
