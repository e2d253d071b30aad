"""BASELINE config 1: the package's bundled fixture `simupats` (reference data/simupats.rda, a
152 x 92 matrix of gene-expression-like doubles) and the `genmech` recipe that turns it into a
treatment-selection data set (reference R/gendata.R:134-254), plus the marshalling ppmxct() does
before its .Call (R/dm_ppmx_ct.R:125-277).  Host-side numpy only: inputs, not the compute path.

The .rda file is a bzip2-compressed RDX2/XDR stream holding one REALSXP with a `dim` attribute;
it is read here without R.  Two steps of the recipe use R facilities that cannot be reproduced
bit for bit and are replaced by documented stand-ins:
  * the outcome draws (`rmultinom` on R's RNG, R/gendata.R:207-211)  -> numpy Generator(seed);
  * the initial partition (`stats::kmeans(nstart = 25)`, Hartigan-Wong, R/dm_ppmx_ct.R:214-225)
    -> Lloyd's k-means with 25 restarts on the same rows.
Everything else (scale(), prcomp(), the ordinal-logit probabilities, the arm assignment, the
Spearman/FDR beta start, the (sigma,kappa) grid) follows the R code.
"""
from __future__ import annotations

import bz2
import struct

import numpy as np

from . import datagen
from ._abi import Dims, HostDataset, default_model

NROW, NCOL = 152, 92


def read_rda_matrix(path: str) -> np.ndarray:
    """The 152 x 92 matrix of data/simupats.rda.  RDX2 layout: 'RDX2\\nX\\n', version triple, a
    pairlist node tagged 'simupats' whose value is a REALSXP (type 14 with the attribute bit, flags
    0x0000020e), its length as a big-endian int32, the doubles in column-major XDR (big-endian)
    order, then the `dim` attribute."""
    d = bz2.decompress(open(path, "rb").read())
    if not d.startswith(b"RDX2\nX\n"):
        raise ValueError("not an RDX2 XDR stream")
    n = NROW * NCOL
    at = d.find(struct.pack(">ii", 0x0000020E, n))
    if at < 0:
        raise ValueError("REALSXP of length 152*92 not found")
    a = np.frombuffer(d, dtype=">f8", count=n, offset=at + 8).astype(np.float64)
    tail = d[at + 8 + n * 8:]
    k = tail.find(b"dim")
    dims = struct.unpack(">ii", tail[k + 3 + 8: k + 3 + 16]) if k >= 0 else (NROW, NCOL)
    if tuple(dims) != (NROW, NCOL):
        raise ValueError(f"unexpected dim attribute {dims}")
    return a.reshape(NCOL, NROW).T.copy()


def _scale(a: np.ndarray) -> np.ndarray:
    """R's scale(): centre, divide by the sample standard deviation (n - 1)."""
    return (a - a.mean(0)) / a.std(0, ddof=1)


def genoutcome_probs(alpha, beta1, beta2, beta3, metx, x2, x3) -> np.ndarray:
    """R/gendata.R:30-47 without the outcome draw: the three category probabilities, rounded to four
    decimals as `round(cbind(myy, prob), 4)` does."""
    eta1 = alpha[0] + beta1[0] * metx ** 3 + beta2[0] * x2 + beta3[0] * x3
    eta2 = alpha[1] + beta1[1] * metx ** 3 + beta2[1] * x2 + beta3[1] * x3
    p1 = 1 / ((np.exp(eta1) + 1) * (np.exp(eta2) + 1))
    p2 = np.exp(eta1) / ((np.exp(eta1) + 1) * (np.exp(eta2) + 1))
    p3 = np.exp(eta2) / (np.exp(eta2) + 1)
    return np.round(np.stack([p1, p2, p3], 1), 4)


def genmech(mat: np.ndarray, npred: int = 10, progscen: int = 1, nset: int = 1, seed: int = 121) -> dict:
    """R/gendata.R:134-254 (`genmech`, predscen = 1): returns Y [nobs, 3, nset] one-hot outcomes,
    treatment [nobs] in {1, 2}, cov = [X (npred columns) | Z (2 columns)], prob = the two arms'
    response probabilities."""
    genenorm = _scale(np.asarray(mat, dtype=np.float64))
    Xp = genenorm[:, :npred]
    Xc = Xp - Xp.mean(0)                                      # prcomp(): centred, not rescaled
    _, _, vt = np.linalg.svd(Xc, full_matrices=False)
    pcs = Xc @ vt.T
    sc = lambda v: (v - v.mean()) / v.std(ddof=1)
    metx1 = sc(pcs[:, 0]) + sc(pcs[:, 1]) + 0.50              # :147  (sign of a PC is arbitrary in prcomp too)
    metx = np.sign(metx1) * np.abs(metx1) ** 0.2 * 0.45      # :148
    nobs = mat.shape[0]
    if progscen == 0:
        z2, z3 = np.zeros(nobs), np.zeros(nobs)
    else:
        z2, z3 = genenorm[:, 90].copy(), genenorm[:, 91].copy()
        if progscen == 2:
            z2 = np.sign(z2) * np.abs(z2) ** 0.5
            z3 = np.sign(z3) * np.abs(z3) ** 0.2
    zero = (0.0, 0.0)
    prob1 = genoutcome_probs((-0.5, -1.0), (2.0, 2.6), zero, zero, metx, z2, z3)      # :171-185
    prob2 = genoutcome_probs((0.7, -1.0), (-1.0, -3.0), zero, zero, metx, z2, z3)
    probprog = genoutcome_probs((1.0, -0.5), zero, (1.0, 0.5), (0.7, 1.0), metx, z2, z3)
    myprob = []
    for pr in (prob1, prob2):                                                          # :190-193
        w = pr * probprog
        myprob.append(w / w.sum(1, keepdims=True))
    trtsgn = np.tile([1, 2], nobs // 2)                                                # :196
    rng = np.random.default_rng(seed)
    Y = np.zeros((nobs, 3, nset))
    for s in range(nset):                                                              # :198-216
        for k in range(nobs):
            Y[k, rng.choice(3, p=myprob[trtsgn[k] - 1][k]), s] = 1.0
    return dict(Y=Y, treatment=trtsgn, cov=np.hstack([Xp, np.stack([z2, z3], 1)]), prob=myprob)


def kmeans_labels(X: np.ndarray, k: int, rng: np.random.Generator, nstart: int = 25, iters: int = 100) -> np.ndarray:
    """Lloyd's k-means, best of `nstart` random starts (stand-in for stats::kmeans): labels 1..k."""
    best, best_ss = None, np.inf
    n = X.shape[0]
    for _ in range(nstart):
        cen = X[rng.choice(n, size=k, replace=False)].copy()
        lab = np.zeros(n, dtype=np.int64)
        for _ in range(iters):
            d2 = ((X[:, None, :] - cen[None, :, :]) ** 2).sum(2)
            new = d2.argmin(1)
            for j in range(k):                      # keep every cluster occupied
                if not (new == j).any():
                    new[d2[:, j].argmin()] = j
            if (new == lab).all():
                break
            lab = new
            for j in range(k):
                cen[j] = X[lab == j].mean(0)
        ss = sum(((X[lab == j] - X[lab == j].mean(0)) ** 2).sum() for j in range(k))
        if ss < best_ss:
            best, best_ss = lab.copy(), ss
    return best.astype(np.int32) + 1


def ppmxct_inputs(sim: dict, replicate: int = 0, n_train: int = 124, cohesion: int = 2, nclu_init: int = 5,
                  seed: int = 121, init=None, **model_kw):
    """What ppmxct() hands to its .Call for the genmech data set (R/dm_ppmx_ct.R:125-277), train /
    new-patient split of R/gendata.R:443-444 (first 124 rows train, last 28 new): returns
    (model, dims, shared, [dataset], init_labels, init_nclu, initbeta) in this package's host types.
    `init` = (labels, nclu) skips the k-means start (the committed fixture carries it).
    Defaults of ppmxct(): PPMx = 1, cohesion = 2, similarity = 1, consim = 1, calibration = 0, CC = 3,
    reuse = 1, nclu_init = 5, hsp = 1, update_hierarchy = 1 (R/dm_ppmx_ct.R:89-93)."""
    Y = sim["Y"][:, :, replicate] if np.ndim(sim["Y"]) == 3 else sim["Y"]
    cov, trt = np.asarray(sim["cov"]), np.asarray(sim["treatment"]).astype(np.int64) - 1   # arms 0-based, :227
    npc = cov.shape[1] - 2
    X, Z = cov[:, :npc], cov[:, npc:]
    tr, te = np.arange(n_train), np.arange(n_train, cov.shape[0])
    T = int(trt.max()) + 1
    n_a = np.bincount(trt[tr], minlength=T)
    ntmax = int(n_a.max())
    model = default_model(cohesion=cohesion, **model_kw)
    dims = datagen.make_dims(G=100, kcap=0, nobs=len(tr), T=T, D=Y.shape[1], Q=Z.shape[1], P=npc, ncat=0, maxC=0,
                             npred=len(te), ntmax=ntmax)
    shared = datagen.make_shared(dims, n_a, cohesion=cohesion)
    ds = HostDataset(treat=trt[tr], y=Y[tr], z=Z[tr], xcon=X[tr], xcat=np.zeros((len(tr), 1), np.int32),
                     zpred=Z[te], xconp=X[te], xcatp=np.zeros((len(te), 1), np.int32))
    if init is not None:
        return model, dims, shared, [ds], np.asfortranarray(init[0], dtype=np.int32), np.asarray(init[1], np.int32), datagen.init_beta(ds)
    rng = np.random.default_rng(seed)
    labels = np.zeros((T, ntmax), dtype=np.int32, order="F")
    nclu = np.zeros(T, dtype=np.int32)
    for t in range(T):                                                                # :214-225
        rows = np.nonzero(trt[tr] == t)[0]
        labels[t, : len(rows)] = kmeans_labels(X[tr][rows], nclu_init, rng)
        nclu[t] = nclu_init
    return model, dims, shared, [ds], labels, nclu, datagen.init_beta(ds)
