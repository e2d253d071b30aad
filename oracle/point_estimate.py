"""CPU restatement (numpy) of the point-estimate partition of the co-clustering matrix.
TEST INFRASTRUCTURE ONLY: imported by tests/, never by treatppmx_b200/.

The reference imports mcclust.ext::minVI / minbinder.ext (R/treatppmx.R:30-31) but does not vendor
the package (GitHub sarawade/mcclust.ext, no version pinned in DESCRIPTION), so the PUBLISHED
algorithm is restated: Wade & Ghahramani (2018), "Bayesian cluster analysis: point estimation and
credible balls", Bayesian Analysis 13(2) section 4; the package's method "avg" / "comp":
hclust(as.dist(1 - psm), "average" | "complete"), cutree for k = 1..max.k (default ceiling(n / 8)),
the cut with the smallest loss (which.min: the smallest k among equals).

  VI lower bound (VI.lb):  (1/n) sum_i [ log2 #{j: c_j = c_i} + log2 sum_j psm_ij - 2 log2 sum_j 1{c_j = c_i} psm_ij ]
  Binder loss:             sum_{i<j} | 1{c_i = c_j} - psm_ij |

Parity pin: mcclust.ext is absent and R is absent, so this restatement is pinned (tests/test_point_estimate.py)
against scipy.cluster.hierarchy (linkage + fcluster, the same Lance-Williams recurrences) on tie-free
matrices and against brute force over all partitions of small sets; ties between equal
dissimilarities go to the smallest (i, j), i < j -- parity with R's hclust on exact ties is unpinned."""
import numpy as np


def vi_lb(cl: np.ndarray, psm: np.ndarray) -> float:
    n = len(cl)
    f = 0.0
    for i in range(n):
        ind = cl == cl[i]
        f += (np.log2(ind.sum()) + np.log2(psm[i].sum()) - 2.0 * np.log2((ind * psm[i]).sum())) / n
    return float(f)


def binder(cl: np.ndarray, psm: np.ndarray) -> float:
    same = (cl[:, None] == cl[None, :]).astype(float)
    return float(np.triu(np.abs(same - psm), 1).sum())


def first_appearance(cl: np.ndarray) -> np.ndarray:
    out, seen = np.zeros(len(cl), dtype=np.int32), {}
    for i, c in enumerate(cl):
        out[i] = seen.setdefault(int(c), len(seen) + 1)
    return out


def cuts(psm: np.ndarray, linkage: str = "average"):
    """Yields (k, cluster ids) for k = n, n-1, ..., 1 of the agglomeration of 1 - psm."""
    n = psm.shape[0]
    dist = 1.0 - psm.astype(float)
    cl, size = np.arange(n), np.ones(n, dtype=int)
    yield n, cl.copy()
    for k in range(n, 1, -1):
        act = [i for i in range(n) if cl[i] == i]
        best, bi, bj = np.inf, -1, -1
        for a, i in enumerate(act):
            for j in act[a + 1:]:
                if dist[i, j] < best:        # strict: the first (smallest (i, j)) minimum wins
                    best, bi, bj = dist[i, j], i, j
        ni, nj = size[bi], size[bj]
        for c in act:
            if c in (bi, bj):
                continue
            di, dj = dist[min(c, bi), max(c, bi)], dist[min(c, bj), max(c, bj)]
            dn = max(di, dj) if linkage == "complete" else (ni * di + nj * dj) / (ni + nj)
            dist[min(c, bi), max(c, bi)] = dn
        cl[cl == bj] = bi
        size[bi] = ni + nj
        yield k - 1, cl.copy()


def point_estimate(psm: np.ndarray, loss: str = "VI", linkage: str = "average", max_k: int = 0):
    """(labels 1-based in order of first appearance, k, loss value)"""
    n = psm.shape[0]
    max_k = min(max_k, n) if max_k > 0 else (n + 7) // 8
    f = vi_lb if loss == "VI" else binder
    best = (np.inf, 0, None)
    for k, cl in cuts(psm, linkage):
        if k <= max_k:
            v = f(cl, psm)
            if v <= best[0]:             # going down in k: the smaller k wins ties
                best = (v, k, cl)
    return first_appearance(best[2]), best[1], best[0]
