"""Generates tests/golden/simupats_config1.npz: BASELINE config 1's inputs -- the reference's bundled
fixture data/simupats.rda read without R and pushed through the genmech recipe
(treatppmx_b200/simupats.py, reference R/gendata.R:134-254), one replicate, plus the k-means start
ppmxct() would compute (R/dm_ppmx_ct.R:214-225).  Needs /root/reference; run from the repo root:

    python tests/golden/make_simupats_fixture.py

The GPU box has no /root/reference: tests and bench.py read this file instead."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

from treatppmx_b200 import simupats  # noqa: E402

RDA = "/root/reference/data/simupats.rda"
OUT = os.path.join(HERE, "simupats_config1.npz")


def main():
    mat = simupats.read_rda_matrix(RDA)
    assert mat.shape == (152, 92) and abs(mat.min() - 0.1) < 1e-12 and abs(mat.mean() - 1.129696495523364) < 1e-12
    sim = simupats.genmech(mat, npred=10, progscen=1, nset=1, seed=121)
    _, dims, _, _, labels, nclu, _ = simupats.ppmxct_inputs(sim, cohesion=1, seed=121)
    np.savez_compressed(OUT, cov=sim["cov"], Y=sim["Y"][:, :, 0], treatment=sim["treatment"], prob1=sim["prob"][0],
                        prob2=sim["prob"][1], init_labels=labels, init_nclu=nclu,
                        raw_checksum=np.array([mat.sum(), (mat ** 2).sum()]))
    print(OUT, os.path.getsize(OUT), "bytes; n_t =", dims.ntmax, "outcome counts", sim["Y"][:, :, 0].sum(0))


if __name__ == "__main__":
    main()
