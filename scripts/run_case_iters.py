"""Full iterations (sweep + updates + predictive + co-clustering every iteration) on a parity case's model switches,
1024 chains over 256 replicate datasets:  python scripts/run_case_iters.py CASE [CASE ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from cases import build_case
from treatppmx_b200.engine import Batch, Context

ctx = Context(0)
for name in sys.argv[1:] or ["default_ngg", "calib1_dp_T3", "nnig_ngg_T3", "categorical_T3", "dd_dp_T3"]:
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=256)
    cds = np.repeat(np.arange(256), 4)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.init(1, labels, nclu, betas[0])
    b.run(1, 0, 20)
    b.reset_summaries()
    b.run(1, 20, 20, burn=0, thin=1, psm=True, predictive=True)
    ms_saved = b.last_timing()[0] / 20
    b.run(1, 40, 20)
    ms_plain = b.last_timing()[0] / 20
    b.predict(1, 60, want=())
    print(f"{name}: n={dims.nobs} T={dims.T} P={dims.P} ncat={dims.ncat}: {ms_saved:.3f} ms per iteration with predictive + co-clustering, "
          f"{ms_plain:.3f} ms without; predictive step alone {b.last_timing()[0] * 1e3:.0f} us; sweep kernel {b.last_sweep_info()[0]}")
    b.close()
