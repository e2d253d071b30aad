"""A small end-to-end exercise of every kernel (for compute-sanitizer runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from cases import build_case
from treatppmx_b200.engine import Batch, Context

ctx = Context(0)
for name, impl, lpu, staged in (("default_ngg", "fast", 16, 0), ("dp_T3", "fast", 8, 4), ("calib2", "fast", 32, 0),
                                ("categorical", "generic", 0, 0), ("nnig_dd", "generic", 0, 0)):
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=2)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=[0, 1, 0])
    b.set_sweep_impl(impl, lpu, staged_clusters=staged)
    b.init(3, labels, nclu, betas[0])
    b.sweep(3, 0, 3)
    if not (model.cohesion == 2 and dims.ncat > 0):
        b.reset_summaries()
        b.run(3, 3, 4, burn=1, thin=1, psm=True, predictive=True)
        b.summaries([0.0, 40.0, 100.0])
    b.predict(3, 9)
    b.score_patient(0, 0, 1)
    b.get_partitions()
    b.close()
    print("ok", name, impl)
ctx.close()
