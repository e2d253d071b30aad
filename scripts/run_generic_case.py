"""Sweep rate on a parity case's model switches (e.g. nnig, categorical, calib1), 1024 chains over 256
replicate datasets: the kernel the library picks, and the generic kernel beside it where that is the fast one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from cases import build_case
from treatppmx_b200.engine import Batch, Context

ctx = Context(0)
for name in sys.argv[1:] or ["nnig", "categorical", "calib1", "double_dipper"]:
    model, dims, shared, dsets, labels, nclu, betas = build_case(name, n_datasets=256)
    cds = np.repeat(np.arange(256), 4)
    b = Batch(ctx, model, dims, shared, dsets, chain_dataset=cds)
    b.init(1, labels, nclu, betas[0])
    b.sweep(1, 0, 10)
    b.sweep(1, 10, 10)   # the automatic lanes per unit may change after the first launches (and a kernel's first launch loads it)
    b.sweep(1, 20, 10)
    ms = b.last_timing()[0]
    line = f"{name}: n={dims.nobs} T={dims.T} P={dims.P} ncat={dims.ncat}: {len(cds) * dims.nobs * 10 / (ms * 1e-3):.3e} reallocations/s ({b.last_sweep_info()[0]} kernel)"
    if b.last_sweep_info()[0] == "fast":
        b.set_sweep_impl("generic")
        b.sweep(1, 30, 10)
        b.sweep(1, 40, 10)
        line += f"; generic kernel {len(cds) * dims.nobs * 10 / (b.last_timing()[0] * 1e-3):.3e}"
    print(line)
    b.close()
