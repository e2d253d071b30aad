"""Device-timed rate of the sweep kernel alone on the bench workload (or a config-4 style shape): a
lean loop for kernel A/B runs -- select the library build with TPPMX_LIB=... .

    python scripts/sweep_rate.py [--reps 8] [--sweeps 20] [--lpu 0] [--n 152 --P 10 --T 3 --datasets 256
                                  --chains-per-dataset 4 --kcap 0 --mixture 0 --v 0.5] [--impl auto]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import bench
from treatppmx_b200.engine import Batch, Context

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--sweeps", type=int, default=20)
ap.add_argument("--burn", type=int, default=50)
ap.add_argument("--full-burn", type=int, default=30, help="full iterations before the sweeps (realistic partition)")
ap.add_argument("--lpu", type=int, default=0)
ap.add_argument("--impl", default="auto")
ap.add_argument("--staged", type=int, default=0)
ap.add_argument("--big-threads", type=int, default=0)
ap.add_argument("--datasets", type=int, default=256)
ap.add_argument("--chains-per-dataset", type=int, default=4)
ap.add_argument("--n", type=int, default=152)
ap.add_argument("--npred", type=int, default=28)
ap.add_argument("--P", type=int, default=10)
ap.add_argument("--Q", type=int, default=2)
ap.add_argument("--T", type=int, default=3)
ap.add_argument("--cohesion", type=int, default=2)
ap.add_argument("--kcap", type=int, default=0)
ap.add_argument("--mixture", type=int, default=0)
ap.add_argument("--v", type=float, default=0.5, help="similparam v (within-cluster variance): smaller => more clusters")
a = ap.parse_args()

w = bench.build_workload(a, 0)
w["model"].similparam[2] = a.v
ctx = Context(0)
b = Batch(ctx, w["model"], w["dims"], w["shared"], w["dsets"], chain_dataset=w["chain_ds"], chain_id=w["chain_id"])
b.set_sweep_impl(a.impl, a.lpu, a.staged, a.big_threads)
b.init(1, w["labels"], w["nclu"], w["beta0"])
it = 0
if a.full_burn:
    b.run(1, 0, a.full_burn)
    it = a.full_burn
b.sweep(1, it, a.burn)
it += a.burn
ms, hand = [], 0
for r in range(a.reps):
    b.sweep(1, it, a.sweeps)
    it += a.sweeps
    ms.append(b.last_timing()[0])
    hand += b.last_sweep_info()[1]
lab, ncl = b.get_partitions()
nre = b.n_chains * w["dims"].nobs * a.sweeps
best, med = min(ms), float(np.median(ms))
print(f"lib={os.path.basename(os.environ.get('TPPMX_LIB', 'libtppmx.so'))} impl={b.last_sweep_info()[0]} lpu={a.lpu} "
      f"chains={b.n_chains} n={a.n} P={a.P} Kbar={ncl.mean():.2f} Kmax={ncl.max()} handovers={hand} "
      f"ms/launch median={med:.4f} best={best:.4f}  realloc/s median={nre / med * 1e3:.4e} best={nre / best * 1e3:.4e}")
