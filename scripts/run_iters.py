"""A few full iterations of the bench workload (for ncu launch lists of tppmx_batch_run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import argparse, bench
from treatppmx_b200.engine import Batch, Context

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 5
npred = int(sys.argv[2]) if len(sys.argv) > 2 else 28   # 10000 = BASELINE config 5's new patients
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # chain groups of run() (0 = automatic)
args = argparse.Namespace(datasets=256, chains_per_dataset=4, n=152, npred=npred, P=10, Q=2, T=3, cohesion=2, kcap=0, mixture=0)
w = bench.build_workload(args, 0)
ctx = Context(0)
b = Batch(ctx, w["model"], w["dims"], w["shared"], w["dsets"], chain_dataset=w["chain_ds"], chain_id=w["chain_id"])
b.init(1, w["labels"], w["nclu"], w["beta0"])
b.set_chain_groups(groups)
b.run(1, 0, 30)
b.reset_summaries()
b.run(1, 30, n_iter, burn=0, thin=1, psm=True, predictive=True)
print("ms per iteration", b.last_timing()[0] / n_iter, "launches", b.last_timing()[1], "chain groups", b.last_chain_groups())
