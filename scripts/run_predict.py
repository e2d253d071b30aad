"""One predictive step of the config-5 shape (10 000 new patients) for ncu captures of predict_fast_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import argparse, bench
from treatppmx_b200.engine import Batch, Context

npred = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
args = argparse.Namespace(datasets=256, chains_per_dataset=4, n=152, npred=npred, P=10, Q=2, T=3, cohesion=2, kcap=0, mixture=0)
w = bench.build_workload(args, 0)
ctx = Context(0)
b = Batch(ctx, w["model"], w["dims"], w["shared"], w["dsets"], chain_dataset=w["chain_ds"], chain_id=w["chain_id"])
b.init(1, w["labels"], w["nclu"], w["beta0"])
b.run(1, 0, 30)
b.reset_summaries()
for l in range(3):
    b.accumulate(1, 30 + l, psm=False, predictive=True)
    print("predictive step ms", b.last_timing()[0])
