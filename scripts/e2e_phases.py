"""Times the phases of the public host API call (what bench.py reports as e2e)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import argparse, bench
from treatppmx_b200.engine import Batch, Context

ap = argparse.ArgumentParser()
args = bench.main.__globals__["argparse"].Namespace(datasets=256, chains_per_dataset=4, n=152, npred=28, P=10, Q=2, T=3,
                                                    cohesion=2, kcap=0, mixture=0)
w = bench.build_workload(args, 0)
ctx = Context(0)
for rep in range(4):
    t = [time.perf_counter()]
    b = Batch(ctx, w["model"], w["dims"], w["shared"], w["dsets"], chain_dataset=w["chain_ds"], chain_id=w["chain_id"]); t.append(time.perf_counter())
    b.init(1, w["labels"], w["nclu"], w["beta0"]); t.append(time.perf_counter())
    b.sweep(1, 0, 20); t.append(time.perf_counter())
    dev = b.last_timing()[0]
    pi, _, _ = b.predict(1, 20, want=("pipred",)); t.append(time.perf_counter())
    lab, ncl = b.get_partitions(); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    names = ["create", "init", "sweep", "predict", "get_partitions", "close"]
    print(rep, " ".join(f"{n}={1e3*(t[i+1]-t[i]):.1f}ms" for i, n in enumerate(names)), f"| sweep device {dev:.2f} ms", b"".join([]) or "", "handovers", None)
