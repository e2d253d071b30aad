"""e2e steps per second of the public host call (bench.e2e_call: create + H2D, init, 20 sweeps, predictive, D2H,
destroy) with N batches in flight (N host threads, one context each):  python scripts/e2e_pipeline.py [reps] [N ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import argparse
from concurrent.futures import ThreadPoolExecutor
import bench
from treatppmx_b200.engine import Context

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ns = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 4, 6]
args = argparse.Namespace(datasets=256, chains_per_dataset=4, n=152, npred=28, P=10, Q=2, T=3, cohesion=2, kcap=0, mixture=0)
w = bench.build_workload(args, 0)
nre = 1024 * 152 * 20
for n in ns:
    ctxs = [Context(0) for _ in range(n)]
    for c in ctxs:
        bench.e2e_call(c, w, 1, 2)
    def work(c):
        for _ in range(reps):
            bench.e2e_call(c, w, 1, 20)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(n) as ex:
        list(ex.map(work, ctxs))
    dt = (time.perf_counter() - t0) / (n * reps)
    print(f"in flight {n}: {1e3 * dt:.3f} ms per step, {nre / dt:.3e} reallocations/s")
    for c in ctxs:
        c.close()
