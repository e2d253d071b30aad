import os, sys, time, cProfile, pstats
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import argparse, bench
from treatppmx_b200.engine import Context
args = argparse.Namespace(datasets=256, chains_per_dataset=4, n=152, npred=28, P=10, Q=2, T=3, cohesion=2, kcap=0, mixture=0)
w = bench.build_workload(args, 0)
ctx = Context(0)
for _ in range(3): bench.e2e_call(ctx, w, 1, 20)
pr = cProfile.Profile(); pr.enable()
t0=time.perf_counter()
for _ in range(10): bench.e2e_call(ctx, w, 1, 20)
dt=time.perf_counter()-t0
pr.disable()
print("per call ms", 1e3*dt/10)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
