"""BASELINE configs 1-2: ONE chain through the drop-in entry point (the reference's 42 arguments ->
18 outputs), iter = 1100, burn = 100, thin = 1 (R/dm_ppmx_ct.R:89-93), on the device and - where
oracle/_ref/libref.so exists - through the reference's own dm_ppmx_ct on one host thread."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import refrun
from treatppmx_b200.engine import Context
from treatppmx_b200.ppmxct import dm_ppmx_ct

iters, burn, thin = 1100, 100, 1
names = ("default_ngg", "dp_T3")
if len(sys.argv) > 1:  # python scripts/single_chain.py CASE [ITERS]  (short runs for launch lists)
    names = (sys.argv[1],)
    if len(sys.argv) > 2:
        iters = int(sys.argv[2]); burn = iters // 10
for name in names:
    args, logV, case = refrun.reference_args(name, iters, burn, thin, False)
    ctx = Context(0)
    dm_ppmx_ct(ctx, **dict(args, iter=20, burn=10), seed=1, logV_compact=logV)  # warm
    t0 = time.perf_counter()
    out = dm_ppmx_ct(ctx, **args, seed=3, logV_compact=logV)
    dt = time.perf_counter() - t0
    ctx.close()
    line = f"{name}: n={args['nobs']} T={args['A']} 1 chain x {iters} iterations: device {dt:.3f} s ({1e3 * dt / iters:.3f} ms/iteration), mean nclus {out['nclus'].mean():.2f}"
    if os.path.exists(refrun.REF_SO):
        argsd, _, _ = refrun.reference_args(name, iters, burn, thin, True)
        R = refrun.RefLib()
        t0 = time.perf_counter()
        R.dm_ppmx_ct(argsd, 5)
        dr = time.perf_counter() - t0
        line += f"; reference's own code, one host thread {dr:.3f} s ({1e3 * dr / iters:.3f} ms/iteration) -> {dr / dt:.1f}x"
    print(line)
