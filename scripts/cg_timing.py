"""Repeated CG + Jacobi solves of the 64^3 hex p=4 diffusion benchmark on one GPU: per-iteration time, run-to-run spread."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import l3ster_b200 as l3b  # noqa: E402
from l3ster_b200.slab import SlabOperator, make_slab  # noqa: E402
from scripts.order_sweep import node_dist, U  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = l3b.Context(0)
xs = node_dist(n)
op = SlabOperator(ctx, make_slab(xs, xs, xs, 4, 0, 1), U, "bench_diffusion3d", [1, 2, 3, 4, 5, 6])
op.solve(1e-6, 2)
for rep in range(reps):
    ctx.synchronize()
    t0 = time.perf_counter()
    _, res, it = op.solve(1e-6, 10000)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print(f"solve {rep}: {it} iterations, residual {res:.3e}, {dt:.3f} s, {dt / it * 1e3:.3f} ms per iteration", flush=True)
# the library's own driver (no Python callback in the loop)
s = op.sys
for rep in range(2):
    t0 = time.perf_counter()
    _, res, it = s.solve(1e-6, 10000)
    dt = time.perf_counter() - t0
    print(f"l3b_mf_solve_cg {rep}: {it} iterations, {dt:.3f} s incl. the D2H of x, {dt / it * 1e3:.3f} ms per iteration", flush=True)
