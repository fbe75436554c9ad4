"""Repeated CG + Jacobi solves of the 64^3 hex p=4 diffusion benchmark on one GPU: per-iteration time and its run-to-run spread, with
the SM / memory clocks, power and temperatures sampled over every solve (pynvml). Usage: python scripts/cg_timing.py [n] [reps]"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import l3ster_b200 as l3b  # noqa: E402
from l3ster_b200.slab import SlabOperator, make_slab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6


def node_dist(k):
    dx, x, out = 1.0 / k, 0.0, []
    for _ in range(k + 1):
        out.append(x)
        x += dx
    return np.array(out)


class Sampler:
    def __init__(self):
        import pynvml

        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.rows, self.on = [], False

    def _loop(self):
        nv = self.nv
        while self.on:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1e3, nv.nvmlDeviceGetTemperature(self.h, nv.NVML_TEMPERATURE_GPU),
                                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception as exc:  # noqa: BLE001
                self.rows.append(("err", str(exc)))
            time.sleep(0.05)

    def __enter__(self):
        self.rows, self.on = [], True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.on = False
        self.t.join()

    def summary(self):
        r = [x for x in self.rows if x[0] != "err"]
        if not r:
            return str(self.rows[:1])
        a = np.array([x[:4] for x in r], dtype=float)
        reasons = 0
        for x in r:
            reasons |= int(x[4])
        return (f"sm {a[:, 0].min():.0f}-{a[:, 0].max():.0f} MHz, mem {a[:, 1].min():.0f}-{a[:, 1].max():.0f} MHz, power {a[:, 2].mean():.0f} W "
                f"(max {a[:, 2].max():.0f}), temp {a[:, 3].max():.0f} C, throttle reasons 0x{reasons:x}, {len(r)} samples")


ctx = l3b.Context(0)
xs = node_dist(n)
op = SlabOperator(ctx, make_slab(xs, xs, xs, 4, 0, 1), 4, "bench_diffusion3d", [1, 2, 3, 4, 5, 6])
op.solve(1e-6, 2)
smp = Sampler()
xsol = torch.zeros(op.n_local_dofs, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for rep in range(reps):
    ctx.synchronize()
    with smp:
        t0 = time.perf_counter()
        if rep % 2 == 0:  # through SlabOperator.solve (allocates the solution vector) ...
            _, res, it = op.solve(1e-6, 10000)
        else:             # ... and straight into the library with a vector that already exists
            res, it = op.sys.solve_device(xsol.data_ptr(), "cg", 1e-6, 10000)
        ctx.synchronize()
        dt = time.perf_counter() - t0
    print(f"solve {rep} ({'SlabOperator.solve' if rep % 2 == 0 else 'l3b_mf_solve_device'}): {it} iterations, residual {res:.3e}, {dt:.3f} s, "
          f"{dt / it * 1e3:.3f} ms per iteration | {smp.summary()}", flush=True)
# the apply alone, 800 in a row, for comparison
x = torch.rand(op.n_local_dofs, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
torch.cuda.synchronize()
for rep in range(3):
    with smp:
        t0 = time.perf_counter()
        for _ in range(800):
            op.apply(x, y)
        ctx.synchronize()
        dt = time.perf_counter() - t0
    print(f"800 applies {rep}: {dt / 800 * 1e3:.3f} ms per apply | {smp.summary()}", flush=True)
