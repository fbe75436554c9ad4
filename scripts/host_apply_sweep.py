"""Host-vector operator apply (l3b_mf_apply) at the benchmark's size: the serial form, the streamed form over a sweep of chunk counts and
block sizes, and the PCIe bounds they sit between (one-way H2D, one-way D2H, both at once on two streams). One JSON line per measurement.
Usage: python scripts/host_apply_sweep.py [n_elements_per_edge] [repeats]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import l3ster_b200 as l3b  # noqa: E402
from l3ster_b200.slab import SlabOperator, make_slab  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ctx = l3b.Context(0)
    xs = np.concatenate([[0.0], np.cumsum(np.full(n, 1.0 / n))])
    slab = make_slab(xs, xs, xs, 4, 0, 1)
    op = SlabOperator(ctx, slab, 4, "bench_diffusion3d", [1, 2, 3, 4, 5, 6])
    nd = op.n_local_dofs
    xh = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, nd)).pin_memory()
    yh = torch.empty(nd, dtype=torch.float64).pin_memory()
    xd, yd = xh.to("cuda"), torch.empty(nd, dtype=torch.float64, device="cuda")
    nbytes = nd * 8

    def wall(fn, k=reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / k

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            xd.copy_(xh, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            yh.copy_(yd, non_blocking=True)

    def both():
        h2d()
        d2h()

    for name, fn in (("h2d", h2d), ("d2h", d2h), ("h2d+d2h concurrently", both)):
        ms = wall(fn)
        print(json.dumps({"pcie": name, "ms": ms, "GB/s per direction": nbytes / ms / 1e6}))
    xn, yn = xh.numpy(), yh.numpy()

    def apply():
        op.sys.apply_raw(xn, yn, 1, 1.0, 0.0)

    op.sys.set_host_apply(0)
    print(json.dumps({"host_apply": "serial", "ms": wall(apply)}))
    ref = yn.copy()
    for chunks, block in ((48, 8192), (12, 8192), (24, 8192), (96, 8192), (192, 8192), (48, 2048), (48, 65536), (24, 65536), (96, 32768)):
        op.sys.set_host_apply(1, chunks, block)
        ms = wall(apply)
        print(json.dumps({"host_apply": "streamed", "chunks": chunks, "block_nodes": block, "ms": ms, "info": op.sys.host_apply_info(),
                          "rel_diff_vs_serial": float(np.linalg.norm(yn - ref) / np.linalg.norm(ref))}))


if __name__ == "__main__":
    main()
