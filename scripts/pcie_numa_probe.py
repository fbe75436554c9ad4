"""PCIe copy rates of one GPU with pinned host memory, optionally with the process bound to the GPU's NUMA node before the memory is
pinned (first touch decides where the pages live). Run one copy per GPU at the same time to see what the ranks of a multi-GPU bench share.
Usage: python scripts/pcie_numa_probe.py <device> <bind 0|1> [MB]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402


def main():
    dev, bind = int(sys.argv[1]), int(sys.argv[2])
    mb = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    info = bind_to_gpu_numa_node(dev) if bind else {"bound": False}
    import torch

    torch.cuda.set_device(dev)
    n = mb * (1 << 20) // 8
    xh, yh = torch.zeros(n, dtype=torch.float64).pin_memory(), torch.zeros(n, dtype=torch.float64).pin_memory()
    xd, yd = xh.to("cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            xd.copy_(xh, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            yh.copy_(yd, non_blocking=True)

    out = {"device": dev, "numa": info, "MB": mb}
    time.sleep(max(0.0, 12.0 - (time.time() % 12.0)) if len(sys.argv) > 4 else 0.0)  # crude rendezvous of concurrent copies of this script
    for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
        for f in fns:
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            for f in fns:
                f()
        torch.cuda.synchronize()
        out[name + "_GBs_per_direction"] = round(20 * n * 8 / (time.perf_counter() - t0) / 1e9, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
