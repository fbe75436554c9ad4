"""Times the 64^3 hex p=4 matrix-free apply with and without the fused x^T A x (ElemArgs::energy)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import l3ster_b200 as l3b  # noqa: E402
from scripts.order_sweep import node_dist, U  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = l3b.Context(0)
hm = l3b.make_cube_mesh(node_dist(n), order=4)
mm = ctx.upload_mesh(hm)
mask = np.zeros(hm.n_nodes * U, dtype=np.uint8)
mask[hm.boundary_nodes([1, 2, 3, 4, 5, 6]) * U] = 1
s = l3b.MatrixFreeSystem(ctx, mm, U, 1, mask, None)
s.assembleProblem("bench_diffusion3d")
s.endAssembly()
x = torch.rand(s.n_dofs, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
e = torch.zeros(1, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream)
for ep in (None, e.data_ptr(), None, e.data_ptr()):
    for _ in range(3):
        s.apply_device(x.data_ptr(), y.data_ptr(), energy_ptr=ep)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        s.apply_device(x.data_ptr(), y.data_ptr(), energy_ptr=ep)
    e1.record(stream)
    ctx.synchronize()
    print("energy" if ep else "plain ", e0.elapsed_time(e1) / 10, "ms per apply", flush=True)
print("x.Ax", float(torch.dot(x, y)), "fused", e.item() / 26)
