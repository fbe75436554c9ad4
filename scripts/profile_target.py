"""Short, fixed workload for ncu: a few assembly steps and matrix-free applies of the benchmark configuration
(hex p=4, U=4, E=7) at a size that keeps ncu's replays cheap. Usage: python scripts/profile_target.py [n_asm] [n_mf]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import l3ster_b200 as l3b  # noqa: E402


def node_dist(n):
    dx, x, out = 1.0 / n, 0.0, []
    for _ in range(n + 1):
        out.append(x)
        x += dx
    return np.array(out)


def main():
    n_asm = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_mf = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    ctx = l3b.Context(0)
    if n_asm > 0:
        host = l3b.make_cube_mesh(node_dist(n_asm), order=4)
        mesh = ctx.upload_mesh(host)
        a = l3b.AssembledSystem(ctx, mesh, 4)
        for _ in range(3):
            a.beginAssembly()
            a.assembleProblem("bench_diffusion3d")
        print("assembly kernel ms", a.last_kernel_ms, "elements", host.n_elems)
    if n_mf > 0:
        host = l3b.make_cube_mesh(node_dist(n_mf), order=4)
        mesh = ctx.upload_mesh(host)
        mask = np.zeros(host.n_nodes * 4, dtype=np.uint8)
        mask[host.boundary_nodes([1, 2, 3, 4, 5, 6]) * 4] = 1
        m = l3b.MatrixFreeSystem(ctx, mesh, 4, 1, mask, None)
        m.assembleProblem("bench_diffusion3d")
        import time
        ctx.synchronize()
        t0 = time.perf_counter()
        m.endAssembly()
        ctx.synchronize()
        print("mf endAssembly (diag + rhs) ms", 1e3 * (time.perf_counter() - t0))
        x = np.random.default_rng(0).uniform(-1, 1, size=(m.n_dofs, 1))
        for _ in range(4):
            y = m.apply(x)
        print("mf apply done, |y| =", float(np.linalg.norm(y)), "dofs", m.n_dofs)


if __name__ == "__main__":
    main()
