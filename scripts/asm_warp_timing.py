"""Experiment (needs a library built with EXTRA=-DL3B_ASM_TIMING, see csrc/Makefile): per-warp clocks of the assembly kernel's chunk
loop — panel build, DMMA contraction, wait at the chunk barrier — averaged per kind of tile. Usage:
    make -C l3ster_b200/csrc EXTRA=-DL3B_ASM_TIMING BUILD=build_alt OUT=../libalt.so
    L3B_LIB_PATH=l3ster_b200/libalt.so python scripts/asm_warp_timing.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import l3ster_b200 as l3b  # noqa: E402


def node_dist(n):
    dx, x, o = 1.0 / n, 0.0, []
    for _ in range(n + 1):
        o.append(x)
        x += dx
    return np.array(o)


ctx = l3b.Context(0)
host = l3b.make_cube_mesh(node_dist(7), order=4)
mesh = ctx.upload_mesh(host)
a = l3b.AssembledSystem(ctx, mesh, 4, 1, host.node_graph())
for _ in range(3):
    a.beginAssembly()
    a.assembleProblem("bench_diffusion3d")
ctx.synchronize()
n = 3430
buf = np.zeros((n, 16, 4), dtype=np.int64)
L = l3b.lib()
L.l3b_debug_asm_timing.argtypes = [C.c_void_p, C.c_int]
assert L.l3b_debug_asm_timing(buf.ctypes.data, n) == 0
print("kernel ms", a.last_kernel_ms)
for key in sorted(set(buf[:, 0, 3])):
    sel = buf[buf[:, 0, 3] == key]
    print(f"n_eq={key // 16} diag={key % 16} CTAs={len(sel)}")
    for w in range(16):
        print(f"   warp {w:2d}: build {sel[:, w, 0].mean():8.0f}  mma {sel[:, w, 1].mean():8.0f}  barrier {sel[:, w, 2].mean():8.0f}")
