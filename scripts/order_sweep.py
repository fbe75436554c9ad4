"""BASELINE configs[4]: local-assembly order sweep on 3-D hexahedra (benchmarks/LocalAssemblyBenchmarks.cpp:41-87), CUDA-core (DFMA,
assemble.cuh) against tensor-core (DMMA, assemble_dmma.cuh) kernel, and the matrix-free apply, per element order.

    python scripts/order_sweep.py            # both assembly kernels (the DFMA one in a subprocess: the selection is per process)

Flops are the reference's own DPFlops count (LocalAssemblyBenchmarks.cpp:71-75). One JSON line per (order, kernel)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import l3ster_b200 as l3b  # noqa: E402

U, E = 4, 7


def node_dist(n):
    dx, x, out = 1.0 / n, 0.0, []
    for _ in range(n + 1):
        out.append(x)
        x += dx
    return np.array(out)


def ref_flops(p):
    nn, q = (p + 1) ** 3, (p + 1) ** 3
    L = nn * U
    return q * (18 * nn + 7 * L * E + (L + 1) ** 2 / 2 * (2 * E + 1))


def run(kind, orders=(1, 2, 3, 4, 5, 6)):
    import torch

    ctx = l3b.Context(0)
    for p in orders:
        line = {"order": p, "kernel": kind}
        if kind != "mf":
            # elements so that the CRS stays below ~6 GB
            n = {1: 40, 2: 28, 3: 20, 4: 14, 5: 10, 6: 8, 7: 6, 8: 5}[p]
            host = l3b.make_cube_mesh(node_dist(n), order=p)
            mesh = ctx.upload_mesh(host)
            a = l3b.AssembledSystem(ctx, mesh, U, 1, host.node_graph())
            ms = []
            for it in range(5):
                a.beginAssembly()
                a.assembleProblem("bench_diffusion3d")
                ms.append(a.last_kernel_ms)
            k_ms = float(np.mean(ms[2:]))
            line.update({"elements": host.n_elems, "kernel_ms": k_ms, "elements_per_s": host.n_elems / (k_ms * 1e-3),
                         "tflops_reference_count": ref_flops(p) * host.n_elems / (k_ms * 1e-3) / 1e12})
            del a
        if kind in ("dmma", "mf"):  # the matrix-free apply of the same order, once
            nm = {1: 128, 2: 96, 3: 80, 4: 64, 5: 48, 6: 40, 7: 32, 8: 28}[p]
            hm = l3b.make_cube_mesh(node_dist(nm), order=p)
            mm = ctx.upload_mesh(hm)
            mask = np.zeros(hm.n_nodes * U, dtype=np.uint8)
            mask[hm.boundary_nodes([1, 2, 3, 4, 5, 6]) * U] = 1
            s = l3b.MatrixFreeSystem(ctx, mm, U, 1, mask, None)
            s.assembleProblem("bench_diffusion3d")
            s.endAssembly()
            x = torch.rand(s.n_dofs, dtype=torch.float64, device="cuda")
            y = torch.zeros_like(x)
            torch.cuda.synchronize()
            stream = torch.cuda.ExternalStream(ctx.stream)
            for _ in range(3):
                s.apply_device(x.data_ptr(), y.data_ptr())
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                s.apply_device(x.data_ptr(), y.data_ptr())
            e1.record(stream)
            ctx.synchronize()
            t = e0.elapsed_time(e1) / 10
            line.update({"mf_elements": hm.n_elems, "mf_dofs": s.n_dofs, "mf_ms_per_apply": t, "mf_gdofs_per_s": s.n_dofs / (t * 1e-3) / 1e9,
                         "mf_bytes_per_dof": (16 * s.n_dofs + hm.n_elems * (192 + 4 * ((p + 1) ** 3 - (p - 1) ** 3 + 1))) / s.n_dofs})
            del s, mm, hm
        print(json.dumps(line), flush=True)


def executed_flops(eq_sets, n_eq_total, nn, q):
    """multiply-adds x 2 the DMMA assembly kernel issues per element (assemble_dmma.cuh): per unknown pair (u <= v) the equations
    E_u ∩ E_v, 32 x 32 warp tiles of the TM x TN node tiles (for u == v: warp tiles above the diagonal skipped),
    quadrature points padded to whole chunks"""
    tn = 128 if nn > 64 else 64 if nn > 32 else 32
    n_blk = -(-nn // tn)
    kcmax = max(32, 4 * n_eq_total)
    macs = 0.0
    for u in range(len(eq_sets)):
        for v in range(u, len(eq_sets)):
            n_eq = len(eq_sets[u] & eq_sets[v])
            if n_eq == 0:
                continue
            qc = min(32, (kcmax // n_eq) & ~3)
            k_rows = -(-q // qc) * qc * n_eq
            units = 0.0
            for rb in range(n_blk):
                for cb in range(n_blk):
                    if u == v and cb * tn > rb * tn + tn - 1:
                        continue
                    for wy in range(tn // 32):
                        for wx in range(tn // 32):
                            r0, c0 = rb * tn + wy * 32, cb * tn + wx * 32
                            if u == v and c0 > r0 + 31:
                                continue
                            units += 1.0  # (a diagonal warp tile issues 10 of its 16 DMMA tiles; counted whole, as bench.py does)
            macs += units * 32 * 32 * k_rows
    return 2.0 * macs


NS3D_EQS = [{0, 1, 2, 4, 5, 6}, {0, 1, 2, 3, 5, 6}, {0, 1, 2, 3, 4, 6}, {0}, {1, 2, 3, 7}, {0, 2, 4, 7}, {0, 1, 5, 7}]  # benchmarks/Kernels.hpp:3-65


def run_ns3d(orders=(2, 4)):
    """BASELINE configs[4] as the reference's harness runs it (benchmarks/LocalAssemblyBenchmarks.cpp:41-87): NS3D kernel, U = 7, E = 8,
    n_fields = 7, quadrature order 4p - 1 (= AssemblyOptions{1, 1}, nq = 2p), DPFlops counter of :71-75"""
    ctx = l3b.Context(0)
    U7, E8, NF = 7, 8, 7
    opts = l3b.AssemblyOptions(value_order=1, derivative_order=1)
    peak = ctx.microbench(1)
    for p in orders:
        n = {2: 16, 3: 10, 4: 6}[p]
        host = l3b.make_cube_mesh(node_dist(n), order=p)
        mesh = ctx.upload_mesh(host)
        fields = ctx.upload_fields(np.random.default_rng(0).uniform(-1, 1, size=(NF, host.n_nodes)))
        a = l3b.AssembledSystem(ctx, mesh, U7, 1, host.node_graph())
        ms = []
        for it in range(5):
            a.beginAssembly()
            a.assembleProblem("ns3d_kernel", fields=fields, asm_opts=opts)
            ms.append(a.last_kernel_ms)
        k_ms = float(np.mean(ms[2:]))
        nn, q = (p + 1) ** 3, (2 * p) ** 3
        L = nn * U7
        ref = q * (18 * nn + 8 * NF * nn + 7 * L * E8 + (L + 1) ** 2 / 2 * (2 * E8 + 1))
        ex = executed_flops(NS3D_EQS, E8, nn, q)
        print(json.dumps({"order": p, "kernel": "ns3d", "assembly_kernel": "dmma" if nn >= 32 else "dfma", "elements": host.n_elems, "nq": 2 * p,
                          "kernel_ms": k_ms, "elements_per_s": host.n_elems / (k_ms * 1e-3),
                          "tflops_reference_count": ref * host.n_elems / (k_ms * 1e-3) / 1e12,
                          "executed_tflops": ex * host.n_elems / (k_ms * 1e-3) / 1e12 if nn >= 32 else None,
                          "dmma_peak_tflops_measured": peak,
                          "frac_executed": ex * host.n_elems / (k_ms * 1e-3) / 1e12 / peak if nn >= 32 else None}), flush=True)
        del a


def run_quads(orders=(1, 2, 3, 4, 5, 6, 7, 8)):
    """the 2-D half: quads, diffusion (U = 3, E = 4): assembly (the kernel the size rule picks) and matrix-free apply per order"""
    import torch

    ctx = l3b.Context(0)
    U2, E2 = 3, 4
    for p in orders:
        n = max(16, 2048 // p)
        host = l3b.make_square_mesh(node_dist(n), order=p)
        mesh = ctx.upload_mesh(host)
        a = l3b.AssembledSystem(ctx, mesh, U2, 1, host.node_graph())
        ms = []
        for it in range(5):
            a.beginAssembly()
            a.assembleProblem("bench_diffusion2d")
            ms.append(a.last_kernel_ms)
        k_ms = float(np.mean(ms[2:]))
        nn = (p + 1) ** 2
        L = nn * U2
        flops = nn * (8 * nn + 5 * L * E2 + (L + 1) ** 2 / 2 * (2 * E2 + 1))  # LocalAssemblyBenchmarks.cpp:71-75 with D = 2
        line = {"order": p, "kernel": "quad", "assembly_kernel": "dmma" if nn >= 32 else "dfma", "elements": host.n_elems, "kernel_ms": k_ms,
                "elements_per_s": host.n_elems / (k_ms * 1e-3), "tflops_reference_count": flops * host.n_elems / (k_ms * 1e-3) / 1e12}
        del a
        mask = np.zeros(host.n_nodes * U2, dtype=np.uint8)
        mask[host.boundary_nodes([1, 2, 3, 4]) * U2] = 1
        s = l3b.MatrixFreeSystem(ctx, mesh, U2, 1, mask, None)
        s.assembleProblem("bench_diffusion2d")
        s.endAssembly()
        x = torch.rand(s.n_dofs, dtype=torch.float64, device="cuda")
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(ctx.stream)
        for _ in range(3):
            s.apply_device(x.data_ptr(), y.data_ptr())
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            s.apply_device(x.data_ptr(), y.data_ptr())
        e1.record(stream)
        ctx.synchronize()
        t = e0.elapsed_time(e1) / 10
        line.update({"mf_dofs": s.n_dofs, "mf_ms_per_apply": t, "mf_gdofs_per_s": s.n_dofs / (t * 1e-3) / 1e9})
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ns3d":
        run_ns3d(tuple(int(a) for a in sys.argv[2:]) or (2, 4))
    elif len(sys.argv) > 1 and sys.argv[1] == "quad":
        run_quads(tuple(int(a) for a in sys.argv[2:]) or (1, 2, 3, 4, 5, 6, 7, 8))
    elif len(sys.argv) > 2:  # e.g. `order_sweep.py mf 4 5 6`: the matrix-free apply only, these orders
        run(sys.argv[1], tuple(int(a) for a in sys.argv[2:]))
    elif len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        run("dmma", (1, 2, 3, 4, 5, 6, 7, 8))
        subprocess.run([sys.executable, os.path.abspath(__file__), "dfma"], env=dict(os.environ, L3B_ASM_FMA="1"), check=True)
        run_quads()
        run_ns3d()
