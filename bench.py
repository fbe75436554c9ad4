#!/usr/bin/env python
"""Benchmark of the hot path named in BASELINE.json (3-D diffusion, hex p=4, U=4, E=7):

  * `matrix_free` (default, BASELINE configs[2] — the configuration BASELINE quotes at 1/2/4/8 B200): sum-factorised operator apply
    with its NCCL halo exchange, metric "matrix-free DOFs/s"; at N > 1 the line carries `parity_vs_n1` (the N-rank apply of a seeded
    global vector against the one-rank apply of the same global mesh);
  * `assembly` (BASELINE configs[1]): element-local least-squares assembly fused with the CRS scatter, metric "assembled elements/s".

One JSON line on stdout (rank 0). The line's `metric`/`value` belong to --workload; the other workload is reported in the
`also` object of the same line with its own roofline. Extras beside the contract keys: `condensed` (assembly workload: the same
elements under CondensationPolicy::ElementBoundary, assembly + per-element Schur complements, the policy
benchmarks/Diffusion3DBenchmark.cpp runs) and `cg_solve` (matrix-free workload: the benchmark's full CG + Jacobi solve, tol 1e-6). `--impl reference` times the CPU restatement of the reference
(oracle/, built -march=native on this host) on a bounded sample of the same workload.

Timing: CUDA events on the library's stream, W warm-up steps, K timed steps between barriers, max over ranks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, U, E = 4, 4, 7
NN = (P + 1) ** 3
L = NN * U
Q = (P + 1) ** 3
# algorithmic work per unit
ASM_FLOPS_PER_ELEM = Q * (18 * NN + 7 * L * E + (L + 1) ** 2 / 2 * (2 * E + 1))  # 238.7 Mflop, reference's own DPFlops formula
MF_FLOPS_PER_ELEM = 215e3  # reference formulation: back 45k + QP 120k + forward 45k + geometry 5k
N_BND_NODES = NN - (P - 1) ** 3
# multiply-adds the matrix-free kernel issues per element (sweeps + point stage), for the fp64-pipe fraction
MF_EXECUTED_DFMA_PER_ELEM = 46.0e3
# dram__bytes_read.sum + dram__bytes_write.sum per element of the two dominant kernels, from the ncu captures under profiles/
MF_TRAFFIC_PER_ELEM = 5.99e3
MF_TRAFFIC_SOURCE = "from profile: profiles/r1b_ncu_raw.txt (5.99 kB per element incl. the y read-modify-write), scaled to this launch"
ASM_TRAFFIC_PER_ELEM = 3.57e6
ASM_TRAFFIC_SOURCE = "from profile: profiles/r1b_ncu_raw.txt (3.57 MB per element), scaled to this launch"


def mf_bytes_per_apply(n_dofs, n_elems):
    return 16 * n_dofs + n_elems * (192 + 4 * (N_BND_NODES + 1))  # x read + y write + vertices + compact node ids


def workload_string(workload, n, world):
    """config.workload — the same text in the product arm and in the reference arm"""
    if workload == "assembly":
        return (f"Diffusion3DBenchmark assembly + CRS scatter: cube [0,1]^3, {n}^3 hex p=4 per GPU, U=4, E=7, nq=5, "
                f"CondensationPolicy::None")
    return (f"Diffusion3DBenchmarkMatrixFree operator apply: {n} x {n} x {n * world} hex p=4 on [0,1]^2 x [0,{world}], "
            f"U=4, E=7, nq=5, Dirichlet T=0 on the six faces, one z-slab of {n}^3 elements per GPU")


def node_dist(n):
    # benchmarks/Diffusion3D.hpp:11-18: running sum x += 1/n (bit-equal vertices)
    dx, x, out = 1.0 / n, 0.0, []
    for _ in range(n + 1):
        out.append(x)
        x += dx
    return np.array(out)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        def num(s):
            try:
                return float(s.split()[0])
            except Exception:
                return None

        sm = [num(r[0]) for r in self.rows if len(r) >= 7 and num(r[0]) is not None]
        mx = [num(r[1]) for r in self.rows if len(r) >= 7 and num(r[1]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "raw_first": self.rows[0] if self.rows and not sm else None}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(workload, n_threads, budget_s=12.0):
    """The reference's CPU path (oracle restatement, -march=native build on this host), bounded sample."""
    from oracle import Oracle, build

    build(native=True)
    orc = Oracle(native=True)

    def make(n):
        return orc.mesh_cube(node_dist(n), order=P)

    if workload == "assembly":
        m = make(2)
        s = m.assembled_system(U)
        t = s.assemble("bench_diffusion3d", n_threads=n_threads)
        per_elem = t / 8
        n = int(max(2, min(16, math.floor((budget_s / per_elem) ** (1 / 3)))))
        for _ in range(3):  # the 2^3 probe overestimates the cost per element (thread start-up): grow until the sample is long enough
            m = make(n)
            s = m.assembled_system(U)
            t = s.assemble("bench_diffusion3d", n_threads=n_threads)
            if t > 0.4 * budget_s or n >= 16:
                break
            n = int(min(16, max(n + 1, math.floor(n * (0.8 * budget_s / t) ** (1 / 3)))))
        return dict(value=n**3 / t, unit="elements/s", cores=n_threads, kind="port",
                    sample=f"oracle assembleGlobalSystem (local assembly + CRS scatter) of {n}^3 hex p=4 elements, {t:.2f} s")
    n = 16
    m = make(n)
    s = m.matrix_free_system(U)
    s.add_kernel("bench_diffusion3d")
    x = np.random.default_rng(5489).uniform(-1, 1, size=(m.n_nodes * U, 1))
    s.apply(x, n_threads=n_threads)
    per_apply = s.last_secs
    reps = int(max(3, min(2000, math.ceil(budget_s / max(per_apply, 1e-6)))))
    s.apply(x, n_threads=n_threads, repeats=reps)
    t = s.last_secs
    return dict(value=m.n_nodes * U * reps / t, unit="DOFs/s", cores=n_threads, kind="port",
                sample=f"oracle sum-factorised operator apply on {n}^3 hex p=4 elements ({m.n_nodes * U} DOFs), {reps} applies, {t:.2f} s")


def bind_to_gpu_numa_node(device):
    """Pin this process to the cores of the NUMA node the GPU hangs off, before any host vector is allocated: pinned pages are placed by
    first touch, and a vector on the far socket crosses the inter-socket link on every PCIe copy. The reference's launch line does the same
    for its ranks (benchmarks/CMakeLists.txt:38: --map-by package --bind-to package). Returns what it found; does nothing when the box
    does not say (one node, a container without sysfs, no NVML)."""
    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]  # sysfs spells the domain with four hex digits
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info.update(pci=bdf, node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:  # a placement hint, never a reason to fail
        info["error"] = str(exc)[:120]
    return info


def physical_cores():
    try:
        import psutil

        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


# ---------------------------------------------------------------------------------------------------------------------
def asm_executed_flops_per_elem():
    """Multiply-adds the DMMA assembly kernel really issues for hex p=4 diffusion (assemble_dmma.cuh): per unknown pair (u <= v)
    the equations E_u ∩ E_v only, 32 x 32 warp tiles of the 128 x 128 node tile (above-diagonal tiles skipped for u == v),
    K = 128 padded quadrature points per equation."""
    eqs = [{1, 2, 3}, {0, 1, 5, 6}, {0, 2, 4, 6}, {0, 3, 4, 5}]  # equations each unknown appears in (benchmarks/Diffusion3D.hpp:50-79)
    macs = 0
    for u in range(4):
        for v in range(u, 4):
            n_eq = len(eqs[u] & eqs[v])
            macs += n_eq * (10 if u == v else 16) * 32 * 32 * 128
    return 2.0 * macs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="matrix_free", choices=["assembly", "matrix_free"])
    ap.add_argument("--n-asm", type=int, default=24, help="elements per edge, assembly workload (per GPU); 24^3 is the largest CRS that leaves room")
    ap.add_argument("--n-parity", type=int, default=32, help="N > 1: edge of the cube whose N-slab apply is checked against the one-rank apply")
    ap.add_argument("--n-mf", type=int, default=64, help="elements per edge, matrix-free workload (per GPU)")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workload")
    ap.add_argument("--cg-max-iters", type=int, default=10000, help="matrix-free workload: iteration cap of the CG solve timed after the applies (0 = skip)")
    ap.add_argument("--no-condensed", dest="condensed", action="store_false",
                    help="assembly workload: skip the CondensationPolicy::ElementBoundary figure (assembly + per-element Schur complements)")
    ap.add_argument("--condensed", dest="condensed", action="store_true", default=True, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-distorted", action="store_true", help="matrix-free workload: skip the distorted-mesh figure")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    metric = {"assembly": ("assembled elements/s (3D diffusion hex p=4, U=4, E=7)", "elements/s"),
              "matrix_free": ("matrix-free DOFs/s (3D diffusion hex p=4, U=4, E=7)", "DOFs/s")}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = physical_cores()
        n_steps = max(1, args.steps)
        budget = max(5.0, min(20.0, 150.0 / (n_steps + max(args.warmup, 0))))
        for _ in range(max(args.warmup, 0)):
            cpu_reference(args.workload, cores, budget_s=budget)
        t0 = time.time()
        vals = [cpu_reference(args.workload, cores, budget_s=budget) for _ in range(n_steps)]
        wall = time.time() - t0
        best = max(vals, key=lambda d: d["value"])
        mean = float(np.mean([d["value"] for d in vals]))
        line = {"impl": "reference", "metric": metric[args.workload][0], "value": mean, "unit": best["unit"], "n_gpus": args.gpus,
                "steps": n_steps, "warmup": max(args.warmup, 0), "ms_per_step": 1e3 * wall / n_steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(args.workload, args.n_asm if args.workload == "assembly" else args.n_mf, max(args.gpus, 1)),
                           "sample": "each step a bounded sample of this workload on the host cores, see cpu_baseline.sample",
                           "note": "the reference needs gcc >= 14, Eigen, Trilinos, oneTBB, MPI — none in this image — so its CPU path is "
                                   "timed through the oracle restatement (oracle/, -O3 -march=native, std::thread over the physical cores)"},
                "cpu_baseline": dict(best, value=mean), "e2e": {"value": mean, "unit": best["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # CPU baseline first (rank 0's host cores, at every N): the other ranks block in the rendezvous of init_process_group meanwhile, so
    # nothing competes for the cores and nothing spins on a GPU
    cpu_main = cpu_also = None
    if rank == 0 and not args.no_cpu_baseline:
        other_wl = "matrix_free" if args.workload == "assembly" else "assembly"
        try:
            cpu_main = cpu_reference(args.workload, physical_cores())
            if not args.no_also:
                cpu_also = cpu_reference(other_wl, physical_cores(), budget_s=8.0)
        except Exception as exc:  # the baseline is a reported extra; never lose the GPU numbers over it
            cpu_main = cpu_main or {"error": str(exc)}

    import torch

    import l3ster_b200 as l3b
    from l3ster_b200.slab import SlabOperator, make_slab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: l3ster_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner to stdout when the communicator is created: send the C-level stdout to stderr until that has
        # happened, so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime

            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=15))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    ctx = l3b.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(step_fn, steps, warmup):
        """W untimed steps, then exactly K steps between barriers: device time (CUDA events on the library's stream) and wall
        time, both as the max over ranks."""
        for _ in range(warmup):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            step_fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, max_over_ranks(1e3 * wall) / steps, clocks

    hbm_peak, hbm_src = measured_peaks()
    z_layers_mf = args.n_mf * world

    # ---- assembly workload (BASELINE configs[1]) --------------------------------------------------------------------
    def run_assembly():
        # the timed region is the reference's assembleProblem (local assembly + scatter into the rank-local CRS), which has no
        # exchange step: shared rows are exported at endAssembly (AssembledSystem.hpp:384-389), outside it. Weak scaling: every
        # rank assembles its own n^3 block.
        n = args.n_asm
        host = l3b.make_cube_mesh(node_dist(n), order=P)
        mesh = ctx.upload_mesh(host)
        graph = host.node_graph()
        sys_ = l3b.AssembledSystem(ctx, mesh, U, 1, graph)
        n_elems = host.n_elems
        crs_nnz, n_dofs_asm = sys_.nnz, sys_.n_dofs
        kernel_ms = []

        def step():
            sys_.beginAssembly()
            sys_.assembleProblem("bench_diffusion3d")
            kernel_ms.append(sys_.last_kernel_ms)

        ms, _, clocks = timed(step, args.steps, args.warmup)
        k_ms = float(np.mean(kernel_ms[-args.steps:]))
        # e2e through the C ABI with host buffers: this step's element geometry H2D (pinned), assembly, rhs D2H (pinned)
        verts_h = torch.from_numpy(np.ascontiguousarray(host.verts)).pin_memory()
        rhs_h = torch.empty(sys_.n_dofs, dtype=torch.float64).pin_memory()
        verts_np = verts_h.numpy()

        def step_e2e():
            mesh.update_verts(verts_np)            # H2D of this step's input
            sys_.beginAssembly()
            sys_.assembleProblem("bench_diffusion3d")
            sys_.download_rhs_into(rhs_h.data_ptr())  # D2H of the step's result (synchronises)

        _, wall_ms, _ = timed(step_e2e, args.steps, 1)
        condensed = None
        if args.condensed:
            # the policy benchmarks/Diffusion3DBenchmark.cpp:6 itself runs: CondensationPolicy::ElementBoundary. Same elements; a step is
            # beginAssembly + assembleProblem into the element-local storage + the Schur complements into the primary-node CRS.
            from l3ster_b200.condensation import CondensedAssembledSystem

            del sys_
            cs = CondensedAssembledSystem(ctx, host, U)

            def step_c():
                cs.beginAssembly()
                cs.assembleProblem("bench_diffusion3d")
                cs.endAssembly()

            def step_a():
                cs.beginAssembly()
                cs.assembleProblem("bench_diffusion3d")

            ms_a, _, _ = timed(step_a, 3, 2)
            cs.endAssembly()
            ms_c, _, _ = timed(step_c, 3, 2)
            condensed = {"value": world * n_elems / (ms_c * 1e-3), "unit": "elements/s", "ms_per_step": ms_c, "assemble_ms": ms_a,
                         "condense_ms": ms_c - ms_a, "primary_dofs_per_gpu": cs.n_primary_dofs, "crs_nnz_per_gpu": cs.condensed.nnz,
                         "steps": 3, "what": "beginAssembly + assembleProblem (element-local, block-diagonal storage) + endAssembly "
                                             "(K_pp - K_pi K_ii^-1 K_ip of every element into the CRS of the primary dofs), CUDA events"}
            del cs
        fp64_fma = ctx.microbench(0)
        fp64_dmma = ctx.microbench(1)
        achieved = ASM_FLOPS_PER_ELEM * n_elems / (k_ms * 1e-3) / 1e12
        executed = asm_executed_flops_per_elem() * n_elems / (k_ms * 1e-3) / 1e12
        return {
            "value": world * n_elems / (ms * 1e-3), "ms_per_step": ms,
            "e2e": {"value": world * n_elems / (wall_ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": int(verts_np.nbytes),
                    "d2h_bytes_per_step": int(n_dofs_asm * 8),
                    "what": "element geometry H2D (pinned) + beginAssembly + assembleProblem + rhs D2H (pinned) through the C ABI, wall clock; "
                            "the CRS values (%.1f GB) stay on the device, where the solve runs" % (crs_nnz * 8 / 1e9)},
            "roofline": {"bound": "tensor", "achieved": executed, "peak": fp64_dmma, "unit": "TFLOP/s", "frac": executed / fp64_dmma,
                         "traffic": ASM_TRAFFIC_PER_ELEM * n_elems,
                         "kernel": "assembleDmmaKernel<bench_diffusion3d, hex p=4> (fp64 DMMA, mma.sync.m8n8k4.f64)", "kernel_ms": k_ms,
                         "algorithmic_flops_per_element": ASM_FLOPS_PER_ELEM,
                         "peak_source": "fp64 DMMA microkernel measured in this run (MEASURED_PEAKS.json carries bf16 and HBM figures only); "
                                        "DFMA and DMMA share one pipe on B200 (interleaved microkernel: %.1f TFLOP/s)" % ctx.microbench(3),
                         "fp64_fma_tflops_measured": fp64_fma,
                         "executed_flops_per_element": asm_executed_flops_per_elem(),
                         "reference_count_tflops": achieved,
                         "note": "`achieved` / `frac` count the multiply-adds the kernel really issues on the DMMA pipe (physical fraction); "
                                 "`reference_count_tflops` is the same time against the reference's own DPFlops formula (dense symmetric "
                                 "rank update, benchmarks/LocalAssemblyBenchmarks.cpp:71-75), larger because products with structurally "
                                 "zero operator entries are skipped (33 of 112 equation x unknown-pair products survive)",
                         "traffic_source": ASM_TRAFFIC_SOURCE},
            "gpu_launches": args.steps, "clocks": clocks, "condensed": condensed,
            "config": {"workload": workload_string("assembly", n, world), "elements_per_gpu": n_elems, "dofs_per_gpu": host.n_nodes * U,
                       "crs_nnz_per_gpu": crs_nnz, "l2": "CRS values (%.1f GB) larger than L2" % (crs_nnz * 8 / 1e9),
                       "step": "beginAssembly (zero values + rhs) + assembleProblem",
                       "multi_gpu": "one n^3 block per rank; assembleProblem has no exchange step in the reference either (shared rows are "
                                    "exported at endAssembly, outside the timed region)"},
        }

    # ---- matrix-free workload (BASELINE configs[2]) -----------------------------------------------------------------
    BND = [1, 2, 3, 4, 5, 6]  # Dirichlet T = 0 on the six faces (benchmarks/Diffusion3D.hpp:39-41)

    def seeded_x(slab):
        """a vector that is a function of the GLOBAL lattice position of a node, so that every rank (and the one-rank run of the same
        global mesh) holds the same values at the nodes it shares"""
        lat = slab.lattice.astype(np.float64)
        base = np.sin(0.37 * lat[:, 0] + 1.0) * np.cos(0.23 * lat[:, 1] - 0.5) + np.sin(0.11 * lat[:, 2] + 0.3)
        return (base[:, None] * (1.0 + 0.25 * np.arange(U))[None, :] + 0.1 * np.arange(U)[None, :]).ravel()

    def parity_vs_n1():
        """strong-scaling check of the halo'd apply: a cube of n_parity^3 elements cut into `world` z-slabs, y = A x of a seeded global
        x; ||y||^2 and x.y summed over the owned dofs of all ranks against the same quantities of the one-rank apply of the whole cube
        (rank 0 computes it after the distributed one). Relative differences, bar 1e-12."""
        n = args.n_parity
        xs = node_dist(n)
        slab = make_slab(xs, xs, xs, P, rank, world)
        op = SlabOperator(ctx, slab, U, "bench_diffusion3d", BND)
        sums = torch.zeros(2, dtype=torch.float64, device="cuda")
        if slab.n_elems > 0:
            xd = torch.from_numpy(seeded_x(slab)).to("cuda")
            yd = torch.zeros_like(xd)
            torch.cuda.synchronize()
            op.apply(xd, yd, 1.0, 0.0)
            ctx.synchronize()
            no = op.n_owned_dofs
            sums = torch.stack([(yd[:no] * yd[:no]).sum(), (xd[:no] * yd[:no]).sum()])
        dist.all_reduce(sums)
        ref = torch.zeros(2, dtype=torch.float64, device="cuda")
        if rank == 0:
            whole = make_slab(xs, xs, xs, P, 0, 1)
            wop = SlabOperator(ctx, whole, U, "bench_diffusion3d", BND)
            xw = torch.from_numpy(seeded_x(whole)).to("cuda")
            yw = torch.zeros_like(xw)
            torch.cuda.synchronize()
            wop.apply(xw, yw, 1.0, 0.0)
            ctx.synchronize()
            ref = torch.stack([(yw * yw).sum(), (xw * yw).sum()])
        dist.broadcast(ref, 0)
        rel = ((sums - ref).abs() / ref.abs()).cpu().numpy()
        return {"ok": bool((rel < 1e-12).all()), "rel_diff_yy": float(rel[0]), "rel_diff_xy": float(rel[1]), "tolerance": 1e-12,
                "yy": float(sums[0]), "xy": float(sums[1]), "yy_one_rank": float(ref[0]), "xy_one_rank": float(ref[1]),
                "mesh": f"{n}^3 hex p=4 cut into {world} z-slabs vs the same cube on one rank",
                "what": "||A x||^2 and x.A x of a seeded global x, all-reduced over the owned dofs"}

    def ragged_partition_parity():
        """the general-partition path at this N (partition import, general neighbour lists, shared-row export): a 4^3 hex p=4 cube cut
        by a ragged, non-slab epart; matrix-free apply and the row-complete assembled matrix (l3b_asm_export_shared_rows) against the
        same objects on one rank (rank 0), relative differences"""
        from l3ster_b200.partition import Partition, bisection_epart
        from l3ster_b200.slab import SlabAssembledOperator

        host = l3b.make_cube_mesh(node_dist(4), order=P)
        rng = np.random.default_rng(2024)
        ep = bisection_epart(host.verts.mean(axis=1), world)
        flip = rng.random(host.n_elems) < 0.2
        ep[flip] = rng.integers(0, world, size=int(flip.sum()))
        part = Partition(3, P, host.n_nodes, host.nodes, host.verts, host.side_boundaries, world, ep.astype(np.int32))

        def seeded_g(gids):
            g = gids.astype(np.float64)
            return (np.sin(0.37 * g + 1.0)[:, None] * (1.0 + 0.25 * np.arange(U))[None, :] + 0.1 * np.arange(U)[None, :]).ravel()

        view = part.rank_view(rank, False)
        op = SlabOperator(ctx, view, U, "bench_diffusion3d", BND)
        no = view.n_owned_nodes * U
        xd = torch.from_numpy(seeded_g(view.gids)).to("cuda")
        yd = torch.zeros_like(xd)
        torch.cuda.synchronize()
        op.apply(xd, yd, 1.0, 0.0)
        ctx.synchronize()
        viewx = part.rank_view(rank, True)
        aop = SlabAssembledOperator(ctx, viewx, U, "bench_diffusion3d", BND)
        nox = viewx.n_owned_nodes * U
        xa = torch.from_numpy(seeded_g(viewx.gids)).to("cuda")
        ya = torch.zeros_like(xa)
        torch.cuda.synchronize()
        aop.apply(xa, ya)
        ctx.synchronize()
        vals, _ = aop.sys.download()
        row_ptr, _ = aop.sys.graph()
        # owned parts, placed by global id, summed over ranks (every dof is owned exactly once)
        full = torch.zeros(3, host.n_nodes * U, dtype=torch.float64, device="cuda")
        full[0, view.first_gid * U:view.first_gid * U + no] = yd[:no]
        full[1, viewx.first_gid * U:viewx.first_gid * U + nox] = ya[:nox]
        row_sq = np.add.reduceat(vals[:row_ptr[nox]] ** 2, row_ptr[:nox]) if nox else np.zeros(0)
        full[2, viewx.first_gid * U:viewx.first_gid * U + nox] = torch.from_numpy(row_sq).to("cuda")
        dist.all_reduce(full)
        out = None
        if rank == 0:
            gnodes = part.new_id[host.nodes.astype(np.int64)].astype(np.uint32)
            whole = Partition(3, P, host.n_nodes, gnodes, host.verts, host.side_boundaries, 1, np.zeros(host.n_elems, dtype=np.int32))
            wv = whole.rank_view(0, True)
            wop = SlabOperator(ctx, whole.rank_view(0, False), U, "bench_diffusion3d", BND)
            xw = torch.from_numpy(seeded_g(wv.gids)).to("cuda")
            yw = torch.zeros_like(xw)
            torch.cuda.synchronize()
            wop.apply(xw, yw, 1.0, 0.0)
            ctx.synchronize()
            wa = SlabAssembledOperator(ctx, wv, U, "bench_diffusion3d", BND)
            wvals, _ = wa.sys.download()
            wrp, _ = wa.sys.graph()
            wrow_sq = torch.from_numpy(np.add.reduceat(wvals ** 2, wrp[:-1])).to("cuda")
            rel = lambda a, b: float(torch.linalg.norm(a - b) / torch.linalg.norm(b))  # noqa: E731
            out = {"mf_apply": rel(full[0], yw), "assembled_apply": rel(full[1], yw), "assembled_row_norms": rel(full[2], wrow_sq),
                   "elements_per_rank": [int((ep == r).sum()) for r in range(world)], "tolerance": 1e-12}
            out["ok"] = bool(max(out["mf_apply"], out["assembled_apply"], out["assembled_row_norms"]) < 1e-12)
        return out

    def run_mf():
        n = args.n_mf
        xs = node_dist(n)
        # weak scaling: the mesh is n x n x (n * world), one z-slab of n layers per rank, halo exchange over NCCL inside the library
        zs = node_dist(n) if world == 1 else np.concatenate([[0.0], np.cumsum(np.full(z_layers_mf, 1.0 / n))])
        slab = make_slab(xs, xs, zs, P, rank, world)
        t0 = time.perf_counter()
        op = SlabOperator(ctx, slab, U, "bench_diffusion3d", BND)
        ctx.synchronize()
        init_s = time.perf_counter() - t0
        n_local, n_owned, n_elems = op.n_local_dofs, op.n_owned_dofs, slab.n_elems
        numa = bind_to_gpu_numa_node(local_rank)  # before the host vectors are pinned (first touch places them)
        xh = torch.from_numpy(seeded_x(slab)).pin_memory()
        yh = torch.empty(n_local, dtype=torch.float64).pin_memory()
        xd = xh.to("cuda")
        yd = torch.zeros_like(xd)
        torch.cuda.synchronize()  # filled on torch's stream; the library works on its own non-blocking stream

        def step():
            op.apply(xd, yd, 1.0, 0.0)

        ms, _, clocks = timed(step, args.steps, args.warmup)
        launches = op.launches

        # the dominant kernel alone, live: K launches of the element phase over all of this rank's elements between two events
        def step_kernel():
            op.sys.apply_phase_device(xd.data_ptr(), yd.data_ptr(), l3b.APPLY_ELEMENTS, 0, n_elems)

        k_ms, _, _ = timed(step_kernel, args.steps, 2)
        # distorted mesh (every element non-affine: the general point stage instead of the axis-aligned one): device time of the same apply
        distorted_ms = None
        if world == 1 and not args.no_distorted:
            verts = np.array(slab.verts)
            bump = 0.15 / n
            verts[..., 0] += bump * np.sin(2 * np.pi * verts[..., 1]) * np.sin(2 * np.pi * verts[..., 2])
            verts[..., 1] += bump * np.sin(2 * np.pi * verts[..., 0]) * np.sin(2 * np.pi * verts[..., 2])
            op.mesh.update_verts(np.ascontiguousarray(verts))
            distorted_ms, _, _ = timed(step, args.steps, 2)
            op.mesh.update_verts(np.ascontiguousarray(slab.verts))

        xh_np, yh_np = xh.numpy(), yh.numpy()

        def step_e2e():  # the reference-facing call with HOST buffers: l3b_mf_apply = H2D x, halo'd apply, D2H y, synchronised
            op.sys.apply_raw(xh_np, yh_np, 1, 1.0, 0.0)

        _, wall_ms, _ = timed(step_e2e, args.steps, 1)
        e2e_info = op.sys.host_apply_info()
        # the same call in its serial form (x in, apply, y out one after the other) and the check that the host-vector call returns
        # what the device-resident apply returns
        op.sys.set_host_apply(0)
        _, serial_wall_ms, _ = timed(step_e2e, max(3, args.steps // 4), 1)
        op.sys.set_host_apply(1)
        step_e2e()
        step()
        ctx.synchronize()
        torch.cuda.synchronize()
        yd_np = yd.cpu().numpy()
        e2e_diff = max_over_ranks(float(np.linalg.norm(yh_np[:n_owned] - yd_np[:n_owned]) / max(np.linalg.norm(yd_np[:n_owned]), 1e-300)))
        del yd_np
        owned_total = sum_over_ranks(float(n_owned))
        parity = parity_vs_n1() if world > 1 else None
        if parity is not None:
            rp = ragged_partition_parity()
            if rank == 0:
                parity["ragged_partition"] = rp
                parity["ok"] = bool(parity["ok"] and rp["ok"])
        cg = None
        if args.cg_max_iters > 0:
            # the benchmark's solve (benchmarks/Diffusion3D.hpp:115-118): CG + native Jacobi, tol 1e-6 (absolute), x0 = 0, rhs of the source f = 1
            del yd
            op.solve(1e-6, 2)  # loads the solver's kernels
            cg_s, cg_all = float("inf"), []
            cg_sampler = ClockSampler(local_rank)
            cg_sampler.start()
            for _ in range(3):  # three solves, the fastest one is reported, all are listed (a 2 s solve can run into the power cap)
                barrier()
                t0 = time.perf_counter()
                _, res, iters = op.solve(1e-6, args.cg_max_iters)
                ctx.synchronize()
                cg_all.append(max_over_ranks(time.perf_counter() - t0))
                cg_s = min(cg_s, cg_all[-1])
            cg_clocks = cg_sampler.stop()
            cg = {"iters": int(iters), "achieved_residual": float(res), "seconds": cg_s, "ms_per_iteration": cg_s * 1e3 / max(iters, 1),
                  "dofs_per_s": owned_total * iters / cg_s, "tol": 1e-6, "max_iters": args.cg_max_iters, "seconds_all": cg_all, "clocks": cg_clocks,
                  "what": "full CG + Jacobi solve on the device (l3b_mf_solve_device: halo'd apply with p.Ap fused + 10 vector passes + "
                          "ncclAllReduce of the dot products per iteration), wall clock, max over ranks, fastest of three solves"}
        bytes_alg = mf_bytes_per_apply(n_owned, n_elems)
        gbs = bytes_alg / (k_ms * 1e-3) / 1e9
        fp64_fma = ctx.microbench(0)
        res = {
            "value": owned_total / (ms * 1e-3), "ms_per_step": ms,
            "e2e": {"value": owned_total / (wall_ms * 1e-3), "unit": "DOFs/s", "h2d_bytes_per_step": int(n_local * 8),
                    "d2h_bytes_per_step": int(n_local * 8), "ms_per_step": wall_ms, "streamed": e2e_info, "serial_form_ms_per_step": serial_wall_ms,
                    "rel_diff_vs_device_apply": e2e_diff, "rel_diff_ok": bool(e2e_diff < 1e-13), "host_numa_binding": numa,
                    "what": "l3b_mf_apply through the C ABI with pinned HOST vectors, synchronised, wall clock, max over ranks. Streamed form "
                            "(l3b_mf_set_host_apply, the default): x blocks H2D, element chunks and y blocks D2H run as three concurrent "
                            "streams, so the call costs about one PCIe copy instead of two copies plus the apply (serial_form_ms_per_step); "
                            "bound by PCIe, not by the kernel"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "traffic": MF_TRAFFIC_PER_ELEM * n_elems if MF_TRAFFIC_PER_ELEM else None,
                         "kernel": "mfHexPlanesKernel<bench_diffusion3d, hex p=4, nq=5>", "kernel_ms": k_ms,
                         "algorithmic_bytes": bytes_alg, "algorithmic_bytes_per_dof": bytes_alg / max(n_owned, 1), "peak_source": hbm_src,
                         "kernel_share_of_step": k_ms / ms,
                         "fp64_pipe": {"executed_dfma_per_element": MF_EXECUTED_DFMA_PER_ELEM,
                                       "executed_tflops": 2 * MF_EXECUTED_DFMA_PER_ELEM * n_elems / (k_ms * 1e-3) / 1e12,
                                       "peak_tflops_measured": fp64_fma,
                                       "frac": 2 * MF_EXECUTED_DFMA_PER_ELEM * n_elems / (k_ms * 1e-3) / 1e12 / fp64_fma,
                                       "reference_formulation_tflops": MF_FLOPS_PER_ELEM * n_elems / (k_ms * 1e-3) / 1e12},
                         "note": "contract roofline is HBM (north_star); the kernel itself is bound by the fp64 pipe (arithmetic intensity "
                                 "~45 flop/B), so the fp64 fraction of the multiply-adds it really issues is reported beside it",
                         "traffic_source": MF_TRAFFIC_SOURCE},
            "gpu_launches": launches * args.steps, "clocks": clocks, "cg_solve": cg,
            "config": {"workload": workload_string("matrix_free", n, world), "elements_per_gpu": n_elems, "owned_dofs_per_gpu": n_owned, "ghost_dofs_per_gpu": n_local - n_owned,
                       "l2": "x and y (%.0f MB each) larger than L2" % (n_local * 8 / 1e6), "init_diag_rhs_s": init_s,
                       "step": "y = A x: zero y, element kernel, Dirichlet rows" + ("; Import of x behind the zeroing, border elements, Export of y "
                               "behind the interior elements, unpack-add" if world > 1 else ""),
                       "multi_gpu": ("NCCL send/recv halo exchange inside the library (l3b_halo: Import x / Export y of the interface plane, "
                                     "%.1f MB per direction), hidden behind the zeroing of y and the interior elements"
                                     % ((n * P + 1) ** 2 * U * 8 / 1e6)) if world > 1 else "single GPU"},
        }
        if distorted_ms is not None:
            res["distorted_mesh"] = {"ms_per_step": distorted_ms, "value": owned_total / (distorted_ms * 1e-3),
                                     "what": "the same apply with every element non-affine (general J^-1 per point instead of the axis-aligned path)"}
        if parity is not None:
            res["parity_vs_n1"] = parity
        return res

    runners = {"assembly": run_assembly, "matrix_free": run_mf}
    main_res = runners[args.workload]()
    also = None
    if not args.no_also:
        other = "matrix_free" if args.workload == "assembly" else "assembly"
        r = runners[other]()
        also = {"metric": metric[other][0], "unit": metric[other][1], "steps": args.steps, **r}

    line = {"metric": metric[args.workload][0], "value": main_res["value"], "unit": metric[args.workload][1], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": main_res["config"],
            "e2e": main_res["e2e"], "roofline": main_res["roofline"], "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"]}
    for extra in ("cg_solve", "condensed", "parity_vs_n1", "distorted_mesh"):
        if main_res.get(extra):
            line[extra] = main_res[extra]
    if also:
        line["also"] = also
    if cpu_main is not None:
        line["cpu_baseline"] = cpu_main
    if also and cpu_also is not None:
        also["cpu_baseline"] = cpu_also
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
