"""torchrun worker of tests/test_gpu_slab.py: every rank applies its slab with NCCL halo exchange (SlabOperator); rank 0
also applies the unpartitioned mesh and checks the gathered result against it to 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import l3ster_b200 as l3b  # noqa: E402
from l3ster_b200.slab import SlabAssembledOperator, SlabOperator, make_slab  # noqa: E402

U, P = 4, 4
BND = [1, 2, 3, 4, 5, 6]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = l3b.Context(local)
    n, nz = 4, 6
    x1, y1, z1 = np.linspace(0, 1, n + 1), np.linspace(0, 1.1, n + 1), np.linspace(0, 1.3, nz + 1)
    slab = make_slab(x1, y1, z1, P, rank, world)
    op = SlabOperator(ctx, slab, U, "bench_diffusion3d", BND)
    n_lat = n * P + 1, n * P + 1
    key = slab.lattice[:, 0] + n_lat[0] * (slab.lattice[:, 1] + n_lat[1] * slab.lattice[:, 2])
    f = np.random.default_rng(5489).uniform(-1, 1, size=(n_lat[0] * n_lat[1] * (nz * P + 1), U))  # global field by lattice key
    xl = f[key].copy()
    xl[slab.n_owned_nodes:] = np.nan
    x = torch.from_numpy(xl.ravel()).cuda()
    y = torch.zeros_like(x)
    torch.cuda.synchronize()  # tensors are filled on torch's stream, the library runs on its own
    for _ in range(3):  # repeated applies: the halo buffers and events are reused
        op.apply(x, y)
    ctx.synchronize()
    torch.cuda.synchronize()
    # gather (key, y) of the owned nodes on rank 0
    no = slab.n_owned_nodes
    mine = torch.cat([torch.from_numpy(key[:no].astype(np.float64)).cuda()[:, None], y.reshape(-1, U)[:no]], dim=1).contiguous()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([no], dtype=torch.int64, device="cuda"))
    n_max = max(int(s.item()) for s in sizes)
    padded = torch.full((n_max, U + 1), -1.0, dtype=torch.float64, device="cuda")
    padded[:no] = mine
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    bufs = [b[: int(s.item())] for b, s in zip(bufs, sizes)]
    ok = True
    if rank == 0:
        whole = make_slab(x1, y1, z1, P, 0, 1)
        wop = SlabOperator(ctx, whole, U, "bench_diffusion3d", BND)
        wkey = whole.lattice[:, 0] + n_lat[0] * (whole.lattice[:, 1] + n_lat[1] * whole.lattice[:, 2])
        xw = torch.from_numpy(f[wkey].ravel().copy()).cuda()
        yw = torch.zeros_like(xw)
        torch.cuda.synchronize()
        wop.apply(xw, yw)
        ctx.synchronize()
        y_ref = np.zeros_like(f)
        y_ref[wkey] = yw.cpu().numpy().reshape(-1, U)
        y_all = np.full_like(f, np.nan)
        for b in bufs:
            b = b.cpu().numpy()
            y_all[b[:, 0].astype(np.int64)] = b[:, 1:]
        err = np.linalg.norm(y_all - y_ref) / np.linalg.norm(y_ref)
        ok = bool(np.isfinite(err) and err < 1e-12)
        print(f"slab apply over {world} ranks vs single GPU: rel err {err:.2e}, launches/apply {op.launches}")
    # ---- the same apply through l3b_mf_apply with HOST vectors, streamed (interior chunks first, border elements + exchange as the last
    # item) and serial: the owned part must agree with the device-resident apply
    xh, worst = xl.ravel().copy(), 0.0
    y_dev = y.cpu().numpy()
    for mode, chunks, block in ((2, 3, 64), (2, 50, 1), (0, 1, 1)):
        op.sys.set_host_apply(mode, chunks, block)
        yh = np.full_like(xh, np.nan)
        op.sys.apply_raw(xh, yh)
        streamed = op.sys.host_apply_info()["streamed"]
        e = np.linalg.norm(yh[: no * U] - y_dev[: no * U]) / max(np.linalg.norm(y_dev[: no * U]), 1e-300)
        worst = max(worst, e if np.isfinite(e) and streamed == (mode == 2) else 1.0)
    worst_t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(worst_t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"host-vector apply (streamed and serial) vs device-resident apply, worst rank: rel diff {worst_t.item():.2e}")
        ok = ok and bool(worst_t.item() < 1e-13)
    # ---- distributed CG + Jacobi (all-reduced dots, export-summed diag / rhs) against the single-GPU solve
    xs, res, its = op.solve(tol=1e-9, max_iters=2000)
    ctx.synchronize()
    mine = torch.cat([torch.from_numpy(key[:no].astype(np.float64)).cuda()[:, None], xs[: no * U].reshape(-1, U)], dim=1).contiguous()
    padded = torch.full((n_max, U + 1), -1.0, dtype=torch.float64, device="cuda")
    padded[:no] = mine
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    bufs = [b[: int(s.item())] for b, s in zip(bufs, sizes)]
    if rank == 0:
        xw_sol, res_w, its_w = wop.solve(tol=1e-9, max_iters=2000)
        ctx.synchronize()
        x_ref = np.zeros_like(f)
        x_ref[wkey] = xw_sol.cpu().numpy().reshape(-1, U)
        x_all = np.full_like(f, np.nan)
        for b in bufs:
            b = b.cpu().numpy()
            x_all[b[:, 0].astype(np.int64)] = b[:, 1:]
        err = np.linalg.norm(x_all - x_ref) / np.linalg.norm(x_ref)
        print(f"distributed CG: {its} iterations, residual {res:.2e}; single GPU: {its_w} iterations, residual {res_w:.2e}; solution rel diff {err:.2e}")
        ok = ok and bool(np.isfinite(err) and err < 1e-6 and abs(its - its_w) <= 2 and res <= 1e-9)
    # ---- the assembled system over the same slabs (p = 2 keeps the CRS small): distributed CG against the single-GPU solve
    P2 = 2
    slab2 = make_slab(x1, y1, z1, P2, rank, world)
    aop = SlabAssembledOperator(ctx, slab2, U, "bench_diffusion3d", BND)
    xa, res_a, its_a = aop.solve(tol=1e-10, max_iters=4000)
    ctx.synchronize()
    n_lat2 = n * P2 + 1, n * P2 + 1
    key2 = slab2.lattice[:, 0] + n_lat2[0] * (slab2.lattice[:, 1] + n_lat2[1] * slab2.lattice[:, 2])
    no2 = slab2.n_owned_nodes
    sizes2 = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes2, torch.tensor([no2], dtype=torch.int64, device="cuda"))
    n_max2 = max(int(s.item()) for s in sizes2)
    padded = torch.full((n_max2, U + 1), -1.0, dtype=torch.float64, device="cuda")
    padded[:no2] = torch.cat([torch.from_numpy(key2[:no2].astype(np.float64)).cuda()[:, None], xa[: no2 * U].reshape(-1, U)], dim=1)
    bufs2 = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs2, padded)
    bufs2 = [b[: int(s.item())] for b, s in zip(bufs2, sizes2)]
    if rank == 0:
        whole2 = make_slab(x1, y1, z1, P2, 0, 1)
        wa = SlabAssembledOperator(ctx, whole2, U, "bench_diffusion3d", BND)
        xw2, res_w2, its_w2 = wa.solve(tol=1e-10, max_iters=4000)
        # and the matrix-free operator of the same mesh must give the same solution
        wm = SlabOperator(ctx, whole2, U, "bench_diffusion3d", BND)
        xm2, _, its_m2 = wm.solve(tol=1e-10, max_iters=4000)
        ctx.synchronize()
        wkey2 = whole2.lattice[:, 0] + n_lat2[0] * (whole2.lattice[:, 1] + n_lat2[1] * whole2.lattice[:, 2])
        n_all = n_lat2[0] * n_lat2[1] * (nz * P2 + 1)
        x_ref2 = np.zeros((n_all, U))
        x_ref2[wkey2] = xw2.cpu().numpy().reshape(-1, U)
        x_all2 = np.full((n_all, U), np.nan)
        for b in bufs2:
            b = b.cpu().numpy()
            x_all2[b[:, 0].astype(np.int64)] = b[:, 1:]
        err2 = np.linalg.norm(x_all2 - x_ref2) / np.linalg.norm(x_ref2)
        err_mf = float(torch.linalg.norm(xm2 - xw2) / torch.linalg.norm(xw2))
        print(f"assembled distributed CG: {its_a} iterations (single GPU {its_w2}, matrix-free {its_m2}); solution rel diff {err2:.2e}, "
              f"assembled vs matrix-free {err_mf:.2e}")
        ok = ok and bool(np.isfinite(err2) and err2 < 1e-7 and abs(its_a - its_w2) <= 2 and err_mf < 1e-7)
    # ---- the same assembled problem with static condensation (CondensationPolicy::ElementBoundary): the condensed operator lives on the
    # primary nodes [owned | ghost] of every slab; distributed CG, then every rank recovers the interior values of its own elements
    cop = SlabAssembledOperator(ctx, slab2, U, "bench_diffusion3d", BND, condensed=True)
    xc, res_c, its_c = cop.solve(tol=1e-10, max_iters=4000)
    nodal = torch.from_numpy(cop.recover(xc)).cuda()  # over the slab's local nodes
    padded = torch.full((n_max2, U + 1), -1.0, dtype=torch.float64, device="cuda")
    padded[:no2] = torch.cat([torch.from_numpy(key2[:no2].astype(np.float64)).cuda()[:, None], nodal[: no2 * U].reshape(-1, U)], dim=1)
    bufs3 = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs3, padded)
    bufs3 = [b[: int(s.item())] for b, s in zip(bufs3, sizes2)]
    if rank == 0:
        x_all3 = np.full((n_all, U), np.nan)
        for b in bufs3:
            b = b.cpu().numpy()
            x_all3[b[:, 0].astype(np.int64)] = b[:, 1:]
        err3 = np.linalg.norm(x_all3 - x_ref2) / np.linalg.norm(x_ref2)
        print(f"condensed distributed CG: {its_c} iterations on {cop.n_owned_dofs} of {aop.n_owned_dofs} owned dofs (uncondensed: {its_a} iterations); "
              f"recovered solution vs uncondensed rel diff {err3:.2e}")
        ok = ok and bool(np.isfinite(err3) and err3 < 1e-7 and res_c <= 1e-10 and its_c <= its_a)
    # ---- fewer element layers than ranks (tests/EmptyPartitionTest.cpp:10-49, both condensation policies): the empty ranks take part in
    # the reductions with zero dofs; the solution is the one-rank solution
    z_one = z1[:2]
    for cond in (False, True):
        eslab = make_slab(x1, y1, z_one, P2, rank, world)
        eop = SlabAssembledOperator(ctx, eslab, U, "bench_diffusion3d", BND, condensed=cond)
        xe, res_e, its_e = eop.solve(tol=1e-10, max_iters=2000)
        mine = float(torch.sum(xe[: eop.n_owned_dofs] ** 2).item()) if eop.n_owned_dofs else 0.0
        tot = torch.tensor([mine, float(eslab.n_elems)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        if rank == 0:
            wslab = make_slab(x1, y1, z_one, P2, 0, 1)
            wop_e = SlabAssembledOperator(ctx, wslab, U, "bench_diffusion3d", BND, condensed=cond)
            xw_e, _, its_we = wop_e.solve(tol=1e-10, max_iters=2000)
            ref = float(torch.sum(xw_e ** 2).item())
            print(f"one element layer over {world} ranks (condensed={cond}): {its_e} iterations (one rank {its_we}), |x|^2 {tot[0].item():.12e} vs {ref:.12e}")
            ok = ok and bool(abs(tot[0].item() - ref) <= 1e-8 * ref and abs(its_e - its_we) <= 1 and int(tot[1].item()) == wslab.n_elems)
    # the matrix-free operator with an empty rank
    eslab = make_slab(x1, y1, z_one, P2, rank, world)
    emf = SlabOperator(ctx, eslab, U, "bench_diffusion3d", BND)
    xm, res_m, its_m = emf.solve(tol=1e-10, max_iters=2000)
    mine = float(torch.sum(xm[: emf.n_owned_dofs] ** 2).item()) if emf.n_owned_dofs else 0.0
    tot = torch.tensor([mine], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        wmf = SlabOperator(ctx, make_slab(x1, y1, z_one, P2, 0, 1), U, "bench_diffusion3d", BND)
        xw_m, _, its_wm = wmf.solve(tol=1e-10, max_iters=2000)
        ref = float(torch.sum(xw_m ** 2).item())
        print(f"matrix-free, one element layer over {world} ranks: {its_m} iterations (one rank {its_wm}), |x|^2 {tot[0].item():.12e} vs {ref:.12e}")
        ok = ok and bool(abs(tot[0].item() - ref) <= 1e-8 * ref and abs(its_m - its_wm) <= 1)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("SLAB_APPLY_OK")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
