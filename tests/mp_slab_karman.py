"""torchrun worker of tests/test_gpu_slab.py: BASELINE configs[3] on more than one GPU — the steady Navier-Stokes kernel of
examples/07-karman-2D (U = 4, previous velocity as external fields, AssemblyOptions{1, 1}) with its outlet boundary kernel, assembled on
y-strips of a channel mesh, Dirichlet u, v on inlet and walls, restarted GMRES + Jacobi with NCCL halo exchange and all-reduced
orthogonalisation; rank 0 solves the unpartitioned problem and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import l3ster_b200 as l3b  # noqa: E402
from l3ster_b200.slab import SlabAssembledOperator, make_slab  # noqa: E402

U, P = 4, 4
BOTTOM, TOP, LEFT, RIGHT = 1, 2, 3, 4  # mesh/primitives/SquareMesh.hpp:14-76
NX, NY = 6, 4
TOL, RESTART = 1e-9, 250  # solve/SolverInterface.hpp:26-37: restart length 250 is the reference default


def node_coords(slab, xs, ys):
    """physical coordinates of the local nodes of an axis-aligned strip from their lattice coordinates (Gauss-Lobatto points inside
    every element)"""
    gll = l3b.tables_gll(P + 1)
    out = np.zeros((slab.n_local_nodes, 2))
    for d, ax in enumerate((xs, ys)):
        lat = slab.lattice[:, d]
        e = np.minimum(lat // P, len(ax) - 2)
        i = lat - e * P
        out[:, d] = ax[e] + (ax[e + 1] - ax[e]) * (1.0 + gll[i]) / 2.0
    return out


def problem(ctx, slab, xs, ys):
    xy = node_coords(slab, xs, ys)
    x, y = xy[:, 0], xy[:, 1]
    h = ys[-1]
    # previous iterate: a developing channel profile with a cross-flow perturbation (any smooth field will do)
    u0 = 4.0 * y * (h - y) / h**2 * (1.0 + 0.1 * np.sin(x))
    v0 = 0.05 * np.sin(np.pi * y / h) * np.cos(0.7 * x)
    fdata = np.stack([u0, v0])
    opts = l3b.AssemblyOptions(value_order=1, derivative_order=1)
    kernels = [dict(name="karman_steady", asm_opts=opts, field_inds=[0, 1]),
               dict(name="karman_outlet", boundary_ids=[RIGHT], dof_inds=[0, 1, 3], asm_opts=opts)]
    walls, inlet = slab.dirichlet_nodes([BOTTOM, TOP]), slab.dirichlet_nodes([LEFT])
    nodes = np.union1d(walls, inlet)
    uin = np.where(np.isin(nodes, inlet), 4.0 * y[nodes] * (h - y[nodes]) / h**2, 0.0)
    uin[np.isin(nodes, walls)] = 0.0
    dofs = np.concatenate([nodes * U, nodes * U + 1])
    vals = np.concatenate([uin, np.zeros(len(nodes))])
    order = np.argsort(dofs)
    return SlabAssembledOperator(ctx, slab, U, kernels, field_data=fdata, dirichlet=(dofs[order], vals[order]))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = l3b.Context(local)
    xs, ys = np.linspace(0.0, 3.0, NX + 1), np.linspace(0.0, 1.0, NY + 1)
    slab = make_slab(xs, ys, None, P, rank, world)
    op = problem(ctx, slab, xs, ys)
    x, res, its = op.solve(tol=TOL, max_iters=4000, gmres=True, restart_length=RESTART)
    ctx.synchronize()
    stride = NX * P + 1
    key = slab.lattice[:, 0] + stride * slab.lattice[:, 1]
    no = slab.n_owned_nodes
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([no], dtype=torch.int64, device="cuda"))
    n_max = max(int(s.item()) for s in sizes)
    padded = torch.full((n_max, U + 1), -1.0, dtype=torch.float64, device="cuda")
    padded[:no] = torch.cat([torch.from_numpy(key[:no].astype(np.float64)).cuda()[:, None], x[: no * U].reshape(-1, U)], dim=1)
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    ok = True
    if rank == 0:
        whole = make_slab(xs, ys, None, P, 0, 1)
        wop = problem(ctx, whole, xs, ys)
        # one rank: the callback-driven driver (l3b_gmres_device with this operator's apply) against the library's own (host vectors)
        xw, res_w, its_w = wop.solve(tol=TOL, max_iters=4000, gmres=True, restart_length=RESTART, callbacks=True)
        xd, res_d, its_d = wop.sys.solve_gmres(tol=TOL, restart_length=RESTART, max_iters=4000)
        ctx.synchronize()
        wkey = whole.lattice[:, 0] + stride * whole.lattice[:, 1]
        n_all = stride * (NY * P + 1)
        x_ref = np.zeros((n_all, U))
        x_ref[wkey] = xw.cpu().numpy().reshape(-1, U)
        x_all = np.full((n_all, U), np.nan)
        for b, s in zip(bufs, sizes):
            b = b[: int(s.item())].cpu().numpy()
            x_all[b[:, 0].astype(np.int64)] = b[:, 1:]
        err = np.linalg.norm(x_all - x_ref) / np.linalg.norm(x_ref)
        err_d = np.linalg.norm(xd - xw.cpu().numpy()) / np.linalg.norm(xd)
        print(f"karman steady, assembled + GMRES({RESTART}) over {world} ranks: {its} iterations, residual {res:.2e}; one rank: {its_w} iterations "
              f"(library driver {its_d}); solution rel diff {err:.2e}, callback vs library driver {err_d:.2e}")
        checks = dict(finite=bool(np.isfinite(err)), same_solution=err < 1e-6, same_iterations=abs(its - its_w) <= 3, converged=res <= TOL,
                      drivers_agree=err_d < 1e-7, nontrivial=np.abs(x_ref).max() > 0.5)
        print(checks)
        ok = all(checks.values())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("SLAB_KARMAN_OK")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
