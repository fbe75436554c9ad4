"""Worker of the partition-import tests (torchrun, one rank per GPU; also runs on one rank): a hex p=4 cube with distorted elements is
cut by a deliberately ragged, non-slab `epart`; every rank builds its views with l3b_partition_* and

  * applies the matrix-free operator through the library's halo (general neighbour lists, NCCL) and solves with CG — against the
    unpartitioned operator on rank 0 and against the oracle;
  * assembles into the row-complete owner graph, export-adds the shared rows (l3b_asm_export_shared_rows), applies the Dirichlet
    conditions — graph bit-exact and values to 1e-12 against the ORACLE's single-rank matrix on the renumbered mesh
    (algsys/SparsityGraph.hpp:83-278, AssembledSystem.hpp:384-389), then solves.
Prints PARTITION_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import l3ster_b200 as l3b  # noqa: E402
from common import distort  # noqa: E402
from l3ster_b200.partition import Partition, bisection_epart  # noqa: E402
from l3ster_b200.slab import SlabAssembledOperator, SlabOperator  # noqa: E402

P, U, N = 4, 4, 3
BND = [1, 2, 3, 4, 5, 6]
TOL = 1e-12


def ragged_epart(host, n_parts):
    rng = np.random.default_rng(2024)
    ep = bisection_epart(host.verts.mean(axis=1), n_parts)
    flip = rng.random(host.n_elems) < 0.2
    ep[flip] = rng.integers(0, n_parts, size=int(flip.sum()))
    return ep.astype(np.int32)


def seeded(gids):
    g = gids.astype(np.float64)
    return (np.sin(0.37 * g + 1.0)[:, None] * (1.0 + 0.25 * np.arange(U))[None, :] + 0.1 * np.arange(U)[None, :]).ravel()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = l3b.Context(local)
    host = l3b.make_cube_mesh(np.linspace(0.0, 1.0, N + 1), np.linspace(0.0, 1.2, N + 1), np.linspace(0.0, 0.9, N + 1), order=P)
    verts = distort(host.verts)
    ep = ragged_epart(host, world)
    part = Partition(3, P, host.n_nodes, host.nodes, verts, host.side_boundaries, world, ep)
    checks = {}

    # ---- matrix-free: halo'd apply + CG
    view = part.rank_view(rank, False)
    op = SlabOperator(ctx, view, U, "bench_diffusion3d", BND)
    no = view.n_owned_nodes * U
    x = torch.from_numpy(seeded(view.gids)).cuda()
    x[no:] = float("nan")  # ghosts must come from the Import
    y = torch.zeros_like(x)
    torch.cuda.synchronize()
    op.apply(x, y)
    ctx.synchronize()
    xs, res, its = op.solve(tol=1e-9, max_iters=3000)
    ctx.synchronize()
    pieces = [None] * world
    dist.all_gather_object(pieces, (int(view.first_gid), y[:no].cpu().numpy(), xs[:no].cpu().numpy(), its))
    # ---- assembled: row-complete owner matrix
    viewx = part.rank_view(rank, True)
    aop = SlabAssembledOperator(ctx, viewx, U, "bench_diffusion3d", BND)
    vals, rhs = aop.sys.download()
    row_ptr, col_ind = aop.sys.graph()
    nox = viewx.n_owned_nodes * U
    gdof = (viewx.gids[:, None] * U + np.arange(U)[None, :]).ravel()
    rows = np.repeat(np.arange(nox), np.diff(row_ptr[:nox + 1]))
    trip = (gdof[rows], gdof[col_ind[:row_ptr[nox]]], vals[:row_ptr[nox]], rhs[:nox, 0])
    ghost_tail_zero = bool(np.all(vals[row_ptr[nox]:] == 0.0))
    xa, res_a, its_a = aop.solve(tol=1e-9, max_iters=3000)
    ctx.synchronize()
    apieces = [None] * world
    dist.all_gather_object(apieces, (int(viewx.first_gid), trip, xa[:nox].cpu().numpy(), its_a, ghost_tail_zero))

    ok = True
    if rank == 0:
        import scipy.sparse as sp

        from oracle import Oracle

        n_dofs = host.n_nodes * U
        y_all, x_all, xa_all = np.zeros(n_dofs), np.zeros(n_dofs), np.zeros(n_dofs)
        for first, yp, xp, _ in pieces:
            y_all[first * U:first * U + len(yp)] = yp
            x_all[first * U:first * U + len(xp)] = xp
        # the unpartitioned operator (product) and the oracle, on the RENUMBERED mesh
        gnodes = part.new_id[host.nodes.astype(np.int64)].astype(np.uint32)
        whole = Partition(3, P, host.n_nodes, gnodes, verts, host.side_boundaries, 1, np.zeros(host.n_elems, dtype=np.int32)).rank_view(0)
        assert np.array_equal(whole.gids, np.arange(host.n_nodes))
        wop = SlabOperator(ctx, whole, U, "bench_diffusion3d", BND)
        xw = torch.from_numpy(seeded(whole.gids)).cuda()
        yw = torch.zeros_like(xw)
        torch.cuda.synchronize()
        wop.apply(xw, yw)
        ctx.synchronize()
        xsw, res_w, its_w = wop.solve(tol=1e-9, max_iters=3000)
        ctx.synchronize()
        yw, xsw = yw.cpu().numpy(), xsw.cpu().numpy()
        orc = Oracle()
        om = orc.mesh_from_nodes(3, P, host.n_nodes, gnodes, verts)
        mask = np.zeros(n_dofs, dtype=np.uint8)
        dir_nodes = part.global_boundary_nodes(BND)
        mask[dir_nodes * U] = 1
        omf = om.matrix_free_system(U, 1, mask, None)
        omf.add_kernel("bench_diffusion3d")
        yo = omf.apply(seeded(whole.gids).reshape(-1, 1))[:, 0]
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
        checks["mf_apply_vs_one_rank"] = rel(y_all, yw) < TOL
        checks["mf_apply_vs_oracle"] = rel(y_all, yo) < TOL
        checks["mf_cg_same_iterations"] = abs(its - its_w) <= 2 and all(p[3] == its for p in pieces)
        checks["mf_cg_same_solution"] = rel(x_all, xsw) < 1e-7
        # assembled: oracle single-rank matrix with the same Dirichlet conditions
        oa = om.assembled_system(U)
        oa.assemble("bench_diffusion3d")
        oa.apply_dirichlet((dir_nodes * U).astype(np.int32), np.zeros(len(dir_nodes)))
        ovals, orhs = oa.get()
        A_o = sp.csr_matrix((ovals, oa.col_ind, oa.row_ptr), shape=(n_dofs, n_dofs))
        r_all = np.concatenate([t[1][0] for t in apieces])
        c_all = np.concatenate([t[1][1] for t in apieces])
        v_all = np.concatenate([t[1][2] for t in apieces])
        rhs_all = np.zeros(n_dofs)
        for first, t, xp, _, _ in apieces:
            rhs_all[first * U:first * U + len(t[3])] = t[3]
            xa_all[first * U:first * U + len(xp)] = xp
        A_p = sp.csr_matrix((v_all, (r_all, c_all)), shape=(n_dofs, n_dofs))
        A_p.sort_indices()
        A_o.sort_indices()
        checks["asm_graph_bit_exact"] = bool(np.array_equal(A_p.indptr, A_o.indptr) and np.array_equal(A_p.indices, A_o.indices))
        checks["asm_values"] = checks["asm_graph_bit_exact"] and rel(A_p.data, A_o.data) < TOL
        checks["asm_rhs"] = rel(rhs_all, orhs[:, 0]) < TOL
        checks["asm_ghost_rows_spent"] = all(t[4] for t in apieces) if world > 1 else True
        checks["asm_solution_matches_matrix_free"] = rel(xa_all, xsw) < 1e-6
        checks["asm_same_iterations_on_all_ranks"] = all(t[3] == its_a for t in apieces)
        print(f"partition over {world} ranks: {[int((ep == r).sum()) for r in range(world)]} elements, owned nodes {np.diff(part.dist).tolist()}; "
              f"CG {its} iterations ({its_w} on one rank), assembled CG {its_a}")
        print(checks)
        ok = all(checks.values())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("PARTITION_OK")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
