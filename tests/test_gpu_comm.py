"""The library's communicator and halo engine (l3b_comm_*, l3b_halo_*: comm/MpiComm.hpp, comm/ImportExport.hpp:29-470) on ONE GPU:
a single-rank NCCL communicator whose rank is its own neighbour exercises pack -> ncclSend/ncclRecv group -> in-place receive and the
Export's unpack-add; the systems' halo hooks (l3b_mf_set_halo / l3b_asm_set_halo) are driven with that loop-back halo and checked
against the same operations done by hand. The multi-rank runs are tests/mp_slab_*.py (two GPUs) and bench.py's `parity_vs_n1`."""
import numpy as np
import pytest

import l3ster_b200 as l3b
from common import PairedMesh, default_dists, oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _torch():
    import torch

    return torch


def test_single_rank_communicator_and_allreduce():
    torch = _torch()
    ctx = l3b.Context(0)
    comm = l3b.Comm(ctx)
    assert comm.rank == 0 and comm.world == 1
    s = torch.tensor([1.5, -2.0, 3.25], dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    comm.allreduce_sum(s.data_ptr(), 3)
    ctx.synchronize()
    assert s.tolist() == [1.5, -2.0, 3.25]


@pytest.mark.parametrize("n_cols", [1, 3])
def test_loopback_halo_import_and_export(n_cols):
    """rank 0 is its own neighbour, twice (two 'neighbours' with different index lists): the Import must copy the listed owned dofs
    into the ghost ranges, the Export must add the ghost ranges into the listed owned dofs (an owned dof listed for both neighbours
    receives both contributions)"""
    torch = _torch()
    ctx = l3b.Context(0)
    comm = l3b.Comm(ctx)
    rng = np.random.default_rng(7)
    n_owned, sizes = 1000, (37, 211)
    lists = [rng.choice(n_owned, size=k, replace=False).astype(np.int32) for k in sizes]
    lists[1][:5] = lists[0][:5]  # shared by both neighbours
    n_ghost = sum(sizes) + 13  # a tail of ghosts nobody sends (owned by a rank that is not a neighbour here)
    halo = l3b.DeviceHalo(comm, n_owned, n_ghost, [(0, lists[0]), (0, lists[1])], [(0, 0, sizes[0]), (0, sizes[0], sizes[1])])
    ld = n_owned + n_ghost
    xh = rng.uniform(-1, 1, size=(n_cols, ld))
    x = torch.from_numpy(xh.copy()).cuda()
    torch.cuda.synchronize()
    halo.import_(x.data_ptr(), n_cols)
    ctx.synchronize()
    want = xh.copy()
    want[:, n_owned:n_owned + sizes[0]] = xh[:, lists[0]]
    want[:, n_owned + sizes[0]:n_owned + sum(sizes)] = xh[:, lists[1]]
    assert np.array_equal(x.cpu().numpy(), want)
    # Export
    yh = rng.uniform(-1, 1, size=(n_cols, ld))
    y = torch.from_numpy(yh.copy()).cuda()
    torch.cuda.synchronize()
    halo.export_add(y.data_ptr(), n_cols)
    ctx.synchronize()
    want = yh.copy()
    for c in range(n_cols):
        np.add.at(want[c], lists[0], yh[c, n_owned:n_owned + sizes[0]])
        np.add.at(want[c], lists[1], yh[c, n_owned + sizes[0]:n_owned + sum(sizes)])
    got = y.cpu().numpy()
    assert np.array_equal(got[:, n_owned:], yh[:, n_owned:])  # the ghost block is only read
    assert np.abs(got - want).max() < 1e-15


def test_halo_rejects_inconsistent_descriptions():
    ctx = l3b.Context(0)
    comm = l3b.Comm(ctx)
    with pytest.raises(l3b.L3BError):
        l3b.DeviceHalo(comm, 10, 4, [(0, np.array([11], dtype=np.int32))], [])  # packed index is not an owned dof
    with pytest.raises(l3b.L3BError):
        l3b.DeviceHalo(comm, 10, 4, [], [(3, 0, 4)])  # neighbour rank outside the communicator
    with pytest.raises(l3b.L3BError):
        l3b.DeviceHalo(comm, 10, 4, [], [(0, 0, 5)])  # range larger than the ghost block


def test_matrix_free_apply_with_a_halo_is_one_library_call():
    """l3b_mf_set_halo: the last nodes of a mesh are declared ghost copies of some owned nodes of the same rank (loop-back halo). The
    halo'd apply (Import, border elements, Export, interior elements, unpack-add, Dirichlet rows in ONE call) and the export-summed
    diag / rhs of endAssembly must equal: import by hand, the plain system over all elements, export-add by hand."""
    torch = _torch()
    ctx = l3b.Context(0)
    U = 4
    pm = PairedMesh(3, default_dists(3, 3), 2)
    host = pm.host
    n_nodes, n_ghost_nodes = host.n_nodes, 40
    n_owned_nodes = n_nodes - n_ghost_nodes
    no = n_owned_nodes * U
    touches = (host.nodes >= n_owned_nodes).any(axis=1)
    order = np.argsort(~touches, kind="stable")  # elements touching ghost nodes first
    n_border = int(touches.sum())
    assert 0 < n_border < host.n_elems
    mesh_plain = l3b.Mesh(ctx, 3, 2, pm.verts, host.nodes, host.side_boundaries, n_nodes, n_nodes)
    mesh_rank = l3b.Mesh(ctx, 3, 2, pm.verts[order], host.nodes[order], host.side_boundaries[order], n_nodes, n_owned_nodes)
    rng = np.random.default_rng(3)
    src_nodes = rng.choice(n_owned_nodes, size=n_ghost_nodes, replace=False)
    src_dofs = (src_nodes[:, None] * U + np.arange(U)[None, :]).ravel().astype(np.int32)
    free_nodes = np.setdiff1d(np.arange(n_owned_nodes), src_nodes)
    mask = np.zeros(n_nodes * U, dtype=np.uint8)
    mask[free_nodes[::7] * U] = 1  # Dirichlet dofs on owned nodes that are not halo sources
    vals = np.zeros(n_nodes * U)
    vals[mask == 1] = rng.uniform(-1, 1, size=int(mask.sum()))
    comm = l3b.Comm(ctx)
    halo = l3b.DeviceHalo(comm, no, n_ghost_nodes * U, [(0, src_dofs)], [(0, 0, n_ghost_nodes * U)])

    def system(mesh, with_halo):
        s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, vals)
        s.assembleProblem("bench_diffusion3d")
        if with_halo:
            s.set_halo(halo, n_border)
        s.endAssembly()
        return s

    plain, ranked = system(mesh_plain, False), system(mesh_rank, True)
    x = rng.uniform(-1, 1, size=n_nodes * U)
    xi = x.copy()
    xi[no:] = x[src_dofs]  # Import by hand
    y_plain = plain.apply(xi.reshape(-1, 1))[:, 0]
    want = y_plain[:no].copy()
    np.add.at(want, src_dofs, y_plain[no:])  # Export-sum by hand
    xd = torch.from_numpy(x.copy()).cuda()
    yd = torch.full_like(xd, 5.0)
    torch.cuda.synchronize()
    ranked.apply_device(xd.data_ptr(), yd.data_ptr())
    ctx.synchronize()
    assert np.array_equal(xd.cpu().numpy(), xi)  # the Import filled the ghost block of x
    assert rel_err(yd.cpu().numpy()[:no], want) < TOL
    # the energy variant (CG's p.Ap): x^T A x over this rank's elements and owned Dirichlet dofs
    e = torch.zeros(1, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ranked.apply_device(xd.data_ptr(), yd.data_ptr(), energy_ptr=e.data_ptr())
    ctx.synchronize()
    assert rel_err(yd.cpu().numpy()[:no], want) < TOL
    assert abs(e.item() - float(xi[:no] @ want)) < 1e-11 * abs(e.item())
    # diag and rhs: export-summed inside l3b_mf_end_assembly, then the Dirichlet dofs
    d_plain, r_plain = plain.download()
    d_rank, r_rank = ranked.download()
    dw, rw = d_plain[:no].copy(), r_plain[:no, 0].copy()
    np.add.at(dw, src_dofs, d_plain[no:])
    np.add.at(rw, src_dofs, r_plain[no:, 0])
    assert rel_err(d_rank[:no], dw) < TOL and rel_err(r_rank[:no, 0], rw) < TOL
    # CG over the "ranks": converges on the owned dofs to the solution of the condensed (identified) operator
    xs, res, its = ranked.solve(tol=1e-10)
    assert res <= 1e-10 and its > 3
    xsd = torch.from_numpy(np.concatenate([xs[:no], xs[src_dofs]])).cuda()
    torch.cuda.synchronize()
    ranked.apply_device(xsd.data_ptr(), yd.data_ptr())
    ctx.synchronize()
    assert np.linalg.norm(yd.cpu().numpy()[:no] - rw) < 1e-8 * max(1.0, np.linalg.norm(rw))
    assert np.array_equal(xs[no:], xs[src_dofs])  # ghost copies refreshed on return


def test_phased_apply_runs_boundary_kernels_exactly_once():
    """a rank without border elements issues ELEMENTS over [0, 0) and then over [0, n): the boundary kernels (the side list is not split)
    must run once (they ran twice); with the explicit L3B_APPLY_BOUNDARY bit they run in the call that carries it"""
    torch = _torch()
    ctx = l3b.Context(0)
    U = 3
    pm = PairedMesh(2, default_dists(2, 3), 2)
    mesh = pm.upload(ctx)
    s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, None, None)
    s.assembleProblem("diffusion_kernel_2D_r1")
    s.assembleProblem("adiabatic_bc_2D", boundary_ids=[0, 1, 2, 3])
    s.endAssembly()
    x = np.random.default_rng(11).uniform(-1, 1, size=pm.n_nodes * U)
    y_ref = s.apply(x.reshape(-1, 1))[:, 0]
    osys = pm.orc.matrix_free_system(U, 1, None, None)
    osys.add_kernel("diffusion_kernel_2D")
    osys.add_kernel("adiabatic_bc_2D", boundary_ids=[0, 1, 2, 3])
    assert rel_err(y_ref, osys.apply(x.reshape(-1, 1))[:, 0]) < TOL
    xd = torch.from_numpy(x).cuda()
    n_el = pm.host.n_elems

    def run(plan, init=True):
        yd = torch.full_like(xd, 3.0) if init else torch.zeros_like(xd)
        torch.cuda.synchronize()
        if init:
            s.apply_phase_device(xd.data_ptr(), yd.data_ptr(), l3b.APPLY_INIT, 0, 0)
        for ph, b, e in plan:
            s.apply_phase_device(xd.data_ptr(), yd.data_ptr(), ph, b, e)
        ctx.synchronize()
        return yd.cpu().numpy()

    E, B, F = l3b.APPLY_ELEMENTS, l3b.APPLY_BOUNDARY, l3b.APPLY_FINISH
    assert rel_err(run([(E, 0, 0), (E, 0, n_el), (F, 0, 0)]), y_ref) < TOL  # no border elements: the ADVICE case
    assert rel_err(run([(E, 2, n_el), (E, 0, 2), (F, 0, 0)]), y_ref) < TOL  # interior first, then the range holding element 0
    assert rel_err(run([(E | B, 2, n_el), (E, 0, 0), (F, 0, 0)]) + run([(E, 0, 2)], init=False) - run([(B, 0, 0)], init=False), y_ref) < TOL


def test_solvers_take_the_initial_guess():
    """Belos starts from the system's persistent solution (AssembledSystem.hpp:113-135): a second solve from the converged solution needs
    no iteration, a solve from a random guess reaches the same solution (x was zeroed)"""
    ctx = l3b.Context(0)
    U = 4
    pm = PairedMesh(3, default_dists(3, 2), 2)
    mesh = pm.upload(ctx)
    mask = np.zeros(pm.n_nodes * U, dtype=np.uint8)
    mask[pm.host.boundary_nodes([1, 2, 3, 4, 5, 6]) * U] = 1
    mf = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    mf.assembleProblem("bench_diffusion3d")
    mf.endAssembly()
    x0, res0, it0 = mf.solve(tol=1e-10)
    assert it0 > 5 and res0 <= 1e-10
    x1, res1, it1 = mf.solve(tol=1e-8, x0=x0)
    assert it1 == 0 and res1 <= 1e-8 and np.array_equal(x1, x0)
    guess = x0 + 0.1 * np.random.default_rng(1).uniform(-1, 1, size=x0.shape) * (1 - mask)
    x2, res2, it2 = mf.solve(tol=1e-10, x0=guess)
    assert 0 < it2 and rel_err(x2, x0) < 1e-7
    # GMRES tests the PRECONDITIONED residual (Belos' implicit test), so converge it on its own criterion first
    xg, resg, itg = mf.solve_gmres(tol=1e-9, x0=x0)
    assert resg <= 1e-9 and rel_err(xg, x0) < 1e-7
    xg2, _, itg2 = mf.solve_gmres(tol=1e-7, x0=xg)
    assert itg2 == 0 and np.array_equal(xg2, xg)
    asm = l3b.AssembledSystem(ctx, mesh, U)
    asm.beginAssembly()
    asm.assembleProblem("bench_diffusion3d")
    dofs = np.nonzero(mask)[0].astype(np.int32)
    asm.endAssembly(dofs, np.zeros(len(dofs)))
    xa, _, ita = asm.solve(tol=1e-10)
    assert ita > 5 and rel_err(xa, x0) < 1e-6
    xb, _, itb = asm.solve(tol=1e-8, x0=xa)
    assert itb == 0 and np.array_equal(xb, xa)
    xc, resc, itc = asm.solve_gmres(tol=1e-9, x0=xa)
    xc2, _, itc2 = asm.solve_gmres(tol=1e-7, x0=xc)
    assert resc <= 1e-9 and itc2 == 0 and rel_err(xc, xa) < 1e-6
