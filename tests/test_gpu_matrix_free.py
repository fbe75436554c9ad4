"""GPU parity: matrix-free operator (sum-factorised and local-element paths, boundary kernels, Dirichlet handling,
diag/rhs init, CG) against the CPU oracle on identical seeded inputs. All calls go through the C ABI.

Tolerance: 1e-12 relative (Frobenius), the fp64 bar stated in BASELINE.json's north_star; fp64 atomics make neither side
bit-reproducible.
"""
import numpy as np
import pytest

import l3ster_b200 as l3b
from common import PairedMesh, default_dists, oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    return l3b.Context(0)


def _fields(n_fields, n_nodes, seed):
    return np.random.default_rng(seed).uniform(-1, 1, size=(n_fields, n_nodes))


def _dirichlet(pm, U, boundary_ids, dof_inds, n_rhs, seed):
    mask = np.zeros(pm.n_nodes * U, dtype=np.uint8)
    nodes = pm.host.boundary_nodes(boundary_ids)
    for d in dof_inds:
        mask[nodes * U + d] = 1
    vals = np.random.default_rng(seed).uniform(-1, 1, size=(pm.n_nodes * U, n_rhs)) * mask[:, None]
    return mask, vals


CASES = [
    # kernel (product name), oracle name, dim, n, order, opts, n_rhs
    ("bench_diffusion3d", "bench_diffusion3d", 3, 2, 4, l3b.AssemblyOptions(), 1),
    ("bench_diffusion3d", "bench_diffusion3d", 3, 3, 2, l3b.AssemblyOptions(), 1),
    ("bench_diffusion3d", "bench_diffusion3d", 3, 2, 1, l3b.AssemblyOptions(), 1),
    ("bench_diffusion3d", "bench_diffusion3d", 3, 2, 6, l3b.AssemblyOptions(), 1),
    ("diffusion_kernel_3D", "diffusion_kernel_3D", 3, 2, 3, l3b.AssemblyOptions(value_order=2), 3),
    ("diffusion_kernel_3D_var", "diffusion_kernel_3D_var", 3, 2, 3, l3b.AssemblyOptions(value_order=2), 2),
    ("dense_probe_3D", "dense_probe_3D", 3, 2, 2, l3b.AssemblyOptions(), 1),
    ("dense_probe_3D", "dense_probe_3D", 3, 2, 3, l3b.AssemblyOptions(value_order=1, derivative_order=1), 1),
    ("diffusion_kernel_2D", "diffusion_kernel_2D", 2, 3, 4, l3b.AssemblyOptions(value_order=2), 2),
    ("diffusion_kernel_2D_var", "diffusion_kernel_2D_var", 2, 3, 3, l3b.AssemblyOptions(), 2),
    ("dense_probe_2D", "dense_probe_2D", 2, 4, 2, l3b.AssemblyOptions(value_order=1, derivative_order=1), 1),
    ("example02_domain", "example02_domain", 2, 4, 4, l3b.AssemblyOptions(), 1),
    # benchmarks/LocalOperatorEvaluationBenchmarks.cpp / LocalAssemblyBenchmarks.cpp:41-87: NS3D, U = 7, E = 8, n_fields = 7
    ("ns3d_kernel", "ns3d_kernel", 3, 2, 2, l3b.AssemblyOptions(value_order=1, derivative_order=1), 1),
    ("ns3d_kernel", "ns3d_kernel", 3, 2, 3, l3b.AssemblyOptions(), 1),
    ("ns3d_kernel", "ns3d_kernel", 3, 1, 4, l3b.AssemblyOptions(), 1),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-d{c[2]}-n{c[3]}-p{c[4]}-vo{c[5].value_order}do{c[5].derivative_order}")
@pytest.mark.parametrize("strategy", [0, 1], ids=["sumfact", "local_element"])
@pytest.mark.parametrize("with_bc", [False, True], ids=["nobc", "dirichlet"])
def test_operator_apply_matches_oracle(ctx, case, strategy, with_bc):
    kname, oname, dim, n, p, opts, n_rhs = case
    opts = l3b.AssemblyOptions(opts.value_order, opts.derivative_order, strategy)
    info = l3b.kernel_info(kname)
    U, NF = info["n_unknowns"], info["n_fields"]
    pm = PairedMesh(dim, default_dists(dim, n), p)
    mesh = pm.upload(ctx)
    time = 0.37
    fdata = _fields(NF, pm.n_nodes, 7) if NF else None
    mask, dvals = _dirichlet(pm, U, [1, 2 * dim], [0, U - 1], n_rhs, 3) if with_bc else (None, None)

    sys_g = l3b.MatrixFreeSystem(ctx, mesh, U, n_rhs, mask, dvals)
    fields = ctx.upload_fields(fdata) if NF else None
    sys_g.assembleProblem(kname, fields=fields, asm_opts=opts, time=time)
    sys_g.endAssembly()
    sys_o = pm.orc.matrix_free_system(U, n_rhs, mask, dvals)
    sys_o.add_kernel(oname, opts.value_order, opts.derivative_order, strategy, time, fdata)
    diag_o, rhs_o = sys_o.init(n_threads=4)

    # init: diagonal + rhs with Dirichlet lifting (MatrixFreeSystem::endAssembly)
    diag_g, rhs_g = sys_g.download()
    assert rel_err(diag_g, diag_o) < TOL
    assert np.linalg.norm(rhs_g - rhs_o) <= TOL * max(np.linalg.norm(rhs_o), np.linalg.norm(diag_o))

    rng = np.random.default_rng(11)
    for n_cols in sorted({n_rhs, 1}):
        x = rng.uniform(-1, 1, size=(pm.n_nodes * U, n_cols))
        y0 = rng.uniform(-1, 1, size=x.shape)
        for alpha, beta in ((1.0, 0.0), (-0.7, 0.4)):
            y_g = sys_g.apply(x, y0, alpha, beta)
            y_o = sys_o.apply(x, y0, alpha, beta, n_threads=4)
            assert rel_err(y_g, y_o) < TOL, (n_cols, alpha, beta)


def test_hex_sumfact_passes_z_zero_like_the_reference(ctx):
    """SumFactorization.hpp:732 hands the kernel point.space = (x, y, 0) on hexes; the non-SF path passes the true point.
    dense_probe_3D depends on x and y only, so both paths must agree; the oracle replicates the quirk and pins it."""
    pm = PairedMesh(3, default_dists(3, 2), 2)
    mesh = pm.upload(ctx)
    f = _fields(2, pm.n_nodes, 1)
    x = np.random.default_rng(2).uniform(-1, 1, size=(pm.n_nodes * 3, 1))
    ys = []
    for strategy in (0, 1):
        s = l3b.MatrixFreeSystem(ctx, mesh, 3)
        s.assembleProblem("dense_probe_3D", fields=ctx.upload_fields(f), asm_opts=l3b.AssemblyOptions(eval_strategy=strategy))
        s.endAssembly()
        ys.append(s.apply(x))
    assert rel_err(ys[0], ys[1]) < TOL


def test_boundary_kernel_and_cg_diffusion2d(ctx):
    """tests/Diffusion2D.hpp:23-120 (matrix-free variants Diffusion2DMF / Diffusion2DMFSF): 4x4 quads p=2, U=3, Dirichlet
    T = x on left/right, adiabatic boundary kernel on bottom/top, CG(1e-10) + Jacobi; exact solution T = x, q = (1, 0)."""
    node_dist = np.linspace(0.0, 1.0, 5)
    host = l3b.make_square_mesh(node_dist, order=2)
    orc_mesh = oracle().mesh_square(node_dist, order=2)
    mesh = ctx.upload_mesh(host)
    U = 3
    # nodal x coordinates through the oracle's reference-to-physical map (test scaffolding only)
    gll = oracle().lobatto(3)
    xs = np.zeros(host.n_nodes)
    for e in range(host.n_elems):
        for a in range(9):
            xs[host.nodes[e, a]] = oracle().map_to_physical(2, host.verts[e], [gll[a % 3], gll[a // 3]])[0]
    bc_nodes = host.boundary_nodes([3, 4])
    mask = np.zeros(host.n_nodes * U, dtype=np.uint8)
    vals = np.zeros((host.n_nodes * U, 1))
    mask[bc_nodes * U] = 1
    vals[bc_nodes * U, 0] = xs[bc_nodes] / node_dist[-1]
    for strategy in (0, 1):
        opts = l3b.AssemblyOptions(1, 0, strategy)
        s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, vals)
        s.assembleProblem("diffusion_kernel_2D_r1", asm_opts=opts)
        s.assembleProblem("adiabatic_bc_2D", boundary_ids=[1, 2])
        s.endAssembly()
        so = orc_mesh.matrix_free_system(U, 1, mask, vals)
        so.add_kernel("diffusion_kernel_2D", 1, 0, strategy)
        so.add_kernel("adiabatic_bc_2D", boundary_ids=[1, 2])
        diag_o, rhs_o = so.init()
        diag_g, rhs_g = s.download()
        assert rel_err(diag_g, diag_o) < TOL and rel_err(rhs_g, rhs_o) < TOL
        x = np.random.default_rng(5).uniform(-1, 1, size=(host.n_nodes * U, 1))
        assert rel_err(s.apply(x), so.apply(x)) < TOL
        sol, tol, iters = s.solve(tol=1e-10)
        sol_o, tol_o, iters_o = so.cg(tol=1e-10)
        assert tol <= 1e-10
        # reference acceptance: L2 error < 1e-8; here the nodal errors
        assert np.abs(sol[0::3] - xs).max() < 1e-8
        assert np.abs(sol[1::3] - 1.0).max() < 1e-8
        assert np.abs(sol[2::3]).max() < 1e-8
        assert np.abs(sol - sol_o).max() < 1e-8
        assert abs(iters - iters_o) <= 2, (iters, iters_o)  # same algorithm; rounding may shift the stopping iteration


def test_single_element_fixtures(ctx):
    """tests/LocalOperatorCommon.hpp:17-61 fixtures through the device path: distorted quad p=4 / hex p=3 with
    asm_opts{.value_order = 2}; y == K_e x (LocalOperatorTests.cpp) and SF == local-element (SumFactorizationTests.cpp)."""
    orc = oracle()
    quad = dict(dim=2, p=4, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0]], k="diffusion_kernel_2D", kv="diffusion_kernel_2D_var", U=3, r=2)
    hexa = dict(dim=3, p=3, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0], [1, 1, 1], [2, 1, 1.5], [1, 3, 2], [3, 4, 3.5]],
                k="diffusion_kernel_3D", kv="diffusion_kernel_3D_var", U=4, r=3)
    for fx in (quad, hexa):
        dim, p, U = fx["dim"], fx["p"], fx["U"]
        nn = (p + 1) ** dim
        nodes = np.arange(nn, dtype=np.uint32)[None, :]
        verts = np.array(fx["verts"], dtype=float)[None]
        mesh = l3b.Mesh(ctx, dim, p, verts, nodes, None, nn, nn)
        opts = l3b.AssemblyOptions(value_order=2)
        rng = np.random.default_rng(9)
        # constant-coefficient kernel vs the oracle's explicitly assembled K_e
        K, _ = orc.assemble_local(fx["k"], dim, p, fx["verts"], n_rhs=fx["r"], value_order=2)
        x = rng.uniform(-1, 1, size=(nn * U, fx["r"]))
        for strategy in (0, 1):
            s = l3b.MatrixFreeSystem(ctx, mesh, U, fx["r"])
            s.assembleProblem(fx["k"], asm_opts=l3b.AssemblyOptions(2, 0, strategy))
            s.endAssembly()
            assert rel_err(s.apply(x), K @ x) < TOL
            diag, _ = s.download()
            assert rel_err(diag, np.diag(K)) < TOL
        # variable coefficient through an external field, n_rhs = 2
        field = rng.uniform(-1, 1, size=(1, nn))
        x = rng.uniform(-1, 1, size=(nn * U, 2))
        ys = []
        for strategy in (0, 1):
            s = l3b.MatrixFreeSystem(ctx, mesh, U, 2)
            s.assembleProblem(fx["kv"], fields=ctx.upload_fields(field), asm_opts=l3b.AssemblyOptions(2, 0, strategy))
            s.endAssembly()
            ys.append(s.apply(x))
        y_ref = orc.eval_local_operator(fx["kv"], dim, p, fx["verts"], x, node_vals=field.T.copy(), value_order=2)
        assert rel_err(ys[0], y_ref) < TOL and rel_err(ys[1], y_ref) < TOL


def test_degenerate_element_is_reported(ctx):
    """AssembleLocalSystem.hpp:249 / EvaluateLocalOperator.hpp:229, 295: |J| <= 0 raises; the SF path does not check (App. B.2)."""
    verts = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, -1], [1, 0, -1], [0, 1, -1], [1, 1, -1]]], dtype=float)
    nodes = np.arange(27, dtype=np.uint32)[None, :]
    mesh = l3b.Mesh(ctx, 3, 2, verts, nodes, None, 27, 27)
    s = l3b.MatrixFreeSystem(ctx, mesh, 4)
    s.assembleProblem("bench_diffusion3d")
    with pytest.raises(l3b.L3BError, match="degenerate element"):
        s.endAssembly()


def test_missing_instance_fails_loudly(ctx):
    pm = PairedMesh(3, default_dists(3, 2), 3)
    mesh = pm.upload(ctx)
    s = l3b.MatrixFreeSystem(ctx, mesh, 4)
    with pytest.raises(l3b.L3BError, match="not compiled for order 3, nq 10"):
        s.assembleProblem("bench_diffusion3d", asm_opts=l3b.AssemblyOptions(value_order=3))


def test_empty_partition(ctx):
    """tests/EmptyPartitionTest.cpp:10-40: a rank with zero elements is legal."""
    mesh = l3b.Mesh(ctx, 3, 2, np.zeros((0, 8, 3)), np.zeros((0, 27), dtype=np.uint32), None, 0, 0)
    s = l3b.MatrixFreeSystem(ctx, mesh, 4)
    s.assembleProblem("bench_diffusion3d")
    s.endAssembly()
    y = s.apply(np.zeros((0, 1)))
    assert y.shape == (0, 1)


def test_line_per_thread_kernel_passes_the_same_suite():
    """L3B_MF_LINES=1 routes hexahedra through mfSumFactApplyKernel (mf_sumfact.cuh) instead of the planes + columns kernel: the
    whole module must pass with it too (the selection is read once per process, hence the subprocess)."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, L3B_MF_LINES="1")
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__, "-k", "not line_per_thread"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]


def test_karman_kernels_matrix_free(ctx):
    """examples/07-karman-2D kernels through the matrix-free system: sum-factorised apply with previous-field access (quad p=4,
    nq = 8 > nb = 5) + the outlet boundary kernel on the dofs (u, v, p), against the assembled oracle matrix."""
    import scipy.sparse as sp

    U = 4
    pm = PairedMesh(2, default_dists(2, 3), 4)
    mesh = pm.upload(ctx)
    fdata = np.random.default_rng(11).uniform(-1, 1, size=(4, pm.n_nodes))
    fields = ctx.upload_fields(fdata)
    opts = l3b.AssemblyOptions(value_order=1, derivative_order=1)
    s = l3b.MatrixFreeSystem(ctx, mesh, U, 1)
    s.assembleProblem("karman_transient", fields=fields, asm_opts=opts, time=0.1)
    s.assembleProblem("karman_outlet", boundary_ids=[2], dof_inds=[0, 1, 3], asm_opts=opts)
    s.endAssembly()
    so = pm.orc.assembled_system(U)
    so.assemble_ex("karman_transient", 1, 1, 0.1, fdata, n_threads=4)
    so.assemble_ex("karman_outlet", 1, 1, 0.0, None, boundary_ids=[2], dof_inds=[0, 1, 3])
    vals_o, rhs_o = so.get()
    A = sp.csr_matrix((vals_o, so.col_ind, so.row_ptr), shape=(so.n_dofs,) * 2)
    x = np.random.default_rng(3).uniform(-1, 1, size=(so.n_dofs, 1))
    assert rel_err(s.apply(x), A @ x) < TOL
    diag, rhs = s.download()
    assert rel_err(diag, A.diagonal()) < TOL and rel_err(rhs, rhs_o) < TOL


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[3], CASES[6], CASES[9], CASES[11]], ids=lambda c: f"{c[0]}-d{c[2]}-p{c[4]}")
@pytest.mark.parametrize("strategy", [0, 1], ids=["sumfact", "local_element"])
def test_apply_energy_is_x_dot_Ax(ctx, case, strategy):
    """The device scalar of l3b_mf_apply_phase_device collects x^T A x (CG's p.Ap) inside the element kernels — planes + columns, line per
    thread and dense local-element kernels, with Dirichlet identity rows — and must equal the dot product of x with the applied y."""
    import torch

    kname, _, dim, n, p, opts, n_rhs = case
    opts = l3b.AssemblyOptions(opts.value_order, opts.derivative_order, strategy)
    info = l3b.kernel_info(kname)
    U, NF = info["n_unknowns"], info["n_fields"]
    pm = PairedMesh(dim, default_dists(dim, n), p)
    mesh = pm.upload(ctx)
    fdata = _fields(NF, pm.n_nodes, 7) if NF else None
    mask, dvals = _dirichlet(pm, U, [1, 2 * dim], [0, U - 1], n_rhs, 3)
    s = l3b.MatrixFreeSystem(ctx, mesh, U, n_rhs, mask, dvals)
    fields = ctx.upload_fields(fdata) if NF else None
    s.assembleProblem(kname, fields=fields, asm_opts=opts, time=0.37)
    if dim == 2 and kname == "example02_domain":
        s.assembleProblem("example02_bc", boundary_ids=[2, 3])  # a boundary kernel in the same operator
    s.endAssembly()
    x = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, size=s.n_dofs)).cuda()
    y = torch.zeros_like(x)
    e = torch.zeros(1, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    s.apply_device(x.data_ptr(), y.data_ptr(), 1, -0.7, 0.0, energy_ptr=e.data_ptr())  # alpha must not enter the energy
    ctx.synchronize()
    xAx = float(torch.dot(x, y).item()) / -0.7
    assert abs(e.item() - xAx) <= 1e-12 * abs(xAx)
    # the same in phases over two element ranges (the multi-rank apply): the shares add up
    e.zero_()
    torch.cuda.synchronize()
    half = pm.host.n_elems // 2
    s.apply_phase_device(x.data_ptr(), y.data_ptr(), l3b.APPLY_INIT, 0, 0)
    s.apply_phase_device(x.data_ptr(), y.data_ptr(), l3b.APPLY_ELEMENTS, half, pm.host.n_elems, energy_ptr=e.data_ptr())
    s.apply_phase_device(x.data_ptr(), y.data_ptr(), l3b.APPLY_ELEMENTS, 0, half, energy_ptr=e.data_ptr())
    s.apply_phase_device(x.data_ptr(), y.data_ptr(), l3b.APPLY_FINISH, 0, 0, energy_ptr=e.data_ptr())
    ctx.synchronize()
    assert abs(e.item() - xAx) <= 1e-12 * abs(xAx)


@pytest.mark.parametrize("p", [2, 3])
def test_3d_boundary_equation_kernel_matrix_free(ctx, p):
    """robin_bc_3D next to the domain kernel in the matrix-free operator (boundary kernels take evaluateLocalOperator on the parent
    element, MatrixFreeSystem.hpp:539-551, 690-709): apply, diag and rhs against the oracle, with Dirichlet dofs"""
    pm = PairedMesh(3, default_dists(3, 2), p)
    mesh = pm.upload(ctx)
    U = 4
    bnd = [1, 2, 4, 5]
    mask, dvals = _dirichlet(pm, U, [3, 6], [0], 1, 5)
    s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, dvals)
    s.assembleProblem("bench_diffusion3d")
    s.assembleProblem("robin_bc_3D", boundary_ids=bnd)
    s.endAssembly()
    so = pm.orc.matrix_free_system(U, 1, mask, dvals)
    so.add_kernel("bench_diffusion3d")
    so.add_kernel("robin_bc_3D", boundary_ids=bnd)
    diag_o, rhs_o = so.init(n_threads=4)
    diag_g, rhs_g = s.download()
    assert rel_err(diag_g, diag_o) < TOL and rel_err(rhs_g, rhs_o) < TOL
    x = np.random.default_rng(2).uniform(-1, 1, size=(pm.n_nodes * U, 1))
    assert rel_err(s.apply(x), so.apply(x, n_threads=4)) < TOL
    # and the boundary kernel's share alone
    sb = l3b.MatrixFreeSystem(ctx, mesh, U, 1, None, None)
    sb.assembleProblem("robin_bc_3D", boundary_ids=bnd)
    sb.endAssembly()
    sob = pm.orc.matrix_free_system(U, 1, None, None)
    sob.add_kernel("robin_bc_3D", boundary_ids=bnd)
    assert rel_err(sb.apply(x), sob.apply(x, n_threads=4)) < TOL


STREAMED_CASES = [
    # kernel, dim, n, order, opts, n_rhs, chunks, block_nodes
    ("bench_diffusion3d", 3, 3, 4, l3b.AssemblyOptions(), 1, 7, 64),
    ("bench_diffusion3d", 3, 4, 2, l3b.AssemblyOptions(), 1, 64, 1),
    ("diffusion_kernel_3D", 3, 2, 3, l3b.AssemblyOptions(value_order=2), 3, 3, 50),
    ("diffusion_kernel_2D", 2, 5, 4, l3b.AssemblyOptions(value_order=2), 2, 4, 32),
    ("bench_diffusion3d", 3, 2, 4, l3b.AssemblyOptions(eval_strategy=1), 1, 5, 100),
]


@pytest.mark.parametrize("case", STREAMED_CASES, ids=lambda c: f"{c[0]}-d{c[1]}-n{c[2]}-p{c[3]}-rhs{c[5]}-chunks{c[6]}-block{c[7]}")
def test_streamed_host_apply_matches_oracle_and_serial_form(ctx, case):
    """l3b_mf_apply with host vectors, streamed (x blocks in, element chunks, y blocks out on three streams, Dirichlet rows per finished
    block) against the oracle and against the serial form of the same call"""
    kname, dim, n, p, opts, n_rhs, chunks, block = case
    U = l3b.kernel_info(kname)["n_unknowns"]
    pm = PairedMesh(dim, default_dists(dim, n), p)
    mesh = pm.upload(ctx)
    mask, dvals = _dirichlet(pm, U, [1, 2 * dim], [0, U - 1], n_rhs, 3)
    sys_g = l3b.MatrixFreeSystem(ctx, mesh, U, n_rhs, mask, dvals)
    sys_g.assembleProblem(kname, asm_opts=opts, time=0.2)
    sys_g.endAssembly()
    sys_o = pm.orc.matrix_free_system(U, n_rhs, mask, dvals)
    sys_o.add_kernel(kname, opts.value_order, opts.derivative_order, opts.eval_strategy, 0.2, None)
    sys_o.init(n_threads=4)
    rng = np.random.default_rng(17)
    for n_cols in sorted({n_rhs, 1}):
        x = rng.uniform(-1, 1, size=(pm.n_nodes * U, n_cols))
        y_o = sys_o.apply(x, np.zeros_like(x), -1.3, 0.0, n_threads=4)
        sys_g.set_host_apply(0)
        y_serial = sys_g.apply(x, None, -1.3, 0.0)
        assert not sys_g.host_apply_info()["streamed"]
        sys_g.set_host_apply(2, chunks, block)
        for _ in range(2):  # the second call reuses the schedule, the streams and the events
            y_g = sys_g.apply(x, None, -1.3, 0.0)
            info = sys_g.host_apply_info()
            assert info["streamed"] and info["items"] == -(-mesh.n_elems // -(-mesh.n_elems // chunks))
            assert rel_err(y_g, y_o) < TOL and rel_err(y_g, y_serial) < TOL
        # beta != 0 needs y over PCIe both ways: the call takes the serial form and stays correct
        y0 = rng.uniform(-1, 1, size=x.shape)
        y_b = sys_g.apply(x, y0, 0.5, -0.25)
        assert not sys_g.host_apply_info()["streamed"]
        assert rel_err(y_b, sys_o.apply(x, y0, 0.5, -0.25, n_threads=4)) < TOL


def test_streamed_host_apply_with_a_numbering_without_locality(ctx):
    """scrambled node ids and element order: the schedule degenerates (most blocks travel with the first / last item), the result does not"""
    pm = PairedMesh(3, default_dists(3, 3), 3)
    rng = np.random.default_rng(23)
    perm = rng.permutation(pm.n_nodes)
    eorder = rng.permutation(pm.host.n_elems)
    nodes = perm[pm.host.nodes.astype(np.int64)][eorder].astype(np.uint32)
    mesh = l3b.Mesh(ctx, 3, 3, pm.verts[eorder], nodes, pm.host.side_boundaries[eorder], pm.n_nodes, pm.n_nodes)
    U = 4
    mask0, _ = _dirichlet(pm, U, [1, 6], [0], 1, 3)
    mask = np.zeros_like(mask0)
    mask.reshape(-1, U)[perm] = mask0.reshape(-1, U)
    sys_g = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    sys_g.assembleProblem("bench_diffusion3d")
    sys_g.endAssembly()
    sys_o = pm.orc.matrix_free_system(U, 1, mask0, None)
    sys_o.add_kernel("bench_diffusion3d")
    sys_o.init(n_threads=4)
    x0 = rng.uniform(-1, 1, size=(pm.n_nodes * U, 1))
    x = np.zeros_like(x0)
    x.reshape(-1, U)[perm] = x0.reshape(-1, U)
    sys_g.set_host_apply(2, 6, 16)
    y = sys_g.apply(x)
    assert sys_g.host_apply_info()["streamed"]
    y_o = sys_o.apply(x0, np.zeros_like(x0), 1.0, 0.0, n_threads=4)
    assert rel_err(y.reshape(-1, U)[perm].reshape(-1, 1), y_o) < TOL


def test_host_apply_with_a_boundary_kernel_takes_the_serial_form(ctx):
    pm = PairedMesh(2, default_dists(2, 4), 4)
    mesh = pm.upload(ctx)
    s = l3b.MatrixFreeSystem(ctx, mesh, 3)
    s.assembleProblem("example02_domain")
    s.assembleProblem("example02_bc", boundary_ids=[1, 2, 3, 4])
    s.endAssembly()
    x = np.random.default_rng(1).uniform(-1, 1, size=(pm.n_nodes * 3, 1))
    s.set_host_apply(0)
    y0 = s.apply(x)
    s.set_host_apply(2, 4, 8)
    y1 = s.apply(x)
    assert not s.host_apply_info()["streamed"] and rel_err(y1, y0) < TOL


def test_streamed_host_apply_with_kernels_restricted_to_element_domains(ctx):
    """assembleProblem(kernel, domain_ids): each use runs over a sub-list of the elements; the streamed call cuts those lists by chunk"""
    pm = PairedMesh(3, default_dists(3, 3), 3)
    mesh = pm.upload(ctx)
    mesh.set_element_domains(np.arange(pm.host.n_elems, dtype=np.int32) % 3)
    U = 4
    mask, _ = _dirichlet(pm, U, [1, 6], [0], 1, 3)
    whole = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    whole.assembleProblem("bench_diffusion3d")
    whole.endAssembly()
    parts = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    parts.assembleProblem("bench_diffusion3d", boundary_ids=[0, 2])
    parts.assembleProblem("bench_diffusion3d", boundary_ids=[1])
    parts.endAssembly()
    x = np.random.default_rng(3).uniform(-1, 1, size=(pm.n_nodes * U, 1))
    whole.set_host_apply(0)
    y_ref = whole.apply(x)
    parts.set_host_apply(2, 5, 40)
    y = parts.apply(x)
    assert parts.host_apply_info()["streamed"] and rel_err(y, y_ref) < TOL
