"""GPU parity: element-local least-squares assembly fused with the CRS scatter, algebraic Dirichlet BCs and the assembled
solve, against the CPU oracle on identical inputs (through the C ABI). Sparsity is compared bit-exactly, values to 1e-12."""
import os

import numpy as np
import pytest

import l3ster_b200 as l3b
from common import PairedMesh, default_dists, oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    return l3b.Context(0)


CASES = [
    ("bench_diffusion3d", "bench_diffusion3d", 3, 2, 4, l3b.AssemblyOptions(), 1),
    ("bench_diffusion3d", "bench_diffusion3d", 3, 3, 2, l3b.AssemblyOptions(), 1),
    ("bench_diffusion3d", "bench_diffusion3d", 3, 2, 1, l3b.AssemblyOptions(), 1),
    ("bench_diffusion3d", "bench_diffusion3d", 3, 1, 6, l3b.AssemblyOptions(), 1),
    ("diffusion_kernel_3D", "diffusion_kernel_3D", 3, 2, 3, l3b.AssemblyOptions(value_order=2), 3),
    ("diffusion_kernel_3D_var", "diffusion_kernel_3D_var", 3, 2, 3, l3b.AssemblyOptions(value_order=2), 2),
    ("dense_probe_3D", "dense_probe_3D", 3, 2, 2, l3b.AssemblyOptions(), 1),
    ("diffusion_kernel_2D", "diffusion_kernel_2D", 2, 3, 4, l3b.AssemblyOptions(value_order=2), 2),
    ("diffusion_kernel_2D_var", "diffusion_kernel_2D_var", 2, 3, 3, l3b.AssemblyOptions(), 2),
    ("dense_probe_2D", "dense_probe_2D", 2, 4, 2, l3b.AssemblyOptions(value_order=1, derivative_order=1), 1),
    # BASELINE configs[4], benchmarks/LocalAssemblyBenchmarks.cpp:41-87: NS3D, U = 7, E = 8, n_fields = 7, QO = 4p - 1 (= AssemblyOptions{1, 1})
    ("ns3d_kernel", "ns3d_kernel", 3, 2, 2, l3b.AssemblyOptions(value_order=1, derivative_order=1), 1),
    ("ns3d_kernel", "ns3d_kernel", 3, 2, 3, l3b.AssemblyOptions(), 1),
    ("ns3d_kernel", "ns3d_kernel", 3, 1, 4, l3b.AssemblyOptions(value_order=1, derivative_order=1), 1),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-d{c[2]}-n{c[3]}-p{c[4]}")
def test_assembled_system_matches_oracle(ctx, case):
    kname, oname, dim, n, p, opts, n_rhs = case
    info = l3b.kernel_info(kname)
    U, NF = info["n_unknowns"], info["n_fields"]
    pm = PairedMesh(dim, default_dists(dim, n), p)
    mesh = pm.upload(ctx)
    fdata = np.random.default_rng(7).uniform(-1, 1, size=(NF, pm.n_nodes)) if NF else None
    time = 0.21
    sys_g = l3b.AssembledSystem(ctx, mesh, U, n_rhs)
    sys_o = pm.orc.assembled_system(U, n_rhs)
    # sparsity graph: bit-exact (tests/SparsityGraphTest.cpp contract)
    row_ptr, col_ind = sys_g.graph()
    assert np.array_equal(row_ptr, sys_o.row_ptr) and np.array_equal(col_ind, sys_o.col_ind)
    for _ in range(2):  # beginAssembly must re-zero
        sys_g.beginAssembly()
        sys_g.assembleProblem(kname, fields=ctx.upload_fields(fdata) if NF else None, asm_opts=opts, time=time)
    sys_o.assemble(oname, opts.value_order, opts.derivative_order, time, fdata, n_threads=4)
    vals_o, rhs_o = sys_o.get()
    vals_g, rhs_g = sys_g.download()
    assert rel_err(vals_g, vals_o) < TOL
    assert np.linalg.norm(rhs_g - rhs_o) <= TOL * max(np.linalg.norm(rhs_o), 1.0)
    # Dirichlet BCs on two faces, dof 0 and the last dof (bcs/DirichletBC.hpp:82-150)
    nodes = pm.host.boundary_nodes([1, 2 * dim])
    dofs = np.sort(np.concatenate([nodes * U, nodes * U + U - 1])).astype(np.int32)
    dvals = np.random.default_rng(3).uniform(-1, 1, size=(len(dofs), n_rhs))
    sys_g.endAssembly(dofs, dvals)
    sys_o.apply_dirichlet(dofs, dvals)
    vals_o, rhs_o = sys_o.get()
    vals_g, rhs_g = sys_g.download()
    assert rel_err(vals_g, vals_o) < TOL
    assert rel_err(rhs_g, rhs_o) < TOL
    x = np.random.default_rng(4).uniform(-1, 1, size=sys_g.n_dofs)
    import scipy.sparse as sp
    A = sp.csr_matrix((vals_o, col_ind, row_ptr), shape=(sys_g.n_dofs,) * 2)
    assert rel_err(sys_g.spmv(x), A @ x) < TOL


def test_assembled_matches_matrix_free(ctx):
    """The two evaluation strategies describe the same operator: A_assembled x == matrix-free apply (no BCs)."""
    pm = PairedMesh(3, default_dists(3, 2), 4)
    mesh = pm.upload(ctx)
    a = l3b.AssembledSystem(ctx, mesh, 4)
    a.beginAssembly()
    a.assembleProblem("bench_diffusion3d")
    a.endAssembly()
    m = l3b.MatrixFreeSystem(ctx, mesh, 4)
    m.assembleProblem("bench_diffusion3d")
    m.endAssembly()
    x = np.random.default_rng(0).uniform(-1, 1, size=a.n_dofs)
    assert rel_err(a.spmv(x), m.apply(x)[:, 0]) < TOL
    diag, rhs = m.download()
    _, rhs_a = a.download(values=False)
    assert rel_err(rhs_a, rhs) < TOL
    assert rel_err(a.getMatrix().diagonal(), diag) < TOL


def test_diffusion2d_assembled_end_to_end(ctx):
    """tests/Diffusion2DAssembledTest.cpp via tests/Diffusion2D.hpp:23-120 and BASELINE config 1's ingredients (domain +
    boundary kernel assembled into one CRS matrix): 4x4 quads p=2, Dirichlet T = x left/right, adiabatic top/bottom."""
    node_dist = np.linspace(0.0, 1.0, 5)
    host = l3b.make_square_mesh(node_dist, order=2)
    orc_mesh = oracle().mesh_square(node_dist, order=2)
    mesh = ctx.upload_mesh(host)
    U = 3
    gll = oracle().lobatto(3)
    xs = np.zeros(host.n_nodes)
    for e in range(host.n_elems):
        for a_ in range(9):
            xs[host.nodes[e, a_]] = oracle().map_to_physical(2, host.verts[e], [gll[a_ % 3], gll[a_ // 3]])[0]
    bc_nodes = host.boundary_nodes([3, 4])
    dofs = (bc_nodes * U).astype(np.int32)
    vals = xs[bc_nodes][:, None]
    s = l3b.AssembledSystem(ctx, mesh, U)
    s.beginAssembly()
    s.assembleProblem("diffusion_kernel_2D_r1")
    s.assembleProblem("adiabatic_bc_2D", boundary_ids=[1, 2])
    so = orc_mesh.assembled_system(U)
    so.assemble("diffusion_kernel_2D")
    so.assemble("adiabatic_bc_2D", boundary_ids=[1, 2])
    v_g, r_g = s.download()
    v_o, r_o = so.get()
    assert rel_err(v_g, v_o) < TOL
    s.endAssembly(dofs, vals)
    sol, tol, iters = s.solve(tol=1e-10)
    assert tol <= 1e-10
    assert np.abs(sol[0::3] - xs).max() < 1e-8
    assert np.abs(sol[1::3] - 1.0).max() < 1e-8
    assert np.abs(sol[2::3]).max() < 1e-8


def test_example02_diffusion2d(ctx):
    """BASELINE config 1 (examples/02-diffusion-2D/source.cpp:10-81) at reduced size: quad p=4, domain kernel with source
    + Robin-type boundary kernel on all four sides, assembled; matrix and rhs against the oracle, solve against a direct
    solve of the oracle system (the reference uses KLU2)."""
    node_dist = np.linspace(0.0, 1.0, 6)
    host = l3b.make_square_mesh(node_dist, order=4)
    orc_mesh = oracle().mesh_square(node_dist, order=4)
    mesh = ctx.upload_mesh(host)
    s = l3b.AssembledSystem(ctx, mesh, 3)
    s.beginAssembly()
    s.assembleProblem("example02_domain")
    s.assembleProblem("example02_bc", boundary_ids=[1, 2, 3, 4])
    s.endAssembly()
    so = orc_mesh.assembled_system(3)
    so.assemble("example02_domain")
    so.assemble("example02_bc", boundary_ids=[1, 2, 3, 4])
    v_o, r_o = so.get()
    v_g, r_g = s.download()
    assert rel_err(v_g, v_o) < TOL and rel_err(r_g, r_o) < TOL
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    A = sp.csr_matrix((v_o, so.col_ind, so.row_ptr), shape=(so.n_dofs,) * 2).tocsc()
    ref = spla.spsolve(A, r_o[:, 0])
    sol, tol, iters = s.solve(tol=1e-12, max_iters=20000)
    assert np.abs(sol - ref).max() < 1e-8 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("force_fma", [False, True], ids=["auto", "dfma"])
def test_single_element_fixtures(ctx, force_fma):
    """tests/LocalAssemblyTests.cpp:3-43 + tests/LocalOperatorCommon.hpp:17-61 through the device path: K_e, F_e of the distorted quad
    p=4 / hex p=3 fixtures (asm_opts{.value_order = 2}) against the oracle's assembleLocalSystem, entry by entry. The single-element
    CRS is the dense row-major K_e. Runs in a subprocess for the DFMA kernel (the selection is read once per process)."""
    if force_fma:
        import subprocess
        import sys

        # the whole module again with each kernel forced for every element size (the default picks by nodes per element)
        for forced in ("1", "0"):
            env = dict(os.environ, L3B_ASM_FMA=forced)
            out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__, "-k", "not dfma"],
                                 env=env, capture_output=True, text=True, timeout=900)
            assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
        return
    orc = oracle()
    quad = dict(dim=2, p=4, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0]], k="diffusion_kernel_2D", U=3, r=2)
    hexa = dict(dim=3, p=3, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0], [1, 1, 1], [2, 1, 1.5], [1, 3, 2], [3, 4, 3.5]],
                k="diffusion_kernel_3D", U=4, r=3)
    for fx in (quad, hexa):
        dim, p, U = fx["dim"], fx["p"], fx["U"]
        nn = (p + 1) ** dim
        nodes = np.arange(nn, dtype=np.uint32)[None, :]
        verts = np.array(fx["verts"], dtype=float)[None]
        mesh = l3b.Mesh(ctx, dim, p, verts, nodes, None, nn, nn)
        s = l3b.AssembledSystem(ctx, mesh, U, fx["r"])
        s.beginAssembly()
        s.assembleProblem(fx["k"], asm_opts=l3b.AssemblyOptions(value_order=2))
        vals, rhs = s.download()
        K, F = orc.assemble_local(fx["k"], dim, p, fx["verts"], n_rhs=fx["r"], value_order=2)
        Kd = vals.reshape(nn * U, nn * U)
        assert rel_err(Kd, K) < TOL
        assert np.abs(Kd - Kd.T).max() <= 1e-13 * np.abs(K).max()
        assert np.abs(rhs - F).max() <= TOL * max(np.abs(F).max(), 1.0)


def test_karman_kernels_assembled(ctx):
    """BASELINE configs[3] (examples/07-karman-2D/source.cpp:21-155): the steady and the BDF2 Navier-Stokes kernels (U=4, previous-field
    values and gradients, AssemblyOptions{1, 1} -> nq = 8) plus the outlet boundary kernel on the dof subset (u, v, p), quad p=4,
    distorted structured mesh instead of the Gmsh file; matrix, rhs and the assembled solve's operator against the oracle."""
    U = 4
    pm = PairedMesh(2, default_dists(2, 4), 4)
    mesh = pm.upload(ctx)
    fdata = np.random.default_rng(11).uniform(-1, 1, size=(6, pm.n_nodes))  # SolutionManager with 6 stored fields
    fields = ctx.upload_fields(fdata)
    opts = l3b.AssemblyOptions(value_order=1, derivative_order=1)
    for kname, finds in (("karman_steady", [4, 1]), ("karman_transient", [0, 1, 2, 3])):
        sys_g = l3b.AssembledSystem(ctx, mesh, U)
        sys_o = pm.orc.assembled_system(U)
        sys_g.beginAssembly()
        sys_g.assembleProblem(kname, fields=fields, field_inds=finds, asm_opts=opts, time=0.3)
        sys_g.assembleProblem("karman_outlet", boundary_ids=[2], dof_inds=[0, 1, 3], asm_opts=opts)
        sys_o.assemble_ex(kname, 1, 1, 0.3, fdata, n_threads=4, field_inds=finds)
        sys_o.assemble_ex("karman_outlet", 1, 1, 0.0, None, n_threads=1, boundary_ids=[2], dof_inds=[0, 1, 3])
        vals_o, rhs_o = sys_o.get()
        vals_g, rhs_g = sys_g.download()
        assert rel_err(vals_g, vals_o) < TOL, kname
        assert rel_err(rhs_g, rhs_o) < TOL, kname
        # Dirichlet u, v on the inlet and the walls (source.cpp:158-180), then the operator of the closed system
        nodes = pm.host.boundary_nodes([1, 3, 4])
        dofs = np.sort(np.concatenate([nodes * U, nodes * U + 1])).astype(np.int32)
        dvals = np.random.default_rng(5).uniform(-1, 1, size=(len(dofs), 1))
        sys_g.endAssembly(dofs, dvals)
        sys_o.apply_dirichlet(dofs, dvals)
        vals_o, rhs_o = sys_o.get()
        vals_g, rhs_g = sys_g.download()
        assert rel_err(vals_g, vals_o) < TOL and rel_err(rhs_g, rhs_o) < TOL, kname


def test_gmres_matches_cg_and_direct_solve(ctx):
    """lstr::Gmres (Belos Pseudoblock GMRES, left Jacobi, restart) on the assembled and on the matrix-free diffusion system: same
    solution as CG and as a direct solve of the oracle matrix (tests/SolverTests.cpp:172-238 check convergence only); a short
    restart length exercises the restart path."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    pm = PairedMesh(2, default_dists(2, 4), 4)  # example02_domain is registered for quad p=4, nq=5
    mesh = pm.upload(ctx)
    U = 3
    nodes = pm.host.boundary_nodes([1, 2, 3, 4])
    dofs = (nodes * U).astype(np.int32)
    dvals = np.random.default_rng(2).uniform(-1, 1, size=(len(dofs), 1))
    s = l3b.AssembledSystem(ctx, mesh, U)
    s.beginAssembly()
    s.assembleProblem("example02_domain")
    s.endAssembly(dofs, dvals)
    so = pm.orc.assembled_system(U)
    so.assemble("example02_domain")
    so.apply_dirichlet(dofs, dvals)
    v_o, r_o = so.get()
    A = sp.csr_matrix((v_o, so.col_ind, so.row_ptr), shape=(so.n_dofs,) * 2).tocsc()
    ref = spla.spsolve(A, r_o[:, 0])
    x_cg, _, it_cg = s.solve(tol=1e-11, max_iters=20000)
    for restart in (250, 30):
        x_g, res, it_g = s.solve_gmres(tol=1e-11, restart_length=restart, max_restarts=400, max_iters=20000)
        assert res <= 1e-11 and it_g > 0
        assert np.abs(x_g - ref).max() < 1e-8 * max(1.0, np.abs(ref).max()), restart
        assert np.abs(x_g - x_cg).max() < 1e-8 * max(1.0, np.abs(ref).max())
    # full GMRES on an SPD matrix needs no more iterations than CG needs for the same (preconditioned) residual
    assert s.solve_gmres(tol=1e-11, restart_length=2000)[2] <= it_cg + 5
    # matrix-free operator, same system
    mask = np.zeros(pm.n_nodes * U, dtype=np.uint8)
    mask[dofs] = 1
    vals = np.zeros((pm.n_nodes * U, 1))
    vals[dofs] = dvals
    mf = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, vals)
    mf.assembleProblem("example02_domain")
    mf.endAssembly()
    x_m, res_m, it_m = mf.solve_gmres(tol=1e-11, restart_length=250)
    assert np.abs(x_m - ref).max() < 1e-8 * max(1.0, np.abs(ref).max())


def test_assembly_state_machine_and_argument_errors(ctx):
    """The OpenForAssembly <-> Closed state machine of AssembledSystem (AssembledSystem.hpp:455-461) and the argument checks of the
    C ABI surface as errors with the reference's wording, never as silent wrong results."""
    pm = PairedMesh(2, default_dists(2, 2), 4)
    mesh = pm.upload(ctx)
    s = l3b.AssembledSystem(ctx, mesh, 3)
    with pytest.raises(l3b.L3BError):  # assembleProblem before beginAssembly
        s.assembleProblem("example02_domain")
    s.beginAssembly()
    with pytest.raises(l3b.L3BError):  # solve while open
        s.solve()
    with pytest.raises(l3b.L3BError):  # a 3-D kernel on a 2-D mesh
        s.assembleProblem("bench_diffusion3d")
    with pytest.raises(l3b.L3BError):  # dof index out of range
        s.assembleProblem("example02_domain", dof_inds=[0, 1, 7])
    with pytest.raises(l3b.L3BError):  # boundary kernel without boundary ids on a mesh without side table
        m2 = l3b.Mesh(ctx, 2, 4, pm.verts, pm.host.nodes, None, pm.n_nodes, pm.n_nodes)
        l3b.AssembledSystem(ctx, m2, 3).assembleProblem("example02_bc", boundary_ids=[1])
    s.assembleProblem("example02_domain")
    s.endAssembly()
    with pytest.raises(l3b.L3BError):  # endAssembly twice
        s.endAssembly()
    # the matrix-free system has the same gate
    mf = l3b.MatrixFreeSystem(ctx, mesh, 3, 1)
    mf.assembleProblem("example02_domain")
    with pytest.raises(l3b.L3BError):  # apply before endAssembly
        mf.apply(np.zeros((pm.n_nodes * 3, 1)))
    mf.endAssembly()
    with pytest.raises(l3b.L3BError):  # assembleProblem after endAssembly
        mf.assembleProblem("example02_domain")
    with pytest.raises(l3b.L3BError):  # 2 columns into a 1-rhs system
        mf.apply(np.zeros((pm.n_nodes * 3, 2)))


@pytest.mark.parametrize("p", [2, 4])
def test_3d_boundary_equation_kernel_assembled(ctx, p):
    """a 3-D boundary EQUATION kernel (robin_bc_3D: adiabatic row + Robin row with derivative operators, normal and boundary point)
    assembled with the domain kernel into one CRS matrix, on a distorted hex mesh: assembleLocalSystem(BoundaryElementView),
    AssembleLocalSystem.hpp:258-280 + mapBoundary"""
    pm = PairedMesh(3, default_dists(3, 2), p)
    mesh = pm.upload(ctx)
    U = 4
    bnd = [1, 2, 4, 5]
    s = l3b.AssembledSystem(ctx, mesh, U)
    s.beginAssembly()
    s.assembleProblem("bench_diffusion3d")
    s.assembleProblem("robin_bc_3D", boundary_ids=bnd)
    so = pm.orc.assembled_system(U)
    so.assemble("bench_diffusion3d", n_threads=4)
    vd, rd = so.get()
    vd, rd = vd.copy(), rd.copy()
    so.assemble("robin_bc_3D", boundary_ids=bnd, n_threads=4)
    v_o, r_o = so.get()
    v_g, r_g = s.download()
    assert rel_err(v_g, v_o) < TOL and rel_err(r_g, r_o) < TOL
    assert np.linalg.norm(v_o - vd) > 1e-3 * np.linalg.norm(vd) and np.linalg.norm(r_o - rd) > 0  # the boundary kernel contributes
    # the boundary kernel alone (its contribution is not hidden behind the domain kernel's larger entries)
    s.beginAssembly()
    s.assembleProblem("robin_bc_3D", boundary_ids=bnd)
    v_b, r_b = s.download()
    assert rel_err(v_b, v_o - vd) < 1e-10 and rel_err(r_b, r_o - rd) < 1e-10
