"""GPU parity: static condensation (CondensationPolicy::ElementBoundary, algsys/StaticCondensationManager.hpp) — the condensed matrix
and right-hand side against the restated reference formula on the oracle's assembled system, the recovered solution against the
uncondensed solve, and tests/Diffusion2DAssembledTest.cpp:9 (the condensed variant of tests/Diffusion2D.hpp) end to end."""
import numpy as np
import pytest
import scipy.sparse as sp

import l3ster_b200 as l3b
from l3ster_b200.condensation import CondensedAssembledSystem, boundary_interior_split
from common import PairedMesh, default_dists, oracle, rel_err
from oracle import condense_element_boundary

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return l3b.Context(0)


def _distorted_host(pm):
    h = pm.host
    h.verts[...] = pm.verts
    return h


CASES = [
    # domain kernel, boundary kernel or None, dim, n, order, opts
    ("diffusion_kernel_2D_r1", "adiabatic_bc_2D", 2, 3, 2, l3b.AssemblyOptions()),
    ("example02_domain", "example02_bc", 2, 2, 4, l3b.AssemblyOptions()),
    ("bench_diffusion3d", None, 3, 2, 2, l3b.AssemblyOptions()),
    ("bench_diffusion3d", None, 3, 1, 4, l3b.AssemblyOptions()),
    ("dense_probe_3D", None, 3, 2, 2, l3b.AssemblyOptions()),
    ("bench_diffusion3d", None, 3, 1, 5, l3b.AssemblyOptions()),  # 256 interior dofs: the largest the register-resident inverse takes
    ("bench_diffusion3d", None, 3, 1, 6, l3b.AssemblyOptions()),  # 500 interior dofs (the shipped benchmark's order): the blocked kernel
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-d{c[2]}-n{c[3]}-p{c[4]}")
def test_condensed_system_matches_reference_formula(ctx, case):
    dom, bnd, dim, n, p, opts = case
    info = l3b.kernel_info(dom)
    U, NF = info["n_unknowns"], info["n_fields"]
    pm = PairedMesh(dim, default_dists(dim, n), p)
    host = _distorted_host(pm)
    fdata = np.random.default_rng(3).uniform(-1, 1, size=(NF, pm.n_nodes)) if NF else None
    cs = CondensedAssembledSystem(ctx, host, U)
    cs.beginAssembly()
    cs.assembleProblem(dom, field_data=fdata, asm_opts=opts, time=0.2)
    so = pm.orc.assembled_system(U)
    so.assemble({"diffusion_kernel_2D_r1": "diffusion_kernel_2D"}.get(dom, dom), opts.value_order, opts.derivative_order, 0.2, fdata)
    if bnd:
        cs.assembleProblem(bnd, boundary_ids=[1, 2])
        so.assemble(bnd, boundary_ids=[1, 2])
    cs.endAssembly()
    vals, rhs = so.get()
    ref = l3b.AssembledSystem(ctx, pm.upload(ctx), U)  # only for the graph of the uncondensed system
    row_ptr, col_ind = ref.graph()
    K = sp.csr_matrix((vals, col_ind, row_ptr), shape=(len(rhs), len(rhs))).toarray()
    bnd_idx, int_idx = boundary_interior_split(dim, p)
    prim, S, Fc, recover = condense_element_boundary(K, rhs, host.nodes, bnd_idx, int_idx, U)
    assert np.array_equal(prim, cs.primary_nodes)
    S_g = cs.condensed.getMatrix().toarray()
    _, F_g = cs.condensed.download(values=False)
    assert rel_err(S_g, S) < 1e-11
    assert rel_err(F_g, Fc) < 1e-11
    # the sparsity of the condensed matrix is that of the primary nodes' graph: nothing outside it, structurally
    assert S_g.shape == (len(prim) * U, len(prim) * U)
    # recovery of the interior values from a condensed solution (any vector will do: it is a linear map)
    xc = np.random.default_rng(5).uniform(-1, 1, size=len(prim) * U)
    assert rel_err(cs.recover(xc), recover(xc)) < 1e-10


def test_condensed_solve_equals_uncondensed_solve(ctx):
    """3-D diffusion benchmark set-up (Dirichlet T = 0 on the six faces, source 1) with and without condensation: same nodal solution"""
    pm = PairedMesh(3, default_dists(3, 2), 3)
    host = _distorted_host(pm)
    U = 4
    bc_nodes = host.boundary_nodes([1, 2, 3, 4, 5, 6])
    dofs, vals = (bc_nodes * U).astype(np.int32), np.zeros((len(bc_nodes), 1))
    full = l3b.AssembledSystem(ctx, pm.upload(ctx), U)
    full.beginAssembly()
    full.assembleProblem("bench_diffusion3d")
    full.endAssembly(dofs, vals)
    x_full, tol_f, it_f = full.solve(tol=1e-12)
    cs = CondensedAssembledSystem(ctx, host, U)
    cs.beginAssembly()
    cs.assembleProblem("bench_diffusion3d")
    cs.endAssembly(dofs, vals)
    x_cond, tol_c, it_c = cs.solve(tol=1e-12)
    assert tol_f <= 1e-12 and tol_c <= 1e-12
    assert cs.n_primary_dofs < full.n_dofs and it_c <= it_f  # fewer unknowns, better conditioned (the point of the policy)
    assert rel_err(x_cond, x_full) < 1e-9


def test_diffusion3d_benchmark_as_shipped(ctx):
    """benchmarks/Diffusion3DBenchmark.cpp:6 + Diffusion3D.hpp:8-118 as shipped: hex p = 6, CondensationPolicy::ElementBoundary, Dirichlet
    T = 0 on the six faces, CG + Jacobi to 1e-6 (4^3 elements here instead of 6^3 to keep the test short; 500 interior dofs per element).
    The nodal solution must equal the matrix-free (uncondensed) solve of the same problem."""
    n, p, U = 4, 6, 4
    dx, x, nd = 1.0 / n, 0.0, []
    for _ in range(n + 1):
        nd.append(x)
        x += dx
    host = l3b.make_cube_mesh(np.array(nd), order=p)
    bc_nodes = host.boundary_nodes([1, 2, 3, 4, 5, 6])
    cs = CondensedAssembledSystem(ctx, host, U)
    cs.beginAssembly()
    cs.assembleProblem("bench_diffusion3d")
    cs.endAssembly((bc_nodes * U).astype(np.int32), np.zeros((len(bc_nodes), 1)))
    x_cond, tol_c, it_c = cs.solve(tol=1e-10)
    assert tol_c <= 1e-10
    mask = np.zeros(host.n_nodes * U, dtype=np.uint8)
    mask[bc_nodes * U] = 1
    mf = l3b.MatrixFreeSystem(ctx, ctx.upload_mesh(host), U, 1, mask, None)
    mf.assembleProblem("bench_diffusion3d")
    mf.endAssembly()
    x_mf, tol_m, it_m = mf.solve(tol=1e-10)
    assert tol_m <= 1e-10 and it_c < it_m  # fewer, better conditioned unknowns
    assert cs.n_primary_dofs < host.n_nodes * U
    assert rel_err(x_cond, x_mf) < 1e-7
    assert np.abs(x_mf[0::U]).max() > 1e-3


def test_diffusion2d_condensed_end_to_end(ctx):
    """tests/Diffusion2DAssembledTest.cpp:9 = tests/Diffusion2D.hpp:23-117 with CondensationPolicy::ElementBoundary"""
    node_dist = np.linspace(0.0, 1.0, 5)
    host = l3b.make_square_mesh(node_dist, order=2)
    mesh = ctx.upload_mesh(host)
    U = 3
    gll = oracle().lobatto(3)
    xs = np.zeros(host.n_nodes)
    for e in range(host.n_elems):
        for a_ in range(9):
            xs[host.nodes[e, a_]] = oracle().map_to_physical(2, host.verts[e], [gll[a_ % 3], gll[a_ // 3]])[0]
    bc_nodes = host.boundary_nodes([3, 4])
    cs = CondensedAssembledSystem(ctx, host, U)
    cs.beginAssembly()
    cs.assembleProblem("diffusion_kernel_2D_r1")
    cs.assembleProblem("adiabatic_bc_2D", boundary_ids=[1, 2])
    cs.endAssembly((bc_nodes * U).astype(np.int32), xs[bc_nodes][:, None])
    sol, tol, _ = cs.solve(tol=1e-10)
    assert tol <= 1e-10
    fields = ctx.upload_fields(np.ascontiguousarray(sol.reshape(-1, U).T))
    assert np.linalg.norm(mesh.computeNormL2("diffusion2d_error_dom", fields=fields)) < 1e-8
    assert np.linalg.norm(mesh.computeNormL2("diffusion2d_error_bnd", boundary_ids=[1, 2, 3, 4], fields=fields)) < 1e-8


def test_dirichlet_on_interior_node_is_rejected(ctx):
    host = l3b.make_square_mesh(np.linspace(0, 1, 3), order=2)
    cs = CondensedAssembledSystem(ctx, host, 3)
    cs.beginAssembly()
    cs.assembleProblem("diffusion_kernel_2D_r1")
    interior_node = int(host.nodes[0, 4])  # the centre node of the first p=2 quad
    with pytest.raises(l3b.L3BError):
        cs.endAssembly(np.array([interior_node * 3], dtype=np.int32), np.zeros((1, 1)))


def test_condensed_slab_operator_on_one_rank(ctx):
    """SlabAssembledOperator(condensed=True), the multi-rank form of the condensed system, on a single slab: same solution as
    CondensedAssembledSystem and as the uncondensed slab operator (distributed: tests/mp_slab_apply.py)"""
    from l3ster_b200.slab import SlabAssembledOperator, make_slab

    x1, y1, z1 = np.linspace(0, 1, 3), np.linspace(0, 1.1, 3), np.linspace(0, 1.3, 4)
    slab = make_slab(x1, y1, z1, 3, 0, 1)
    bnd = [1, 2, 3, 4, 5, 6]
    full = SlabAssembledOperator(ctx, slab, 4, "bench_diffusion3d", bnd)
    cond = SlabAssembledOperator(ctx, slab, 4, "bench_diffusion3d", bnd, condensed=True)
    xf, res_f, it_f = full.solve(tol=1e-11)
    xc, res_c, it_c = cond.solve(tol=1e-11)
    assert res_f <= 1e-11 and res_c <= 1e-11 and cond.n_local_dofs < full.n_local_dofs
    assert rel_err(cond.recover(xc), xf.cpu().numpy()) < 1e-9
