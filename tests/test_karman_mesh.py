"""The real mesh of the reference's examples/07-karman-2D (BASELINE configs[3]).

CPU: the Gmsh reader on the real file (when /root/reference is present) against the committed fixture, and the order-4 conversion
against the oracle's restatement of convertMeshToOrder's geometric node matching (mesh/ConvertMeshToOrder.hpp:52-104) — node numbering
and boundary matching bit-exact. GPU: assembly of the steady + outlet kernels on that mesh against the oracle, and the solve."""
import os

import numpy as np
import pytest

import l3ster_b200 as l3b
from l3ster_b200 import meshio
import karman_common as kc
from common import oracle, rel_err

REAL = "/root/reference/examples/07-karman-2D/karman.msh"


def test_fixture_is_what_the_reader_makes_of_the_real_file():
    if not os.path.exists(REAL):
        pytest.skip("the reference tree is not present on this machine")
    m, f = meshio.read_gmsh(REAL, [kc.INLET, kc.WALL, kc.OUTLET]), kc.load_order1()
    for name in ("coords", "elems", "elem_ids", "elem_domains", "bnd_elems", "bnd_ids", "bnd_domains"):
        assert np.array_equal(getattr(m, name), getattr(f, name)), name
    assert m.dim == 2 and len(m.elems) == 914 and len(m.coords) == 996 and np.unique(m.elem_domains).tolist() == [kc.DOMAIN]


def test_order4_conversion_of_the_real_mesh_is_bit_exact():
    m = kc.load_order1()
    host = meshio.convert_to_order(m, kc.P)
    om = oracle().mesh_from_arrays(2, m.coords, m.elems, m.bnd_elems, m.bnd_domains, m.bnd_ids, kc.P)
    assert host.n_nodes == om.n_nodes == 14952
    assert np.array_equal(host.nodes.astype(np.uint64), om.elem_nodes)
    assert np.array_equal(host.verts, om.elem_verts)
    # boundary matching: every boundary line found its parent element and side (mesh/MeshPartition.hpp:505-596)
    sb = np.full((host.n_elems, 4), l3b.NO_BOUNDARY, dtype=np.uint16)
    sb[om.bnd_parent, om.bnd_side] = om.bnd_domain
    assert np.array_equal(sb, host.side_boundaries)
    assert sorted(np.unique(host.side_boundaries).tolist()) == [kc.INLET, kc.WALL, kc.OUTLET, l3b.NO_BOUNDARY]
    # all elements have positive Jacobians at their vertices (the reader flips clockwise quads)
    v = host.verts
    for a, b, c in ((0, 1, 2), (1, 3, 0), (2, 0, 3), (3, 2, 1)):
        cross = (v[:, b, 0] - v[:, a, 0]) * (v[:, c, 1] - v[:, a, 1]) - (v[:, b, 1] - v[:, a, 1]) * (v[:, c, 0] - v[:, a, 0])
        assert (cross > 0).all()


@pytest.mark.gpu
def test_karman_steady_assembled_on_the_real_mesh_matches_the_oracle():
    """assembleProblem(kernel_steady) + assembleProblem(kernel_outlet, {outlet}) + Dirichlet u, v on inlet and wall
    (source.cpp:194-214): CRS graph bit-exact, values and rhs to 1e-12 against the oracle; then the GMRES driver on that matrix"""
    import scipy.sparse as sp

    ctx = l3b.Context(0)
    m = kc.load_order1()
    host = meshio.convert_to_order(m, kc.P)
    om = oracle().mesh_from_arrays(2, m.coords, m.elems, m.bnd_elems, m.bnd_domains, m.bnd_ids, kc.P)
    xy = kc.node_coords(host)
    fdata = kc.previous_velocity(xy)
    mesh = ctx.upload_mesh(host)
    s = l3b.AssembledSystem(ctx, mesh, kc.U, 1, host.node_graph())
    so = om.assembled_system(kc.U)
    row_ptr, col_ind = s.graph()
    assert np.array_equal(row_ptr, so.row_ptr) and np.array_equal(col_ind, so.col_ind)
    fields = ctx.upload_fields(fdata)
    s.beginAssembly()
    for k in kc.KERNELS:
        s.assembleProblem(k["name"], k.get("boundary_ids", ()), fields if "field_inds" in k else None, k.get("field_inds"), k.get("dof_inds"),
                          k["asm_opts"])
        so.assemble_ex(k["name"], 1, 1, 0.0, fdata if "field_inds" in k else None, n_threads=4, boundary_ids=k.get("boundary_ids", ()),
                       dof_inds=k.get("dof_inds"), field_inds=k.get("field_inds"))
    v_g, r_g = s.download()
    v_o, r_o = so.get()
    # the rhs entries u.grad(u) tested against B cancel by four orders of magnitude on this mesh (h ~ 0.01): the error bar is 1e-12 of
    # the scale of the summands, |F_i| <~ sqrt(K_ii) |f|, not of the cancelled result
    import scipy.sparse as sp_

    rhs_scale = np.linalg.norm(np.sqrt(np.abs(sp_.csr_matrix((v_o, col_ind, row_ptr), shape=(s.n_dofs,) * 2).diagonal())))
    assert rel_err(v_g, v_o) < 1e-12 and np.linalg.norm(r_g - r_o) < 1e-12 * max(np.linalg.norm(r_o), rhs_scale)
    dofs, vals = kc.dirichlet(host.boundary_nodes([kc.WALL]), host.boundary_nodes([kc.INLET]), xy)
    s.endAssembly(dofs.astype(np.int32), vals)
    so.apply_dirichlet(dofs.astype(np.int32), vals)
    v_g, r_g = s.download()
    v_o, r_o = so.get()
    assert rel_err(v_g, v_o) < 1e-12 and np.linalg.norm(r_g - r_o) < 1e-12 * max(np.linalg.norm(r_o), rhs_scale)
    # The example solves this system with the direct solver KLU2 (source.cpp:189): Jacobi-preconditioned GMRES(250) stagnates on it
    # (checked with scipy: no convergence in 10 000 iterations), so the Krylov layer is checked for CONSISTENCY here — after a fixed
    # number of iterations the residual it reports is the true preconditioned residual of the iterate it returns, and it went down.
    A = sp.csr_matrix((v_o, col_ind, row_ptr), shape=(s.n_dofs,) * 2)
    d = A.diagonal()
    x, res, its = s.solve_gmres(tol=1e-12, restart_length=250, max_restarts=39, max_iters=300)
    assert its == 300
    true_res = np.linalg.norm((r_o[:, 0] - A @ x) / d)
    assert abs(true_res - res) < 1e-6 * true_res
    assert res < 0.5 * np.linalg.norm(r_o[:, 0] / d)
    assert np.abs(x[kc.IU::kc.U]).max() > 0.5  # a flow, not the trivial solution
