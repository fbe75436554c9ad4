"""DOF maps with inactive dofs and several domains (l3b_dofmap_*), against restatements of the reference's own tests.

CPU: tests/SparsityGraphTest.cpp:99-146 — cube, order 2, definitions ({0},{1}), ({1},{0}), ({3},{0,1}) (volume domain 0, boundary
faces 1 and 3): numbering (dofs/NodeToDofMap.hpp:248-264) and graph against the brute-force DenseGraph of that test (:14-48).
GPU: tests/MultiDomainTest.cpp — a square with four volume domains, dof i active in domain i only, kernel A0 = 1, rhs = i + 1: compact
matrix against the oracle, assembled and matrix-free solves return the prescribed value per domain."""
import numpy as np
import pytest

import l3ster_b200 as l3b


def dense_graph(n_dofs, dof, cliques):
    """DenseGraph of tests/SparsityGraphTest.cpp:14-48: for every (entity, definition) all (row, col) pairs of its dofs"""
    g = np.zeros((n_dofs, n_dofs), dtype=bool)
    for nodes, dofs in cliques:
        ids = dof[np.asarray(nodes, dtype=np.int64)][:, dofs].ravel()
        g[np.ix_(ids, ids)] = True
    return g


def cliques_of(host, elem_domains, definitions):
    out = []
    for doms, dofs in definitions:
        for e in range(host.n_elems):
            if elem_domains[e] in doms:
                out.append((host.nodes[e], list(dofs)))
            for s in range(host.n_sides):
                if host.side_boundaries[e, s] != l3b.NO_BOUNDARY and host.side_boundaries[e, s] in doms:
                    out.append((host.nodes[e][l3b.side_node_inds(host.dim, host.order, s)], list(dofs)))
    return out


def test_sparsity_graph_test_restated():
    host = l3b.make_cube_mesh(np.linspace(0.0, 1.0, 5), order=2)
    ed = np.zeros(host.n_elems, dtype=np.int32)  # mesh/primitives/CubeMesh.hpp: volume domain 0, faces 1..6
    defs = [([0], [1]), ([1], [0]), ([3], [0, 1])]
    dm = l3b.DofMap(3, 2, host.n_nodes, host.nodes, ed, host.side_boundaries, 2, defs)
    # activity: dof 1 everywhere (domain 0 holds every node), dof 0 on the nodes of faces 1 and 3
    on_faces = np.zeros(host.n_nodes, dtype=bool)
    on_faces[host.boundary_nodes([1, 3])] = True
    assert dm.active[:, 1].all() and np.array_equal(dm.active[:, 0], on_faces)
    # numbering: node-major over the active pairs
    expect = np.full((host.n_nodes, 2), -1, dtype=np.int64)
    expect[dm.active] = np.arange(int(dm.active.sum()))
    assert np.array_equal(dm.dof, expect) and dm.n_dofs == int(dm.active.sum())
    dense = dense_graph(dm.n_dofs, dm.dof, cliques_of(host, ed, defs))
    for r in range(dm.n_dofs):
        cols = dm.col_ind[dm.row_ptr[r]:dm.row_ptr[r + 1]]
        assert (np.diff(cols) > 0).all()
        assert np.array_equal(np.nonzero(dense[r])[0], cols)


def test_full_problem_definition_reduces_to_the_node_block_graph():
    """all dofs active on one domain: the compact graph is the dpn-fold expansion of the node graph (what l3b_graph_expand returns)"""
    host = l3b.make_square_mesh(np.linspace(0.0, 1.0, 4), order=3)
    dm = l3b.DofMap(2, 3, host.n_nodes, host.nodes, None, host.side_boundaries, 3, [([0], [0, 1, 2])])
    row_ptr, col_ind = l3b.expand_graph(*host.node_graph(), 3)
    assert dm.n_dofs == host.n_nodes * 3 and np.array_equal(dm.row_ptr, row_ptr) and np.array_equal(dm.col_ind, col_ind)
    assert np.array_equal(dm.dof.ravel(), np.arange(dm.n_dofs))


def _multidomain_square(order):
    """a square of 4 x 4 elements whose quadrants are the domains 13..16 (the role of tests/data/gmsh_ascii4_square_multidom.msh)"""
    host = l3b.make_square_mesh(np.linspace(0.0, 1.0, 5), order=order)
    c = host.verts.mean(axis=1)
    ed = (13 + (c[:, 0] > 0.5) + 2 * (c[:, 1] > 0.5)).astype(np.int32)
    return host, ed


def test_multidomain_dofmap_numbering():
    host, ed = _multidomain_square(2)
    defs = [([13 + i], [i]) for i in range(4)]
    dm = l3b.DofMap(2, 2, host.n_nodes, host.nodes, ed, None, 4, defs)
    for i in range(4):
        nodes_i = np.unique(host.nodes[ed == 13 + i])
        expect = np.zeros(host.n_nodes, dtype=bool)
        expect[nodes_i] = True
        assert np.array_equal(dm.active[:, i], expect)
    dense = dense_graph(dm.n_dofs, dm.dof, cliques_of(host, ed, defs))
    for r in range(dm.n_dofs):
        assert np.array_equal(np.nonzero(dense[r])[0], dm.col_ind[dm.row_ptr[r]:dm.row_ptr[r + 1]])
    # the nodes on the lines between quadrants carry two dofs, the centre node four
    assert dm.active.sum(axis=1).max() == 4 and dm.active.sum(axis=1).min() == 1


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", [0, 1], ids=["sumfact", "local_element"])
def test_multi_domain_test_restated(strategy):
    """tests/MultiDomainTest.cpp:15-103: unknown i lives in domain 13 + i only; kernel A0(0, 0) = 1, rhs = i + 1 assembled per domain
    with dof_inds = {i}; CG to 1e-10; the solution of dof i is i + 1 on its domain's nodes. Assembled (compact matrix against the
    oracle's per-domain assembly) and matrix-free."""
    from common import oracle, rel_err

    ctx = l3b.Context(0)
    p = 2
    host, ed = _multidomain_square(p)
    defs = [([13 + i], [i]) for i in range(4)]
    dm = l3b.DofMap(2, p, host.n_nodes, host.nodes, ed, None, 4, defs)
    mesh = ctx.upload_mesh(host)
    mesh.set_element_domains(ed)
    opts = l3b.AssemblyOptions(1, 0, strategy)
    # assembled
    a = l3b.AssembledSystem(ctx, mesh, 4)
    a.set_dofmap(dm)
    a.beginAssembly()
    for i in range(4):
        a.assembleProblem("multidomain_mass", boundary_ids=[13 + i], dof_inds=[i], time=float(i + 1))
    vals, rhs = a.download_compact(dm)
    # oracle: the same kernel assembled element by element into the compact graph
    import scipy.sparse as sp

    orc = oracle()
    A_o = sp.lil_matrix((dm.n_dofs, dm.n_dofs))
    b_o = np.zeros(dm.n_dofs)
    for e in range(host.n_elems):
        i = int(ed[e]) - 13
        K, F = orc.assemble_local("multidomain_mass", 2, p, host.verts[e], time=float(i + 1))
        ids = dm.dof[host.nodes[e].astype(np.int64), i]
        A_o[np.ix_(ids, ids)] += K
        b_o[ids] += F[:, 0]
    A_o = A_o.tocsr()
    A_o.sort_indices()
    assert np.array_equal(A_o.indptr, dm.row_ptr) and np.array_equal(A_o.indices, dm.col_ind)
    assert rel_err(vals, A_o.data) < 1e-12 and rel_err(rhs[:, 0], b_o) < 1e-12
    a.endAssembly()
    x, res, its = a.solve(tol=1e-10)
    xs = x.reshape(host.n_nodes, 4)
    for i in range(4):
        assert np.abs(xs[dm.active[:, i], i] - (i + 1)).max() < 1e-6
        assert np.abs(xs[~dm.active[:, i], i]).max() == 0.0  # closed pairs stay zero
    # matrix-free
    mask, dvals = dm.padded_dirichlet()
    mf = l3b.MatrixFreeSystem(ctx, mesh, 4, 1, mask, dvals)
    for i in range(4):
        mf.assembleProblem("multidomain_mass", boundary_ids=[13 + i], dof_inds=[i], asm_opts=opts, time=float(i + 1))
    mf.endAssembly()
    diag, rhs_mf = mf.download()
    assert rel_err(dm.compact(diag), A_o.diagonal()) < 1e-12 and rel_err(dm.compact(rhs_mf[:, 0]), b_o) < 1e-12
    xr = np.random.default_rng(0).uniform(-1, 1, size=dm.n_dofs)
    xp = np.zeros(host.n_nodes * 4)
    xp[dm.active.ravel()] = xr
    assert rel_err(dm.compact(mf.apply(xp.reshape(-1, 1))[:, 0]), A_o @ xr) < 1e-12
    xm, res_m, its_m = mf.solve(tol=1e-10)
    xs = xm.reshape(host.n_nodes, 4)
    for i in range(4):
        assert np.abs(xs[dm.active[:, i], i] - (i + 1)).max() < 1e-6
