"""CPU-only tests of the product's host logic against the oracle's restatement of the reference algorithms, and of the
C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import re
import os

import numpy as np
import pytest

import l3ster_b200 as l3b
from common import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "l3ster_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(l3b_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(l3b.EXPORTED_SYMBOLS)
    L = l3b.lib()
    for sym in declared:
        assert hasattr(L, sym), sym


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(l3b.L3BError, match="no CPU fallback"):
        l3b.Context(0)


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "l3ster_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("the oracle", "").replace("oracle's", "").replace(
                    "oracle/kernels.cpp", "").replace("against the oracle", ""), f


@pytest.mark.parametrize("dim,n,p", [(3, 3, 4), (3, 2, 2), (3, 4, 3), (2, 4, 2), (2, 5, 4), (3, 3, 1), (2, 3, 1), (3, 2, 5), (3, 2, 8), (2, 3, 7)])
def test_node_numbering_boundaries_and_sparsity_are_bit_exact(dim, n, p):
    """mesh/ConvertMeshToOrder.hpp:52-104 node ids, mesh/MeshPartition.hpp:505-596 boundary sides,
    algsys/SparsityGraph.hpp:25-81 + 254-278 graph — integer contracts, compared bit for bit."""
    orc = oracle()
    x = np.linspace(0, 1, n + 1)
    y = np.linspace(0, 2, n + 2)
    z = np.linspace(-1, 1, n)
    if dim == 3:
        pm, om = l3b.make_cube_mesh(x, y, z, order=p), orc.mesh_cube(x, y, z, order=p)
    else:
        pm, om = l3b.make_square_mesh(x, y, order=p), orc.mesh_square(x, y, order=p)
    assert pm.n_nodes == om.n_nodes and pm.n_elems == om.n_elems
    assert np.array_equal(pm.nodes.astype(np.uint64), om.elem_nodes)
    assert np.array_equal(pm.verts, om.elem_verts)
    sb = pm.side_boundaries
    for b in range(om.n_boundary):
        assert sb[om.bnd_parent[b], om.bnd_side[b]] == om.bnd_domain[b]
    assert int((sb != l3b.NO_BOUNDARY).sum()) == om.n_boundary
    if (p + 1) ** dim * n**dim > 60000:
        return
    U = 3
    ptr, nbr = pm.node_graph()
    row_ptr, col_ind = l3b.expand_graph(ptr, nbr, U)
    asys = om.assembled_system(U)
    assert np.array_equal(row_ptr, asys.row_ptr) and np.array_equal(col_ind, asys.col_ind)
    # rows sorted by local column id (SparsityGraph.hpp:275)
    for r in (0, len(row_ptr) // 2, len(row_ptr) - 2):
        assert np.all(np.diff(col_ind[row_ptr[r]:row_ptr[r + 1]]) > 0)


def test_mesh_tests_node_counts():
    """tests/MeshTests.cpp:261-279: order-p node count (n p + 1)^3, ids contiguous; interior nodes of each element form one
    contiguous ascending run (mesh/LocalMeshView.hpp:219-222)."""
    for n, p in ((3, 4), (2, 6)):
        m = l3b.make_cube_mesh(np.linspace(0, 1, n + 1), order=p)
        assert m.n_nodes == (n * p + 1) ** 3
        assert np.array_equal(np.unique(m.nodes), np.arange(m.n_nodes))
        nb = p + 1
        a = np.arange(nb**3)
        i, j, k = a % nb, (a // nb) % nb, a // nb**2
        interior = (i > 0) & (i < p) & (j > 0) & (j < p) & (k > 0) & (k < p)
        ids = m.nodes[:, interior].astype(np.int64)
        assert np.all(np.diff(ids, axis=1) == 1)


def test_structured_row_lengths():
    """structured hex p=4, U=4 row lengths 500 / 900 / 1620 / 2916 (interior / face / edge / vertex)."""
    m = l3b.make_cube_mesh(np.linspace(0, 1, 4), order=4)
    ptr, nbr = m.node_graph()
    deg = np.diff(ptr) * 4
    assert set(np.unique(deg)) >= {500, 900, 1620, 2916}
    # the centre vertex of the 3x3x3 mesh touches 8 elements
    assert deg.max() == 2916


@pytest.mark.parametrize("p", range(1, 9))
def test_tables_agree_with_reference_algorithm(p):
    """Product tables (Newton + product-form Lagrange in long double) vs the oracle's restatement of the reference
    (monomial coefficients + Horner, math/LagrangeInterpolation.hpp:12-40 — 'accurate until N ~ 16')."""
    orc = oracle()
    assert np.abs(l3b.tables_gll(p + 1) - orc.lobatto(p + 1)).max() < 1e-15
    for nq in (p + 1, p + 2, 2 * p + 1):
        gp, gw = l3b.tables_gauss(nq)
        op, ow = orc.gauss(nq)
        assert np.abs(gp - op).max() < 1e-15 and np.abs(gw - ow).max() < 1e-15
        a, b, c = l3b.tables_1d(p, nq)
        oa, ob = orc.sumfact_tables(p, 2 * (nq - 1))
        tol = 1e-15 * 10 ** max(0, p - 3) * 10  # monomial/Horner conditioning of the reference grows with p
        assert np.abs(a - oa).max() < tol
        assert np.abs(b - ob).max() < tol * np.abs(ob).max()
        # collocation derivative: D_q applied to interpolated values reproduces the interpolated derivative (nq >= nb)
        assert np.abs(a @ c - b).max() < 1e-12 * np.abs(b).max()


@pytest.mark.parametrize("dim,p,nq", [(2, 2, 3), (3, 2, 3), (3, 4, 5), (2, 4, 9)])
def test_dense_side_tables_match_reference_side_quadratures(dim, p, nq):
    """basisfun/ReferenceElementBasisAtQuadrature.hpp:56-96 with mapping/ReferenceBoundaryToSideMapping.hpp:14-48: same
    point sets per side (order may differ — every consumer sums over points), same weights, same basis values."""
    orc = oracle()
    qo = 2 * (nq - 1)
    for side in range(-1, 2 * dim):
        pts, wts, vals, ders = l3b.tables_dense(dim, p, nq, side)
        opts_, owts, ovals, oders = orc.ref_basis_at_quad(dim, p, qo, side)
        key = lambda P: np.lexsort(np.round(P, 12).T[::-1])
        i, j = key(pts), key(opts_)
        assert np.abs(pts[i] - opts_[j]).max() < 1e-15
        assert np.abs(wts[i] - owts[j]).max() < 1e-15
        assert np.abs(vals[i] - ovals[j]).max() < 1e-13
        assert np.abs(ders[i] - oders[j]).max() < 1e-12


def _check_apply_plan(n_nodes, nodes, items, up, down, n_owned=None, halo_nodes=(), block=1):
    """the invariants of csrc/apply_plan_host.hpp: every node is copied in exactly once, not later than the first item that touches it;
    copied out exactly once, not before the last item touching it and not before it came in"""
    k = len(items)
    came_in, went_out = np.full(n_nodes, -1), np.full(n_nodes, -1)
    for i in range(k):
        for a, b in up[i]:
            assert 0 <= a < b <= n_nodes and (np.all(came_in[a:b] == -1))
            came_in[a:b] = i
        for a, b in down[i]:
            assert 0 <= a < b <= n_nodes and (np.all(went_out[a:b] == -1))
            went_out[a:b] = i
    assert np.all(came_in >= 0) and np.all(went_out >= came_in)
    first, last = np.full(n_nodes, k), np.full(n_nodes, -1)
    for i, (e0, e1) in enumerate(items):
        touched = np.unique(nodes[e0:e1])
        if i == k - 1 and n_owned is not None and (n_owned < n_nodes or len(halo_nodes) or (e0 == 0 and e1 > 0)):
            touched = np.unique(np.concatenate([touched, np.asarray(halo_nodes, dtype=np.int64), np.arange(n_owned, n_nodes)]))
        first[touched] = np.minimum(first[touched], i)
        last[touched] = np.maximum(last[touched], i)
    seen = last >= 0
    assert np.all(came_in[seen] <= first[seen]) and np.all(went_out[seen] >= last[seen])
    # block granularity: every range starts on a block boundary
    for i in range(k):
        for a, b in up[i] + down[i]:
            assert a % block == 0 and (b % block == 0 or b == n_nodes)
    return came_in, went_out


@pytest.mark.parametrize("order,n,chunk,block", [(4, 4, 8, 64), (2, 5, 7, 1), (3, 3, 100, 16), (1, 6, 1, 5)])
def test_host_apply_plan_streams_a_structured_cube(order, n, chunk, block):
    host = l3b.make_cube_mesh(np.linspace(0, 1, n + 1), order=order)
    nodes = host.nodes.astype(np.int64)
    items, up, down = l3b.host_apply_plan(host.n_nodes, host.nodes, chunk, block)
    assert len(items) == -(-host.n_elems // chunk) and items[0][0] == 0 and items[-1][1] == host.n_elems
    assert np.array_equal(items[1:, 0], items[:-1, 1])
    came_in, went_out = _check_apply_plan(host.n_nodes, nodes, items, up, down, block=block)
    if len(items) >= 4:
        # the generator's numbering is monotone along z: the copies really are spread over the items (this is what makes the call overlap)
        spread = min(len(items), -(-host.n_nodes // block)) // 4
        assert len(np.unique(came_in)) >= spread and len(np.unique(went_out)) >= spread


def test_host_apply_plan_survives_a_numbering_without_locality():
    host = l3b.make_cube_mesh(np.linspace(0, 1, 4), order=3)
    rng = np.random.default_rng(5)
    perm = rng.permutation(host.n_nodes)
    nodes = perm[host.nodes.astype(np.int64)]
    order = rng.permutation(host.n_elems)
    nodes = nodes[order]
    items, up, down = l3b.host_apply_plan(host.n_nodes, nodes, 4, 32)
    _check_apply_plan(host.n_nodes, nodes, items, up, down, block=32)


def test_host_apply_plan_with_border_elements_and_halo_nodes():
    # a z-slab in the middle of a chain: the lowest node plane is owned and sent to the rank below, the top plane is ghost
    host = l3b.make_cube_mesh(np.linspace(0, 1, 4), order=2)
    nodes = host.nodes.astype(np.int64)
    zmax = host.verts[..., 2].max()
    border = np.where(np.isclose(host.verts[..., 2].max(axis=1), zmax))[0]
    ghost_mask = np.zeros(host.n_nodes, dtype=bool)
    # ghost = nodes on the top plane; found through the elements' local top layer (local nodes 18..26 of a p=2 hex)
    ghost_mask[np.unique(nodes[border][:, 18:])] = True
    new_id = np.empty(host.n_nodes, dtype=np.int64)
    n_owned = int((~ghost_mask).sum())
    new_id[~ghost_mask] = np.arange(n_owned)
    new_id[ghost_mask] = n_owned + np.arange(int(ghost_mask.sum()))
    interior = np.setdiff1d(np.arange(host.n_elems), border)
    local = new_id[nodes[np.concatenate([border, interior])]]
    zmin_elems = np.where(np.isclose(host.verts[..., 2].min(axis=1), 0.0))[0]
    halo_nodes = np.unique(new_id[nodes[zmin_elems][:, :9]])
    items, up, down = l3b.host_apply_plan(host.n_nodes, local, 5, 8, n_owned_nodes=n_owned, n_border_elems=len(border), halo_nodes=halo_nodes)
    assert tuple(items[-1]) == (0, len(border)) and items[0][0] == len(border) and items[-2][1] == host.n_elems
    came_in, went_out = _check_apply_plan(host.n_nodes, local, items, up, down, n_owned=n_owned, halo_nodes=halo_nodes, block=8)
    last = len(items) - 1
    # what the exchange reads or writes leaves with the last item only
    assert np.all(went_out[halo_nodes] == last) and np.all(went_out[n_owned:] == last)


def test_host_apply_plan_of_an_empty_rank_and_bad_input():
    items, up, down = l3b.host_apply_plan(10, np.zeros((0, 8), dtype=np.uint32), 4, 4)
    assert len(items) == 1 and tuple(items[0]) == (0, 0) and up[0] == [(0, 10)] and down[0] == [(0, 10)]
    with pytest.raises(l3b.L3BError, match="outside the local node range"):
        l3b.host_apply_plan(4, np.array([[0, 1, 2, 9]], dtype=np.uint32), 1, 2)
    with pytest.raises(l3b.L3BError, match="chunk size"):
        l3b.host_apply_plan(4, np.array([[0, 1, 2, 3]], dtype=np.uint32), 0, 2)
