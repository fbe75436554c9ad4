"""Mesh front end for unstructured meshes (l3ster_b200/meshio.py): the Gmsh reader's conventions (mesh/ReadMesh.hpp) and the order-p
node numbering (mesh/ConvertMeshToOrder.hpp) against the structured generators and against the oracle's restatement of the reference's
geometric matching, on meshes whose elements are rotated against each other and meet three at a vertex."""
import numpy as np
import pytest

import l3ster_b200 as l3b
from l3ster_b200 import meshio
from common import oracle


def _roundtrip(tmp_path, m1, bids):
    f = str(tmp_path / "mesh.msh")
    meshio.write_gmsh(f, m1)
    return meshio.read_gmsh(f, bids)


@pytest.mark.parametrize("dim,n,p", [(2, 3, 3), (2, 4, 4), (3, 2, 2), (3, 2, 4)])
def test_gmsh_roundtrip_reproduces_the_structured_numbering(tmp_path, dim, n, p):
    d = np.linspace(0.0, 1.0, n + 1)
    make = (lambda o: l3b.make_square_mesh(d, d, order=o)) if dim == 2 else (lambda o: l3b.make_cube_mesh(d, d, d, order=o))
    h1, hp = make(1), make(p)
    m1 = meshio.order1_from_host(h1)
    r = _roundtrip(tmp_path, m1, sorted(set(int(b) for b in m1.bnd_domains)))
    assert np.array_equal(r.elems, m1.elems) and np.array_equal(r.bnd_elems, m1.bnd_elems)
    u = meshio.convert_to_order(r, p)
    assert u.n_nodes == hp.n_nodes
    assert np.array_equal(u.nodes, hp.nodes)
    assert np.array_equal(u.side_boundaries, hp.side_boundaries)
    assert np.array_equal(u.verts, hp.verts)


def _rotate_elements(m1, rng):
    """the same mesh with the local axes of some elements rotated (even permutations: positive Jacobians), so that neighbours see their
    shared edges and faces with different orientations"""
    elems = m1.elems.copy()
    for e in range(len(elems)):
        k = rng.integers(0, 3)
        v = elems[e].copy()
        if m1.dim == 2:
            for _ in range(k):  # quarter turn: new (0,1,2,3) = old (1,3,0,2)
                v = v[[1, 3, 0, 2]]
        else:
            for _ in range(k):  # cyclic permutation of the axes: new vertex (bx, by, bz) = old vertex with (x, y, z) bits (bz, bx, by)
                v = np.array([v[((i >> 2) & 1) + 2 * (i & 1) + 4 * ((i >> 1) & 1)] for i in range(8)])
        elems[e] = v
    return meshio.Order1Mesh(m1.dim, m1.coords, elems, m1.elem_ids, m1.elem_domains, m1.bnd_elems, m1.bnd_ids, m1.bnd_domains)


@pytest.mark.parametrize("dim,n,p", [(2, 3, 2), (2, 3, 4), (3, 2, 2), (3, 2, 3)])
def test_order_conversion_matches_the_reference_algorithm_on_rotated_elements(tmp_path, dim, n, p):
    d = np.linspace(0.0, 1.0, n + 1)
    h1 = l3b.make_square_mesh(d, d, order=1) if dim == 2 else l3b.make_cube_mesh(d, d, d, order=1)
    m1 = _rotate_elements(meshio.order1_from_host(h1), np.random.default_rng(4))
    u = meshio.convert_to_order(m1, p)
    o = oracle().mesh_from_arrays(dim, m1.coords, m1.elems, m1.bnd_elems, m1.bnd_domains, m1.bnd_ids, p)
    assert o.n_nodes == u.n_nodes
    assert np.array_equal(o.elem_nodes, u.nodes.astype(np.uint64))
    # sides carrying boundary ids: the oracle's (parent, side, domain) triples
    got = {(e, s): int(u.side_boundaries[e, s]) for e in range(u.n_elems) for s in range(2 * dim) if u.side_boundaries[e, s] != meshio.NO_BOUNDARY}
    ref = {(int(o.bnd_parent[b]), int(o.bnd_side[b])): int(o.bnd_domain[b]) for b in range(o.n_boundary)}
    assert got == ref


def test_three_quads_around_a_vertex_and_a_flipped_element(tmp_path):
    """an unstructured patch: a hexagon cut into three quads meeting at its centre (vertex valence 3); one quad is written clockwise
    in the file and must come back flipped (ReadMesh.hpp:70-90)"""
    ang = np.deg2rad(np.arange(6) * 60.0)
    coords = np.zeros((7, 3))
    coords[:6, 0], coords[:6, 1] = np.cos(ang), np.sin(ang)
    ccw = [[6, 0, 1, 2], [6, 2, 3, 4], [6, 4, 5, 0]]  # gmsh order (counter-clockwise)
    lex = np.array([[q[0], q[1], q[3], q[2]] for q in ccw])  # reference order
    bnd = np.array([[i, (i + 1) % 6] for i in range(6)])
    m1 = meshio.Order1Mesh(2, coords, lex, np.arange(3), np.full(3, 7), bnd, 3 + np.arange(6), np.array([8, 8, 9, 9, 8, 8]))
    f = str(tmp_path / "hexagon.msh")
    meshio.write_gmsh(f, m1)
    txt = open(f).read().split("\n")
    k = next(i for i, l in enumerate(txt) if l.split()[:1] == ["2"] and len(l.split()) == 5 and l.split()[0] == "2")  # second quad: reverse it
    w = txt[k].split()
    txt[k] = " ".join([w[0], w[1], w[4], w[3], w[2]])
    open(f, "w").write("\n".join(txt))
    r = meshio.read_gmsh(f, [8, 9])
    c = r.coords[r.elems[1]]
    assert np.cross(c[1] - c[0], c[2] - c[0])[2] > 0.0  # positive Jacobian after the flip
    u = meshio.convert_to_order(r, 3)
    o = oracle().mesh_from_arrays(2, r.coords, r.elems, r.bnd_elems, r.bnd_domains, r.bnd_ids, 3)
    assert np.array_equal(o.elem_nodes, u.nodes.astype(np.uint64))
    assert u.n_nodes == 7 + 9 * 2 + 3 * 4  # 7 vertices, 9 edges x 2, 3 interiors x 4
    assert sorted(set(u.side_boundaries.ravel().tolist())) == [8, 9, meshio.NO_BOUNDARY]


def test_reader_rejects_what_the_reference_rejects(tmp_path):
    f = tmp_path / "bad.msh"
    f.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError, match="Only the ASCII v4"):
        meshio.read_gmsh(str(f), [])
    f.write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError, match="required section"):
        meshio.read_gmsh(str(f), [])
