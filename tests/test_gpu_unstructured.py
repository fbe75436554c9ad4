"""GPU parity on an unstructured mesh read from a Gmsh file: three p=4 quads around a valence-3 vertex with rotated local axes —
assembly (domain + boundary kernel), sparsity and the matrix-free apply against the oracle on the mesh the oracle converts itself."""
import numpy as np
import pytest

import l3ster_b200 as l3b
from l3ster_b200 import meshio
from common import oracle, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return l3b.Context(0)


def _hexagon(tmp_path, order):
    ang = np.deg2rad(np.arange(6) * 60.0)
    coords = np.zeros((7, 3))
    coords[:6, 0], coords[:6, 1] = 1.3 * np.cos(ang), np.sin(ang) + 0.1 * np.cos(2 * ang)
    coords[6, :2] = 0.07, -0.05
    ccw = [[6, 0, 1, 2], [2, 3, 4, 6], [4, 5, 0, 6]]  # each quad starts at a different corner: rotated local axes
    lex = np.array([[q[0], q[1], q[3], q[2]] for q in ccw])
    bnd = np.array([[i, (i + 1) % 6] for i in range(6)])
    m1 = meshio.Order1Mesh(2, coords, lex, np.arange(3), np.full(3, 7), bnd, 3 + np.arange(6), np.array([8, 8, 9, 9, 8, 8]))
    f = str(tmp_path / "hexagon.msh")
    meshio.write_gmsh(f, m1)
    r = meshio.read_gmsh(f, [8, 9])
    host = meshio.convert_to_order(r, order)
    orc = oracle().mesh_from_arrays(2, r.coords, r.elems, r.bnd_elems, r.bnd_domains, r.bnd_ids, order)
    assert np.array_equal(orc.elem_nodes, host.nodes.astype(np.uint64))
    return host, orc


def test_assembly_and_apply_on_a_gmsh_mesh(ctx, tmp_path):
    host, orc = _hexagon(tmp_path, 4)
    mesh = ctx.upload_mesh(host)
    U = 3
    s = l3b.AssembledSystem(ctx, mesh, U)
    s.beginAssembly()
    s.assembleProblem("example02_domain")
    s.assembleProblem("example02_bc", boundary_ids=[8])
    so = orc.assembled_system(U)
    so.assemble("example02_domain")
    so.assemble("example02_bc", boundary_ids=[8])
    v_g, r_g = s.download()
    v_o, r_o = so.get()
    assert rel_err(v_g, v_o) < 1e-12 and rel_err(r_g, r_o) < 1e-12
    # matrix-free operator of the same problem with Dirichlet rows on boundary 9
    mask = np.zeros(host.n_nodes * U, dtype=np.uint8)
    mask[host.boundary_nodes([9]) * U] = 1
    mf = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    mf.assembleProblem("example02_domain")
    mf.assembleProblem("example02_bc", boundary_ids=[8])
    mf.endAssembly()
    mo = orc.matrix_free_system(U, 1, mask, None)
    mo.add_kernel("example02_domain")
    mo.add_kernel("example02_bc", boundary_ids=[8])
    d_o, f_o = mo.init()
    d_g, f_g = mf.download()
    assert rel_err(d_g, d_o) < 1e-12
    x = np.random.default_rng(2).uniform(-1, 1, size=(host.n_nodes * U, 1))
    assert rel_err(mf.apply(x), mo.apply(x)) < 1e-12
