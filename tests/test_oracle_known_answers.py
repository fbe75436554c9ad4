"""Pins the CPU oracle against every known-answer test the reference's own suite holds for the hot path.

Each test names the reference test it restates (paths relative to /root/reference). The reference ships no stored
numeric dumps for K_e / y_e, only analytic and self-consistency checks — all of them are encoded here with the
reference's own tolerances.
"""
import math

import numpy as np
import pytest

from oracle import HEX, LINE, QUAD, Oracle


@pytest.fixture(scope="module")
def orc():
    return Oracle()


# ---------------------------------------------------------------------------------------------------------------------
# tests/MathTests.cpp:146-214
def test_legendre_polynomials(orc):
    np.testing.assert_allclose(orc.legendre(2), [1.5, 0.0, -0.5], atol=1e-15)
    np.testing.assert_allclose(orc.legendre(3), [2.5, 0.0, -1.5, 0.0], atol=1e-15)
    np.testing.assert_allclose(orc.legendre(4), [4.375, 0.0, -3.75, 0.0, 0.375], atol=1e-15)


def test_lobatto_abscissas(orc):
    assert list(orc.lobatto(2)) == [-1.0, 1.0]
    assert list(orc.lobatto(3)) == [-1.0, 0.0, 1.0]
    a = 0.2 * math.sqrt(5.0)
    np.testing.assert_allclose(orc.lobatto(4), [-1, -a, a, 1], atol=1e-14, rtol=0)
    a = math.sqrt(21.0) / 7.0
    la = orc.lobatto(5)
    np.testing.assert_allclose(la, [-1, -a, 0, a, 1], atol=1e-14, rtol=0)
    assert la[2] == 0.0 and la[0] == -1.0 and la[4] == 1.0
    a14 = math.sqrt((7.0 + 2 * math.sqrt(7.0)) / 21.0)
    a23 = math.sqrt((7.0 - 2 * math.sqrt(7.0)) / 21.0)
    np.testing.assert_allclose(orc.lobatto(6), [-1, -a14, -a23, a23, a14, 1], atol=1e-14, rtol=0)


# tests/MathTests.cpp:120-144 — Lagrange interpolation reproduces its nodes' values
def test_lagrange_interpolation(orc):
    x = np.linspace(-1, 1, 7)
    y = np.sin(3 * x) + x**2
    c = orc.lagrange_interp(x, y)
    np.testing.assert_allclose(np.polyval(c, x), y, atol=5e-3)
    np.testing.assert_allclose(np.polyval(c, x), y, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# tests/QuadratureTests.cpp:10-61
def test_gauss_legendre_1d(orc):
    tol = 1e-10
    p, w = orc.gauss(1)
    assert abs(p[0]) < tol and abs(w[0] - 2) < tol
    p, w = orc.gauss(2)
    np.testing.assert_allclose(p, [-0.57735026919, 0.57735026919], atol=tol)
    np.testing.assert_allclose(w, [1, 1], atol=tol)
    p, w = orc.gauss(3)
    np.testing.assert_allclose(p, [-0.77459666924, 0, 0.77459666924], atol=tol)
    np.testing.assert_allclose(w, [0.55555555556, 0.88888888889, 0.55555555556], atol=tol)


# tests/QuadratureTests.cpp:129-219
def test_quad_quadrature_exactness(orc):
    tol = 1e-10
    funs = [
        (lambda x, y: 1.0 + 0 * x, 4.0),
        (lambda x, y: 2 * x + 3 * y + 1, 4.0),
        (lambda x, y: 2 * x * x + x + 3 * y * y + 2 * y + 1, 10.666666666667),
        (lambda x, y: 3 * x**3 + 2 * x * x + x + 4 * y**3 + 3 * y * y + 2 * y + 1, 10.666666666667),
        (lambda x, y: 4 * x**4 + 3 * x**3 + 2 * x * x + x + 5 * y**4 + 4 * y**3 + 3 * y * y + 2 * y + 1, 17.866666666667),
        (lambda x, y: 5 * x**5 + 4 * x**4 + 3 * x**3 + 2 * x * x + x + 6 * y**5 + 5 * y**4 + 4 * y**3 + 3 * y * y + 2 * y + 1,
         17.866666666667),
    ]
    p, w = orc.quadrature(QUAD, 1)
    assert len(w) == 1 and abs(w[0] - 4) < tol and np.all(np.abs(p) < tol)
    p, w = orc.quadrature(QUAD, 3)
    assert len(w) == 4
    for f, ref in funs[:4]:
        assert abs(np.sum(w * f(p[:, 0], p[:, 1])) - ref) < tol
    p, w = orc.quadrature(QUAD, 5)
    assert len(w) == 9
    for f, ref in funs:
        assert abs(np.sum(w * f(p[:, 0], p[:, 1])) - ref) < tol


# tests/QuadratureTests.cpp:222-295
def test_hex_quadrature_exactness(orc):
    tol = 1e-10
    funs = [
        (lambda x, y, z: 1.0 + 0 * x, 8.0),
        (lambda x, y, z: x * y * z + x * y + y * z - z * x + x + y + z + 1, 8.0),
        (lambda x, y, z: x * x * y * (y + 2) * z * (z + 1) + x * (y - 1) + z * y * (y - 2), 8.0 / 27.0),
        (lambda x, y, z: z * z * (z + 1) * (x * x + x) + (y + 1) * y * y, 32.0 / 9.0),
    ]
    p, w = orc.quadrature(HEX, 1)
    assert len(w) == 1 and abs(w[0] - 8) < tol
    p, w = orc.quadrature(HEX, 3)
    assert len(w) == 8
    for f, ref in funs:
        assert abs(np.sum(w * f(p[:, 0], p[:, 1], p[:, 2])) - ref) < tol
    p, w = orc.quadrature(HEX, 15)
    assert len(w) == 512
    trig = np.sin(p[:, 0]) * np.tan(p[:, 0]) + np.sin(p[:, 1]) * np.cos(p[:, 2]) ** 2
    ref = -8.0 * (math.sin(1) - 2.0 * math.atanh(math.tan(0.5)))
    assert np.sum(w * trig) == pytest.approx(ref)  # Catch2 Approx default epsilon
    # quad/GenerateQuadrature.hpp:62-75: first coordinate slowest
    g, _ = orc.gauss(8)
    assert p[1, 2] == g[1] and p[1, 0] == g[0] and p[64, 0] == g[1]


# ---------------------------------------------------------------------------------------------------------------------
LINE_EL = [[0, 0, 0], [1, 0, 0]]
QUAD_EL = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 2, 0]]
HEX_EL = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1.5], [0, 1, 1.5], [1, 1, 2]]


# tests/MappingTests.cpp:49-97
def test_reference_to_physical_mapping(orc):
    m = orc.map_to_physical(LINE, [[1, 1, 1], [0.5, 0.5, 0.5]], [0.0])
    np.testing.assert_allclose(m, [0.75] * 3, atol=1e-15)
    m = orc.map_to_physical(QUAD, [[1, -1, 0], [2, -1, 0], [1, 1, 1], [2, 1, 1]], [0.5, -0.5])
    np.testing.assert_allclose(m, [1.75, -0.5, 0.25], atol=1e-15)
    v = [[0.5, 0.5, 0.5], [1, 0.5, 0.5], [0.5, 1, 0.5], [1, 1, 0.5], [0.5, 0.5, 1], [1, 0.5, 1], [0.5, 1, 1], [1, 1, 1]]
    np.testing.assert_allclose(orc.map_to_physical(HEX, v, [0, 0, 0]), [0.75] * 3, atol=1e-15)


# tests/MappingTests.cpp:99-135
def test_jacobi_matrix(orc):
    assert orc.jacobi_mat(LINE, LINE_EL, [0.42])[0, 0] == pytest.approx(0.5, abs=1e-13)
    np.testing.assert_allclose(orc.jacobi_mat(QUAD, QUAD_EL, [0.5, 0.5]), [[7 / 8, 3 / 8], [3 / 8, 7 / 8]], atol=1e-13)
    np.testing.assert_allclose(orc.jacobi_mat(HEX, HEX_EL, [0.5, 0.5, 0.5]), [[0.5, 0, 3 / 16], [0, 0.5, 3 / 16], [0, 0, 7 / 8]],
                               atol=1e-13)


# tests/MappingTests.cpp:137-218
def test_boundary_normals(orc):
    n = lambda et, el, side, pt: orc.boundary_normal(et, side, orc.jacobi_mat(et, el, pt))
    assert n(LINE, LINE_EL, 0, [0.0])[0] == pytest.approx(-1, abs=1e-13)
    assert n(LINE, LINE_EL, 1, [1.0])[0] == pytest.approx(1, abs=1e-13)
    s5 = math.sqrt(5.0)
    np.testing.assert_allclose(n(QUAD, QUAD_EL, 0, [0, -1]), [0, -1], atol=1e-13)
    np.testing.assert_allclose(n(QUAD, QUAD_EL, 1, [0, 1]), [-1 / s5, 2 / s5], atol=1e-13)
    np.testing.assert_allclose(n(QUAD, QUAD_EL, 2, [-1, 0]), [-1, 0], atol=1e-13)
    np.testing.assert_allclose(n(QUAD, QUAD_EL, 3, [1, 0]), [2 / s5, -1 / s5], atol=1e-13)
    np.testing.assert_allclose(n(HEX, HEX_EL, 0, [0, 0, -1]), [0, 0, -1], atol=1e-13)
    np.testing.assert_allclose(n(HEX, HEX_EL, 1, [0, 0, 1]), [-math.sqrt(1 / 6), -math.sqrt(1 / 6), math.sqrt(2 / 3)], atol=1e-13)
    np.testing.assert_allclose(n(HEX, HEX_EL, 2, [0, -1, 0]), [0, -1, 0], atol=1e-13)
    np.testing.assert_allclose(n(HEX, HEX_EL, 3, [0, 1, 0]), [0, 1, 0], atol=1e-13)
    np.testing.assert_allclose(n(HEX, HEX_EL, 4, [-1, 0, 0]), [-1, 0, 0], atol=1e-13)
    np.testing.assert_allclose(n(HEX, HEX_EL, 5, [1, 0, 0]), [1, 0, 0], atol=1e-13)


# tests/MappingTests.cpp:220-330
def test_basis_function_values(orc):
    v = lambda et, i, pt: orc.ref_basis_value(et, 1, i, pt)
    assert v(LINE, 0, [-1.0]) == pytest.approx(1, abs=1e-13) and v(LINE, 0, [1.0]) == pytest.approx(0, abs=1e-13)
    assert v(LINE, 1, [-1.0]) == pytest.approx(0, abs=1e-13) and v(LINE, 1, [1.0]) == pytest.approx(1, abs=1e-13)
    p0, p1, p2 = [-0.5, -0.5], [0.5, 0.5], [1.0, 1.0]
    exp = {0: (0.75 * 0.75, 0.25 * 0.25, 0), 1: (0.25 * 0.75, 0.25 * 0.75, 0), 2: (0.25 * 0.75, 0.25 * 0.75, 0), 3: (0.25 * 0.25, 0.75 * 0.75, 1)}
    for i, e in exp.items():
        for p, r in zip((p0, p1, p2), e):
            assert v(QUAD, i, p) == pytest.approx(r, abs=1e-13)
    p0, p1, p2, p3 = [-0.5] * 3, [0.5] * 3, [1.0, 1.0, -1.0], [0.0, 1.0, 1.0]
    a, b = 0.25, 0.75
    exp = {
        0: (b * b * b, a * a * a, 0, 0), 1: (a * b * b, a * a * b, 0, 0), 2: (a * b * b, a * a * b, 0, 0), 3: (a * a * b, a * b * b, 1, 0),
        4: (a * b * b, a * a * b, 0, 0), 5: (a * a * b, a * b * b, 0, 0), 6: (a * a * b, a * b * b, 0, 0.5), 7: (a * a * a, b * b * b, 0, 0.5),
    }
    for i, e in exp.items():
        for p, r in zip((p0, p1, p2, p3), e):
            assert v(HEX, i, p) == pytest.approx(r, abs=1e-13)


# tests/MappingTests.cpp:332-403
def test_basis_function_derivatives(orc):
    np.testing.assert_allclose(orc.phys_basis_ders(LINE, 1, LINE_EL, [0.0]), [[-1, 1]], atol=1e-13)
    d = orc.phys_basis_ders(QUAD, 1, QUAD_EL, [0.0, 0.0])
    np.testing.assert_allclose(d, [[-0.25, 0.5, -0.5, 0.25], [-0.25, -0.5, 0.5, 0.25]], atol=1e-13)
    cube = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]]
    d = orc.phys_basis_ders(HEX, 1, cube, [0.0, 0.0, 0.0])
    sx = [-1, 1, -1, 1, -1, 1, -1, 1]
    sy = [-1, -1, 1, 1, -1, -1, 1, 1]
    sz = [-1, -1, -1, -1, 1, 1, 1, 1]
    np.testing.assert_allclose(d, 0.25 * np.array([sx, sy, sz]), atol=1e-13)


# tests/MappingTests.cpp:405-427
def test_reference_basis_at_domain_qps(orc):
    _, _, vals, ders = orc.ref_basis_at_quad(HEX, 4, 4)
    np.testing.assert_allclose(vals.sum(axis=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(ders.sum(axis=2), 0.0, atol=1e-13)


# tests/MappingTests.cpp:429-523 ("Generated" sections)
def test_reference_basis_at_boundary_qps(orc):
    node_pos = [0.0, 0.25, 0.5, 0.75, 1.0]
    for side, x in ((0, 0.0), (1, 1.0)):
        pts, _, _, _ = orc.ref_basis_at_quad(LINE, 1, 1, side)
        phys = orc.map_to_physical(LINE, [[node_pos[0], 0, 0], [node_pos[-1], 0, 0]], pts[0])
        np.testing.assert_allclose(phys, [x, 0, 0], atol=1e-15)
    # boundary ids → (axis, offset): SquareMesh.hpp ids bottom 1, top 2, left 3, right 4
    mesh = orc.mesh_square(node_pos)
    plane = {1: (1, 0.0), 2: (1, 1.0), 3: (0, 0.0), 4: (0, 1.0)}
    for b in range(mesh.n_boundary):
        axis, offs = plane[int(mesh.bnd_domain[b])]
        pts, _, _, _ = orc.ref_basis_at_quad(QUAD, 1, 5, int(mesh.bnd_side[b]))
        for qp in pts:
            assert orc.map_to_physical(QUAD, mesh.elem_verts[mesh.bnd_parent[b]], qp)[axis] == pytest.approx(offs, abs=1e-15)
    # CubeMesh.hpp ids: back 1 (z=min), front 2 (z=max), bottom 3, top 4, left 5, right 6
    mesh = orc.mesh_cube(node_pos)
    plane = {1: (2, 0.0), 2: (2, 1.0), 3: (1, 0.0), 4: (1, 1.0), 5: (0, 0.0), 6: (0, 1.0)}
    assert mesh.n_boundary == 6 * 16
    for b in range(mesh.n_boundary):
        axis, offs = plane[int(mesh.bnd_domain[b])]
        pts, _, _, _ = orc.ref_basis_at_quad(HEX, 1, 5, int(mesh.bnd_side[b]))
        for qp in pts:
            assert orc.map_to_physical(HEX, mesh.elem_verts[mesh.bnd_parent[b]], qp)[axis] == pytest.approx(offs, abs=1e-15)


# tests/MappingTests.cpp:567-605
def test_boundary_integration(orc):
    def area(et, el, side):
        pts, wts, _, _ = orc.ref_basis_at_quad(et, 1, 10, side)
        return sum(w * orc.boundary_jacobian(et, side, orc.jacobi_mat(et, el, p)) for p, w in zip(pts, wts))

    assert area(LINE, LINE_EL, 0) == 0.0 and area(LINE, LINE_EL, 1) == 0.0
    for side, ref in enumerate((1.0, math.sqrt(5.0), 1.0, math.sqrt(5.0))):
        assert area(QUAD, QUAD_EL, side) == pytest.approx(ref, abs=1e-15)
    for side, ref in enumerate((1.0, math.sqrt(1.5), 1.25, 1.75, 1.25, 1.75)):
        assert area(HEX, HEX_EL, side) == pytest.approx(ref, abs=2e-15)


# ---------------------------------------------------------------------------------------------------------------------
# tests/LocalOperatorCommon.hpp:17-61 fixtures
QUAD_FIX = dict(et=QUAD, order=4, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0]], kernel="diffusion_kernel_2D", U=3)
HEX_FIX = dict(et=HEX, order=3,
               verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0], [1, 1, 1], [2, 1, 1.5], [1, 3, 2], [3, 4, 3.5]],
               kernel="diffusion_kernel_3D", U=4)


def _node_locations(orc, fix):
    et, p = fix["et"], fix["order"]
    gll = orc.lobatto(p + 1)
    n = p + 1
    locs = []
    for i in range(n**et):
        xi = [gll[i % n], gll[(i // n) % n]] + ([gll[i // (n * n)]] if et == HEX else [])
        locs.append(orc.map_to_physical(et, fix["verts"], xi))
    return np.array(locs)


def _apply_dirichlet(A, b, phi, bc_dofs, bnd_nodes):
    # tests/LocalOperatorCommon.hpp:190-205
    b = b - A[:, bc_dofs] @ phi[bnd_nodes, :]
    b[bc_dofs, :] = phi[bnd_nodes, :]
    A = A.copy()
    A[bc_dofs, :] = 0
    A[:, bc_dofs] = 0
    A[bc_dofs, bc_dofs] = 1
    return A, b


# tests/LocalAssemblyTests.cpp:3-43
@pytest.mark.parametrize("fix", [QUAD_FIX, HEX_FIX], ids=["diffusion2d_quad_p4", "diffusion3d_hex_p3"])
def test_local_system_assembly(orc, fix):
    et, p, U = fix["et"], fix["order"], fix["U"]
    phi = _node_locations(orc, fix)[:, :et]  # phi(x) = x_d, one rhs per dimension
    A, b = orc.assemble_local(fix["kernel"], et, p, fix["verts"], n_rhs=et, value_order=2)
    np.testing.assert_allclose(A, A.T, rtol=0, atol=0)
    bnd = orc.boundary_node_inds(et, p)
    bc_dofs = bnd * U
    A, b = _apply_dirichlet(A, b, phi, bc_dofs, bnd)
    x = np.linalg.solve(A, b)
    np.linalg.cholesky(A)  # the reference solves with LLT: the system must be SPD
    for node in range((p + 1) ** et):
        np.testing.assert_allclose(x[node * U], phi[node], rtol=1e-6)


# tests/LocalOperatorTests.cpp:3-95
@pytest.mark.parametrize("fix", [QUAD_FIX, HEX_FIX], ids=["diffusion2d_quad_p4", "diffusion3d_hex_p3"])
def test_local_operator_evaluation(orc, fix):
    et, p, U = fix["et"], fix["order"], fix["U"]
    n_rhs = et
    phi = _node_locations(orc, fix)[:, :et]
    A, b = orc.assemble_local(fix["kernel"], et, p, fix["verts"], n_rhs=n_rhs, value_order=2)
    bnd = orc.boundary_node_inds(et, p)
    bc_dofs = bnd * U
    A, b = _apply_dirichlet(A, b, phi, bc_dofs, bnd)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=b.shape)
    x_bc = x.copy()
    x_bc[bc_dofs, :] = 0
    diag, rhs = orc.precompute_diag_rhs(fix["kernel"], et, p, fix["verts"], n_rhs=n_rhs, value_order=2, dir_inds=bc_dofs,
                                        dir_vals=phi[bnd, :])
    y = orc.eval_local_operator(fix["kernel"], et, p, fix["verts"], x_bc, value_order=2)
    y[bc_dofs, :] = x[bc_dofs, :]
    diag[bc_dofs] = 1.0
    rhs[bc_dofs, :] = phi[bnd, :]
    eps = 1e-8
    assert np.linalg.norm(y - A @ x) < eps
    assert np.linalg.norm(np.diag(A) - diag) < eps
    assert np.linalg.norm(rhs - b) < eps


# tests/SumFactorizationTests.cpp:3-53 — variable-coefficient kernel, random external field, n_rhs = 2
@pytest.mark.parametrize("fix,kernel", [(QUAD_FIX, "diffusion_kernel_2D_var"), (HEX_FIX, "diffusion_kernel_3D_var")],
                         ids=["diffusion2d_var_quad_p4", "diffusion3d_var_hex_p3"])
@pytest.mark.parametrize("strategy", [0, 2, 3], ids=["auto", "standard", "odd_even"])
def test_sum_factorized_evaluation(orc, fix, kernel, strategy):
    et, p, U = fix["et"], fix["order"], fix["U"]
    nn = (p + 1) ** et
    rng = np.random.default_rng(5489)
    field = rng.uniform(-1, 1, size=(nn, 1))
    x = rng.uniform(-1, 1, size=(nn * U, 2))
    y_loc = orc.eval_local_operator(kernel, et, p, fix["verts"], x, node_vals=field, value_order=2)
    y_sf = orc.eval_sumfact(kernel, et, p, fix["verts"], x, node_vals=field, value_order=2, eval_strategy=strategy)
    # NOTE: the hex SF path passes z = 0 to the kernel (SumFactorization.hpp:732) — invisible for this kernel
    assert np.linalg.norm(y_loc - y_sf) < 1e-8


# tests/SumFactorizationTests.cpp:58-129
@pytest.mark.parametrize("EO,QO", [(3, 3), (3, 4), (4, 3), (4, 4)])
@pytest.mark.parametrize("kind", [0, 1, 2, 3], ids=["back_interp", "back_der", "fwd_interp_assign", "fwd_der_accumulate"])
def test_odd_even_decomposition(orc, EO, QO, kind):
    size = 33
    nb, nq = EO + 1, QO // 2 + 1
    rng = np.random.default_rng(EO * 10 + QO)
    x = rng.uniform(-1, 1, size=(nb if kind < 2 else nq, size))
    y0 = rng.uniform(-1, 1, size=(size, nb)) if kind == 3 else None
    y_sf = orc.sumfact_sweep(EO, QO, kind, False, x, y0)
    y_oe = orc.sumfact_sweep(EO, QO, kind, True, x, y0)
    assert np.linalg.norm(y_sf - y_oe) < 1e-8
    # and both equal the plain matrix product out = in^T * M
    interp, der = orc.sumfact_tables(EO, QO)
    M = [interp, der, interp.T, der.T][kind]
    ref = x.T @ M + (y0 if y0 is not None else 0)
    assert np.linalg.norm(y_sf - ref) < 1e-12


# benchmarks/Common.hpp:10-31 — the benchmark hex (vertex 7 at (2,2,2)); K_e x == matrix-free apply == SF apply
def test_benchmark_element_consistency(orc):
    verts = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [2, 2, 2]]
    rng = np.random.default_rng(3)
    for p in (2, 4):
        K, F = orc.assemble_local("bench_diffusion3d", HEX, p, verts)
        x = rng.uniform(-1, 1, size=(K.shape[0], 1))
        y1 = orc.eval_local_operator("bench_diffusion3d", HEX, p, verts, x)
        y2 = orc.eval_sumfact("bench_diffusion3d", HEX, p, verts, x)
        scale = np.linalg.norm(K @ x)
        assert np.linalg.norm(K @ x - y1) < 1e-12 * scale
        assert np.linalg.norm(K @ x - y2) < 1e-12 * scale
        diag, rhs = orc.precompute_diag_rhs("bench_diffusion3d", HEX, p, verts)
        assert np.linalg.norm(np.diag(K) - diag) < 1e-12 * np.linalg.norm(diag)
        assert np.linalg.norm(F - rhs) < 1e-12 * np.linalg.norm(rhs)


# AssembleLocalSystem.hpp:249 — degenerate elements throw
def test_degenerate_element_throws(orc):
    verts = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, -1], [1, 0, -1], [0, 1, -1], [1, 1, -1]]
    with pytest.raises(RuntimeError, match="degenerate element"):
        orc.assemble_local("diffusion_kernel_3D", HEX, 2, verts)


@pytest.mark.parametrize("dim", [2, 3])
def test_integrals_of_residual_kernels(orc, dim):
    """post/Integral.hpp / post/NormL2.hpp restatement against closed forms: measure, a polynomial moment, the perimeter / surface and
    the divergence theorem on a box [0, 2]^D; then the divergence theorem again on a distorted mesh (both sides polynomial)."""
    d = np.linspace(0.0, 2.0, 4)
    m = orc.mesh_square(d, d, order=2) if dim == 2 else orc.mesh_cube(d, d, d, order=2)
    f = np.zeros((2, m.n_nodes))
    vol = 2.0**dim
    got = m.compute_integral(f"integrand_probe_{dim}D", fields=f, time=0.5, value_order=2)
    moment = (8 / 3) * 2 if dim == 2 else 2 * (8 / 3) * 2  # int x^2 y over [0,2]^2; int x y^2 z over [0,2]^3
    assert np.allclose(got, [vol, moment + 0.5 * vol, 0.0], rtol=1e-13, atol=1e-13)
    sides = list(range(1, 2 * dim + 1))
    got = m.compute_integral(f"boundary_probe_{dim}D", boundary_ids=sides, fields=f)
    assert np.allclose(got, [2 * dim * 2.0 ** (dim - 1), dim * vol, 0.0], rtol=1e-13, atol=1e-13)
    # a nodal field equal to x: its integral, its x-derivative and the L2 norm of (1, ., .)
    got = m.compute_integral(f"integrand_probe_{dim}D", fields=f, norm_l2=True)
    assert abs(got[0] - np.sqrt(vol)) < 1e-13
    from common import distort

    m.set_verts(distort(m.elem_verts))
    v = m.compute_integral(f"integrand_probe_{dim}D", fields=f, value_order=3)[0]
    s = m.compute_integral(f"boundary_probe_{dim}D", boundary_ids=sides, fields=f, value_order=3)[1]
    assert abs(s / dim - v) < 1e-12 * v


def test_static_condensation_is_exact(orc):
    """The restated ElementBoundary condensation (StaticCondensationManager.hpp:330-535): solving the condensed system and recovering
    the interior values gives the solution of the full system (2-D diffusion, p=3, a penalty on T along the boundary makes K regular)."""
    from oracle import condense_element_boundary

    d = np.linspace(0.0, 1.0, 3)
    m = orc.mesh_square(d, d, order=3)
    U = 3
    s = m.assembled_system(U)
    s.assemble("example02_domain")
    vals, rhs = s.get()
    import l3ster_b200 as l3b
    import scipy.sparse as sp

    ptr, nbr = l3b.node_graph(m.n_nodes, m.elem_nodes.astype(np.uint32))
    row_ptr, col_ind = l3b.expand_graph(ptr, nbr, U)
    K = sp.csr_matrix((vals, col_ind, row_ptr), shape=(len(rhs), len(rhs))).toarray()
    nb = 4
    idx = np.arange(nb * nb)
    on_bnd = ((idx % nb) % 3 == 0) | ((idx // nb) % 3 == 0)
    bnd_idx, int_idx = np.flatnonzero(on_bnd), np.flatnonzero(~on_bnd)
    for n in np.unique(m.bnd_nodes):
        K[int(n) * U, int(n) * U] += 1e3
    prim, S, Fc, recover = condense_element_boundary(K, rhs, m.elem_nodes, bnd_idx, int_idx, U)
    assert len(prim) == m.n_nodes - m.n_elems * len(int_idx)
    x_full = np.linalg.solve(K, rhs)
    x_cond = recover(np.linalg.solve(S, Fc))
    assert np.abs(x_cond - x_full).max() < 1e-10 * np.abs(x_full).max()


# the element routine of the sum-factorised apply is instantiated with compile-time sizes for the timed configurations (oracle/sumfact.cpp:
# evalLocalOperatorSumFact) and falls back to run-time sizes elsewhere: every branch must reproduce K_e x of the dense assembly
@pytest.mark.parametrize("kernel,et,p,n_fields", [("ns3d_kernel", HEX, 2, 7), ("example02_domain", QUAD, 3, 0), ("dense_probe_3D", HEX, 2, 2),
                                                  ("bench_diffusion3d", HEX, 3, 0), ("diffusion_kernel_2D", QUAD, 4, 0)],
                         ids=["ns3d_8x7", "example02_4x3", "dense_probe_runtime_sizes", "diffusion3d_7x4", "diffusion2d"])
def test_sized_and_unsized_sum_factorised_routines_agree_with_the_dense_element_matrix(orc, kernel, et, p, n_fields):
    rng = np.random.default_rng(7)
    nn = (p + 1) ** et
    verts = np.array([[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)], dtype=float)[: 2 ** et]
    verts[-1] += 0.3  # not affine
    if et == QUAD:
        verts[:, 2] = 0.0
    field = rng.uniform(-1, 1, size=(nn, n_fields)) if n_fields else None
    K, _ = orc.assemble_local(kernel, et, p, verts, node_vals=field)
    x = rng.uniform(-1, 1, size=(K.shape[0], 1))
    # dense_probe_3D reads the point: the hex sum-factorised path hands it z = 0 (SumFactorization.hpp:732), the dense paths the true z —
    # that kernel depends on x and y only (tests/test_gpu_matrix_free.py pins the quirk)
    y_sf = orc.eval_sumfact(kernel, et, p, verts, x, node_vals=field)
    y_lo = orc.eval_local_operator(kernel, et, p, verts, x, node_vals=field)
    scale = np.linalg.norm(K @ x)
    assert np.linalg.norm(K @ x - y_lo) < 1e-12 * scale
    assert np.linalg.norm(K @ x - y_sf) < 1e-12 * scale


# tests/Diffusion2D.hpp:23-120 (the fixture of Diffusion2DAssembledTest / Diffusion2DMFTest / Diffusion2DMFSFTest): 4 x 4 quads p = 2 on the
# unit square, first-order diffusion system (T, qx, qy), T = x on the left / right sides as Dirichlet values, adiabatic top / bottom through
# the boundary kernel; the exact solution T = x, q = (1, 0) must come out to 1e-8 — here through the ORACLE's global assembly (scatter, dof
# numbering, algebraic Dirichlet rows) and its matrix-free system (lifting, Jacobi-CG), so that the checker's global half is pinned by the
# reference's own end-to-end answer without a GPU
def _diffusion2d_fixture(orc):
    node_dist = np.linspace(0.0, 1.0, 5)
    mesh = orc.mesh_square(node_dist, order=2)
    host_nodes = mesh.elem_nodes.astype(np.int64)
    gll = orc.lobatto(3)
    xs = np.zeros(mesh.n_nodes)
    for e in range(host_nodes.shape[0]):
        for a in range(9):
            xs[host_nodes[e, a]] = orc.map_to_physical(2, mesh.elem_verts[e], [gll[a % 3], gll[a // 3]])[0]
    bc_nodes = np.where((np.abs(xs) < 1e-14) | (np.abs(xs - 1.0) < 1e-14))[0]
    return mesh, xs, bc_nodes


def _check_diffusion2d_solution(sol, xs):
    assert np.abs(sol[0::3] - xs).max() < 1e-8
    assert np.abs(sol[1::3] - 1.0).max() < 1e-8
    assert np.abs(sol[2::3]).max() < 1e-8


def test_diffusion2d_end_to_end_assembled(orc):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    mesh, xs, bc_nodes = _diffusion2d_fixture(orc)
    s = mesh.assembled_system(3)
    s.assemble("diffusion_kernel_2D")
    s.assemble("adiabatic_bc_2D", boundary_ids=[1, 2])
    s.apply_dirichlet((bc_nodes * 3).astype(np.int32), xs[bc_nodes][:, None])
    vals, rhs = s.get()
    A = sp.csr_matrix((vals, s.col_ind, s.row_ptr), shape=(s.n_dofs,) * 2).tocsc()
    _check_diffusion2d_solution(spla.spsolve(A, rhs[:, 0]), xs)


@pytest.mark.parametrize("strategy", [1, 2], ids=["local_element", "sum_factorisation"])
def test_diffusion2d_end_to_end_matrix_free(orc, strategy):
    mesh, xs, bc_nodes = _diffusion2d_fixture(orc)
    mask = np.zeros(mesh.n_nodes * 3, dtype=np.uint8)
    mask[bc_nodes * 3] = 1
    vals = np.zeros((mesh.n_nodes * 3, 1))
    vals[bc_nodes * 3, 0] = xs[bc_nodes]
    s = mesh.matrix_free_system(3, 1, mask, vals)
    s.add_kernel("diffusion_kernel_2D", eval_strategy=strategy)
    s.add_kernel("adiabatic_bc_2D", boundary_ids=[1, 2])
    s.init()
    sol, res, iters = s.cg(tol=1e-12, max_iters=5000)
    assert res <= 1e-12 and iters < 5000
    _check_diffusion2d_solution(sol, xs)


# the same identity over every extent the compile-time sweep tables hold (1 .. 10 in both directions), odd and even sizes, all four kinds
@pytest.mark.parametrize("EO", [1, 2, 3, 4, 5, 6, 8, 9])
def test_odd_even_and_standard_sweeps_equal_the_matrix_product_for_all_tabulated_extents(orc, EO):
    rng = np.random.default_rng(100 + EO)
    nb = EO + 1
    for QO in (0, 1, 2, 3, 5, 6, 9, 12, 15, 18):
        nq = QO // 2 + 1
        interp, der = orc.sumfact_tables(EO, QO)
        assert interp.shape == (nb, nq)
        for kind in range(4):
            x = rng.uniform(-1, 1, size=(nb if kind < 2 else nq, 7))
            y0 = rng.uniform(-1, 1, size=(7, nb)) if kind == 3 else None
            ref = x.T @ [interp, der, interp.T, der.T][kind] + (y0 if y0 is not None else 0)
            for odd_even in (False, True):
                y = orc.sumfact_sweep(EO, QO, kind, odd_even, x, y0)
                assert np.linalg.norm(y - ref) < 1e-12 * max(1.0, np.linalg.norm(ref)), (EO, QO, kind, odd_even)
