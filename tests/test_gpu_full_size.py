"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle needs minutes there):
the operator is symmetric positive semi-definite (x.Ay == y.Ax, x.Ax >= 0), linear, Dirichlet rows are identity rows, the
assembled matrix applies like the matrix-free operator, and both kernels of each path agree with each other."""
import numpy as np
import pytest

import l3ster_b200 as l3b

pytestmark = pytest.mark.gpu
P, U = 4, 4
BND = [1, 2, 3, 4, 5, 6]


def node_dist(n):  # benchmarks/Diffusion3D.hpp:11-18
    dx, x, out = 1.0 / n, 0.0, []
    for _ in range(n + 1):
        out.append(x)
        x += dx
    return np.array(out)


@pytest.fixture(scope="module")
def ctx():
    return l3b.Context(0)


def test_matrix_free_apply_properties_at_64_cubed(ctx):
    """BASELINE configs[2]: 64^3 hex p=4, 67.9 M dofs."""
    import torch

    host = l3b.make_cube_mesh(node_dist(64), order=P)
    mesh = ctx.upload_mesh(host)
    mask = np.zeros(host.n_nodes * U, dtype=np.uint8)
    dir_dofs = host.boundary_nodes(BND) * U
    mask[dir_dofs] = 1
    s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    s.assembleProblem("bench_diffusion3d")
    s.endAssembly()
    n = s.n_dofs
    g = torch.Generator(device="cuda").manual_seed(5489)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    Ax, Ay, Axy = (torch.empty_like(x) for _ in range(3))
    z = 0.3 * x - 1.7 * y
    torch.cuda.synchronize()  # the tensors were filled on torch's stream, the library works on its own (non-blocking) stream
    s.apply_device(x.data_ptr(), Ax.data_ptr())
    s.apply_device(y.data_ptr(), Ay.data_ptr())
    s.apply_device(z.data_ptr(), Axy.data_ptr())
    ctx.synchronize()
    scale = float(torch.linalg.norm(Ax) * torch.linalg.norm(y))
    assert abs(float(torch.dot(x, Ay) - torch.dot(y, Ax))) <= 1e-12 * scale          # symmetry
    assert float(torch.dot(x, Ax)) > 0.0                                             # least-squares operator: SPD off the null space
    assert float(torch.linalg.norm(Axy - (0.3 * Ax - 1.7 * Ay)) / torch.linalg.norm(Axy)) < 1e-12  # linearity
    d = torch.from_numpy(dir_dofs).cuda()
    assert torch.equal(Ax[d], x[d])                                                  # Dirichlet rows are identity rows
    # y = alpha A x + beta y
    y0 = y.clone()
    torch.cuda.synchronize()
    s.apply_device(x.data_ptr(), y.data_ptr(), 1, -0.5, 2.0)
    ctx.synchronize()
    assert float(torch.linalg.norm(y - (2.0 * y0 - 0.5 * Ax)) / torch.linalg.norm(y)) < 1e-12


def test_assembled_matrix_applies_like_the_operator_at_12_cubed(ctx):
    """BASELINE configs[1] at 12^3 (1728 elements, 3 GB of CRS values): A_assembled x == A_matrix-free x, symmetric values."""
    host = l3b.make_cube_mesh(node_dist(12), order=P)
    mesh = ctx.upload_mesh(host)
    a = l3b.AssembledSystem(ctx, mesh, U, 1, host.node_graph())
    a.beginAssembly()
    a.assembleProblem("bench_diffusion3d")
    mf = l3b.MatrixFreeSystem(ctx, mesh, U, 1)
    mf.assembleProblem("bench_diffusion3d")
    mf.endAssembly()
    rng = np.random.default_rng(5489)
    x = rng.uniform(-1, 1, size=a.n_dofs)
    y_asm = a.spmv(x)
    y_mf = mf.apply(x.reshape(-1, 1)).ravel()
    assert np.linalg.norm(y_asm - y_mf) / np.linalg.norm(y_mf) < 1e-12
    z = rng.uniform(-1, 1, size=a.n_dofs)
    assert abs(z @ y_asm - x @ a.spmv(z)) <= 1e-12 * np.linalg.norm(z) * np.linalg.norm(y_asm)
    # the rhs the two paths build is the same vector
    _, rhs_asm = a.download(values=False)
    _, rhs_mf = mf.download()
    assert np.linalg.norm(rhs_asm - rhs_mf) / np.linalg.norm(rhs_mf) < 1e-12
