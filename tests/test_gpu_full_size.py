"""Parity at BASELINE.json's full sizes.

* against the ORACLE (threaded over the host cores; its mesh is built from the product's order-p arrays because the faithful
  restatement of convertMeshToOrder's geometric node matching needs minutes at 64^3 — the numbering itself is pinned bit-exactly at
  small sizes in test_gpu_assembly.py / test_host_logic.py): the 32^3 and 64^3 matrix-free applies on the axis-aligned benchmark mesh and
  on a distorted one, diag + rhs of the 64^3 init, and the 16^3 assembled CRS values + rhs entry for entry;
* through size-independent properties: symmetry, definiteness, linearity, Dirichlet identity rows, y = alpha A x + beta y, assembled
  matrix == matrix-free operator."""
import os

import numpy as np
import pytest

import l3ster_b200 as l3b
from common import distort, oracle, rel_err

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


def _host_threads():
    try:
        import psutil

        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


@pytest.mark.parametrize("n,distorted", [(32, False), (32, True), (64, False), (64, True)])
def test_matrix_free_apply_matches_the_oracle_at_benchmark_size(ctx, n, distorted):
    """BASELINE configs[2] (64^3 hex p=4, 67.9 M dofs; 32^3 as the mid-size point): y = A x with Dirichlet T = 0 on the six faces, and
    the diagonal + rhs of the init, against the oracle's sum-factorised apply (SumFactorization.hpp:882-917) at 1e-12"""
    host = l3b.make_cube_mesh(node_dist(n), order=P)
    verts = distort(host.verts, amp=0.1 / n) if distorted else np.array(host.verts)
    mesh = l3b.Mesh(ctx, 3, P, verts, host.nodes, host.side_boundaries, host.n_nodes, host.n_nodes)
    mask = np.zeros(host.n_nodes * U, dtype=np.uint8)
    mask[host.boundary_nodes(BND) * U] = 1
    s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
    s.assembleProblem("bench_diffusion3d")
    s.endAssembly()
    x = np.random.default_rng(5489).uniform(-1, 1, size=(s.n_dofs, 1))
    y = s.apply(x)
    om = oracle().mesh_from_nodes(3, P, host.n_nodes, host.nodes, verts)
    osys = om.matrix_free_system(U, 1, mask, None)
    osys.add_kernel("bench_diffusion3d")
    yo = osys.apply(x, n_threads=_host_threads())
    assert rel_err(y, yo) < 1e-12
    if n == 64:
        diag, rhs = s.download()
        odiag, orhs = osys.init(n_threads=_host_threads())
        assert rel_err(diag, odiag) < 1e-12 and rel_err(rhs, orhs) < 1e-12


def test_assembled_values_match_the_oracle_at_16_cubed(ctx):
    """BASELINE configs[1]: 16^3 hex p=4 (4096 elements, 0.91 G non-zeros): CRS values entry for entry and the rhs against the oracle's
    assembleGlobalSystem (AssembleLocalSystem.hpp:145-208 + ScatterLocalSystem.hpp:25-54) at 1e-12; falls back to 12^3 on a host with
    less than 48 GB of free memory (two 7.3 GB value arrays and the graphs live on the host at once)"""
    n = 16
    try:
        import psutil

        if psutil.virtual_memory().available < 48e9:
            n = 12
    except Exception:
        pass
    host = l3b.make_cube_mesh(node_dist(n), order=P)
    mesh = ctx.upload_mesh(host)
    a = l3b.AssembledSystem(ctx, mesh, U, 1, host.node_graph())
    a.beginAssembly()
    a.assembleProblem("bench_diffusion3d")
    vals, rhs = a.download()
    om = oracle().mesh_from_nodes(3, P, host.n_nodes, host.nodes, host.verts)
    oa = om.assembled_system(U)
    row_ptr, _ = l3b.expand_graph(a.node_ptr, a.node_nbr, U, with_cols=False)
    assert np.array_equal(row_ptr, oa.row_ptr)  # same graph (col_ind equality is pinned at small sizes; here the row structure)
    oa.assemble("bench_diffusion3d", n_threads=_host_threads())
    ovals, orhs = oa.get()
    assert len(vals) == len(ovals)
    err2, ref2 = 0.0, 0.0
    for beg in range(0, len(vals), 1 << 26):  # chunked: no third 7 GB temporary
        d = vals[beg:beg + (1 << 26)] - ovals[beg:beg + (1 << 26)]
        err2 += float(d @ d)
        ref2 += float(ovals[beg:beg + (1 << 26)] @ ovals[beg:beg + (1 << 26)])
    assert np.sqrt(err2 / ref2) < 1e-12
    assert rel_err(rhs, orhs) < 1e-12
