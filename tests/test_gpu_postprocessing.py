"""GPU parity: integrals and L2 norms of residual kernels (post/Integral.hpp, post/NormL2.hpp) against the CPU oracle and against
closed forms, through the C ABI; the reference's own end-to-end check of tests/Diffusion2D.hpp:80-117 on the device."""
import numpy as np
import pytest

import l3ster_b200 as l3b
from common import PairedMesh, default_dists, oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    return l3b.Context(0)


@pytest.mark.parametrize("dim,order,n", [(2, 1, 5), (2, 2, 4), (2, 4, 3), (3, 1, 3), (3, 2, 2), (3, 3, 2), (3, 4, 2)])
@pytest.mark.parametrize("opts", [l3b.AssemblyOptions(), l3b.AssemblyOptions(value_order=2, derivative_order=1)])
def test_integrals_match_oracle(ctx, dim, order, n, opts):
    pm = PairedMesh(dim, default_dists(dim, n), order)  # distorted: the Jacobian varies inside every element
    mesh = pm.upload(ctx)
    data = np.random.default_rng(7).uniform(-1, 1, size=(3, pm.n_nodes))
    fields = ctx.upload_fields(data)
    bnd = list(range(1, 2 * dim + 1))
    dom, bk = f"integrand_probe_{dim}D", f"boundary_probe_{dim}D"
    kw = dict(value_order=opts.value_order, der_order=opts.derivative_order)
    for norm in (False, True):
        f_gpu = mesh.computeNormL2 if norm else mesh.computeIntegral
        got = f_gpu(dom, fields=fields, field_inds=[2, 0], asm_opts=opts, time=0.25)
        ref = pm.orc.compute_integral(dom, fields=data, field_inds=[2, 0], time=0.25, norm_l2=norm, **kw)
        assert rel_err(got, ref) < TOL
        got = f_gpu(bk, boundary_ids=bnd[:-1], fields=fields, field_inds=[1, 2], asm_opts=opts)
        ref = pm.orc.compute_integral(bk, boundary_ids=bnd[:-1], fields=data, field_inds=[1, 2], norm_l2=norm, **kw)
        assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_closed_forms_on_a_distorted_mesh(ctx, dim):
    """measure of the domain two ways: the volume integral of 1 and, by the divergence theorem, the closed-surface integral of x . n / D
    (the bilinear / trilinear map makes both integrands polynomial: exact with p = 2 and doubled orders)"""
    pm = PairedMesh(dim, default_dists(dim, 3), 2)
    mesh = pm.upload(ctx)
    fields = ctx.upload_fields(np.zeros((2, pm.n_nodes)))
    opts = l3b.AssemblyOptions(value_order=3)
    vol = mesh.computeIntegral(f"integrand_probe_{dim}D", fields=fields, asm_opts=opts)[0]
    flux = mesh.computeIntegral(f"boundary_probe_{dim}D", boundary_ids=list(range(1, 2 * dim + 1)), fields=fields, asm_opts=opts)[1]
    assert abs(flux / dim - vol) < 1e-12 * vol


def test_diffusion2d_error_norms(ctx):
    """tests/Diffusion2D.hpp:23-117: solve, move the solution into the field storage, L2 norms of the error against the analytic
    solution over the domain and over the boundary, both below 1e-8"""
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
    s = l3b.AssembledSystem(ctx, mesh, U)
    s.beginAssembly()
    s.assembleProblem("diffusion_kernel_2D_r1")
    s.assembleProblem("adiabatic_bc_2D", boundary_ids=[1, 2])
    s.endAssembly((bc_nodes * U).astype(np.int32), xs[bc_nodes][:, None])
    sol, tol, _ = s.solve(tol=1e-10)
    assert tol <= 1e-10
    fields = ctx.upload_fields(np.ascontiguousarray(sol.reshape(-1, U).T))  # updateSolution: dof d of node n -> field d
    err_dom = mesh.computeNormL2("diffusion2d_error_dom", fields=fields, asm_opts=l3b.AssemblyOptions(value_order=1))
    err_bnd = mesh.computeNormL2("diffusion2d_error_bnd", boundary_ids=[2, 1, 3, 4], fields=fields)
    assert np.linalg.norm(err_dom) < 1e-8
    assert np.linalg.norm(err_bnd) < 1e-8
    # and it does measure something: a perturbed field is seen with the right size (|| 0.1 ||_L2 over the unit square = 0.1)
    pert = sol.reshape(-1, U).T.copy()
    pert[0] += 0.1
    fields.update(np.ascontiguousarray(pert))
    assert abs(mesh.computeNormL2("diffusion2d_error_dom", fields=fields)[0] - 0.1) < 1e-8


def test_kernel_kind_errors(ctx):
    host = l3b.make_square_mesh(np.linspace(0, 1, 3), order=2)
    mesh = ctx.upload_mesh(host)
    with pytest.raises(l3b.L3BError):
        mesh.computeIntegral("diffusion_kernel_2D_r1")  # an equation kernel is not an integrand
    a = l3b.AssembledSystem(ctx, mesh, 3)
    a.beginAssembly()
    with pytest.raises(l3b.L3BError):
        a.assembleProblem("diffusion2d_error_dom")  # and vice versa
    with pytest.raises(l3b.L3BError):
        mesh.computeIntegral("diffusion2d_error_dom")  # fields missing
    with pytest.raises(l3b.L3BError):
        mesh.computeIntegral("integrand_probe_3D")  # dimension mismatch


@pytest.mark.parametrize("dim,p,kernel,bnd", [(2, 2, "integrand_probe_2D", None), (2, 4, "integrand_probe_2D", None), (3, 2, "integrand_probe_3D", None),
                                              (2, 4, "boundary_probe_2D", [1, 4]), (3, 3, "boundary_probe_3D", [2, 5, 6])])
def test_values_at_nodes_match_oracle(dim, p, kernel, bnd):
    """computeValuesAtNodes (algsys/ComputeValuesAtNodes.hpp:316-506): residual kernels (fields, gradients, point, time, normal) evaluated
    at the nodes of the visited elements / sides on a distorted mesh, averaged over the elements sharing a node, written to chosen dofs of
    a 4-dof-per-node vector; untouched dofs keep their values"""
    import torch

    ctx = l3b.Context(0)
    pm = PairedMesh(dim, default_dists(dim, 3), p)
    mesh = pm.upload(ctx)
    dpn, dof_inds, time = 4, [2, 0, 3], 0.6
    fdata = np.random.default_rng(3).uniform(-1, 1, size=(2, pm.n_nodes))
    v0 = np.random.default_rng(4).uniform(-1, 1, size=pm.n_nodes * dpn)
    vd = torch.from_numpy(v0.copy()).cuda()
    torch.cuda.synchronize()
    mesh.computeValuesAtNodes(kernel, vd.data_ptr(), dpn, ids=bnd or (), dof_inds=dof_inds, fields=ctx.upload_fields(fdata), time=time)
    ctx.synchronize()
    vo = pm.orc.values_at_nodes(kernel, v0.reshape(-1, 1), dpn, boundary_ids=bnd or (), dof_inds=dof_inds, fields=fdata, time=time)[:, 0]
    got = vd.cpu().numpy()
    assert rel_err(got, vo) < 1e-12
    touched = got != v0
    if bnd is None:
        assert touched.reshape(-1, dpn)[:, dof_inds].all() and not touched.reshape(-1, dpn)[:, 1].any()
    else:
        on = np.zeros(pm.n_nodes, dtype=bool)
        on[pm.host.boundary_nodes(bnd)] = True
        assert not touched.reshape(-1, dpn)[~on].any() and touched.reshape(-1, dpn)[on][:, dof_inds].any()


def test_update_solution_on_the_device():
    """updateSolution (AssembledSystem.hpp:140-160): dof columns of the solution vector -> nodal fields, without leaving the GPU"""
    import torch

    ctx = l3b.Context(0)
    n_nodes, dpn = 1000, 4
    x = np.random.default_rng(0).uniform(-1, 1, size=n_nodes * dpn)
    f0 = np.random.default_rng(1).uniform(-1, 1, size=(6, n_nodes))
    fields = ctx.upload_fields(f0)
    xd = torch.from_numpy(x).cuda()
    torch.cuda.synchronize()
    fields.update_from_solution(xd.data_ptr(), dpn, [0, 1, 3], [2, 5, 0])
    ctx.synchronize()
    from l3ster_b200.slab import _DevicePtr

    view = torch.as_tensor(_DevicePtr(fields.device_ptr, f0.size), device="cuda").cpu().numpy().reshape(f0.shape)
    want = f0.copy()
    want[2], want[5], want[0] = x[0::dpn], x[1::dpn], x[3::dpn]
    assert np.array_equal(view, want)
