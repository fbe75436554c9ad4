"""Multi-rank matrix-free apply on the GPU.

* one process, one GPU: the ranks of a z-slab partition are emulated side by side — every slab runs the phased apply of the
  C ABI (INIT / ELEMENTS border + interior / FINISH), the halo moves with device copies in place of NCCL — and the stitched
  result must equal the apply on the unpartitioned mesh and the oracle's;
* two processes, two GPUs (skipped when the box has one): the real SlabOperator with NCCL point-to-point calls."""
import os
import subprocess
import sys

import numpy as np
import pytest

import l3ster_b200 as l3b
from l3ster_b200.slab import Halo, make_slab
from common import oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
U = 4
BND = [1, 2, 3, 4, 5, 6]


def _dists(n, nz):
    return np.linspace(0.0, 1.0, n + 1), np.linspace(0.0, 1.2, n + 1), np.linspace(0.0, 0.9, nz + 1)


def _global_key(lat, n_lat):
    return lat[:, 0] + n_lat[0] * (lat[:, 1] + n_lat[1] * lat[:, 2])


@pytest.mark.parametrize("order,world,nz", [(2, 3, 4), (4, 2, 3), (3, 4, 3)])
def test_emulated_ranks_match_the_single_mesh_apply(order, world, nz):
    import torch

    ctx = l3b.Context(0)
    x1, y1, z1 = _dists(2, nz)
    whole = make_slab(x1, y1, z1, order, 0, 1)
    n_lat = (len(x1) - 1) * order + 1, (len(y1) - 1) * order + 1
    key_to_global = {int(k): i for i, k in enumerate(_global_key(whole.lattice, n_lat))}
    rng = np.random.default_rng(5489)
    xg = rng.uniform(-1, 1, size=(whole.n_local_nodes, U))

    def system(slab):
        mesh = l3b.Mesh(ctx, 3, order, slab.verts, slab.nodes, slab.side_boundaries, slab.n_local_nodes, slab.n_owned_nodes)
        mask = np.zeros(slab.n_local_nodes * U, dtype=np.uint8)
        mask[slab.dirichlet_nodes(BND) * U] = 1
        s = l3b.MatrixFreeSystem(ctx, mesh, U, 1, mask, None)
        s.assembleProblem("bench_diffusion3d")
        s.endAssembly()
        return s, mask

    # reference: the unpartitioned mesh, product and oracle
    sys_w, mask_w = system(whole)
    y_ref = sys_w.apply(xg.reshape(-1, 1)).reshape(-1, U)
    om = oracle().mesh_cube(x1, y1, z1, order=order)
    assert np.array_equal(om.elem_nodes, whole.nodes.astype(np.uint64))
    osys = om.matrix_free_system(U, 1, mask_w, None)
    osys.add_kernel("bench_diffusion3d")
    assert rel_err(y_ref, osys.apply(xg.reshape(-1, 1)).reshape(-1, U)) < TOL

    slabs = [make_slab(x1, y1, z1, order, r, world) for r in range(world)]
    systems, halos, xs, ys, gids = [], [], [], [], []
    for s in slabs:
        gid = np.array([key_to_global[int(k)] for k in _global_key(s.lattice, n_lat)], dtype=np.int64)
        gids.append(gid)
        if s.n_elems == 0:
            systems.append(None), halos.append(None), xs.append(None), ys.append(None)
            continue
        systems.append(system(s)[0])
        halos.append(Halo(s, U, "cuda", ctx))
        xl = xg[gid].copy()
        xl[s.n_owned_nodes:] = np.nan  # ghosts must come from the Import
        xs.append(torch.from_numpy(xl.ravel()).cuda())
        ys.append(torch.full((s.n_local_nodes * U,), 7.0, dtype=torch.float64, device="cuda"))
    live = [r for r in range(world) if slabs[r].n_elems > 0]
    for r in live:  # pack
        halos[r].pack(xs[r])
    ctx.synchronize()
    for r in live:  # Import: device copy in place of NCCL
        if slabs[r].lower >= 0:
            xs[r][halos[r].n_owned_dofs:].copy_(halos[slabs[r].lower].send_up)
    torch.cuda.synchronize()
    for r in live:  # interior first, then border — as the overlapped schedule issues them
        s, nb = systems[r], slabs[r].n_border_elems
        s.apply_phase_device(xs[r].data_ptr(), ys[r].data_ptr(), l3b.APPLY_INIT, 0, 0)
        s.apply_phase_device(xs[r].data_ptr(), ys[r].data_ptr(), l3b.APPLY_ELEMENTS, nb, slabs[r].n_elems)
        s.apply_phase_device(xs[r].data_ptr(), ys[r].data_ptr(), l3b.APPLY_ELEMENTS, 0, nb)
    ctx.synchronize()
    for r in live:  # Export
        if slabs[r].upper >= 0:
            up = slabs[r].upper
            halos[r].recv_up.copy_(ys[up][halos[up].n_owned_dofs:])
    torch.cuda.synchronize()
    for r in live:
        halos[r].unpack_add(ys[r])
        systems[r].apply_phase_device(xs[r].data_ptr(), ys[r].data_ptr(), l3b.APPLY_FINISH, 0, 0)
    ctx.synchronize()
    y_stitched = np.full_like(y_ref, np.nan)
    for r in live:
        no = slabs[r].n_owned_nodes
        y_stitched[gids[r][:no]] = ys[r].cpu().numpy().reshape(-1, U)[:no]
    assert not np.isnan(y_stitched).any()
    assert rel_err(y_stitched, y_ref) < TOL


def test_two_process_nccl_apply_matches_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mp_slab_apply.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SLAB_APPLY_OK" in out.stdout


def test_two_process_karman_assembled_gmres():
    """BASELINE configs[3] partitioned: steady Navier-Stokes kernel + outlet kernel assembled on y-strips, distributed GMRES (NCCL)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mp_slab_karman.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29537", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SLAB_KARMAN_OK" in out.stdout


@pytest.mark.parametrize("nproc", [1, 2, 4])
def test_partition_import_worker(nproc):
    """tests/mp_partition.py: ragged external partition -> views, halo'd matrix-free apply + CG, row-complete assembled matrix with the
    shared-row export, all against the oracle's single-rank objects (one rank: the same code with an empty halo)"""
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs (gpurun --gpus {nproc})")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mp_partition.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                          "--master-port", str(29541 + nproc), script], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "PARTITION_OK" in out.stdout


@pytest.mark.parametrize("nproc", [1, 2])
def test_real_karman_mesh_partitioned_worker(nproc):
    """tests/mp_karman_mesh.py: BASELINE configs[3] on the real examples/07-karman-2D mesh, partitioned, row-complete assembled matrix,
    GMRES over NCCL"""
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs (gpurun --gpus {nproc})")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mp_karman_mesh.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                          "--master-port", str(29551 + nproc), script], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "KARMAN_MESH_OK" in out.stdout


def test_one_process_karman_worker():
    """the same worker on one rank (one GPU): the callback-driven GMRES against the library's own driver"""
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mp_slab_karman.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
                          "--master-port", "29539", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SLAB_KARMAN_OK" in out.stdout


def test_assembled_operators_on_an_empty_rank():
    """tests/EmptyPartitionTest.cpp:10-49: a rank without elements builds its (empty) assembled system under both condensation policies
    and can take part in the solves (the distributed solve with an empty rank runs in tests/mp_slab_apply.py)"""
    from l3ster_b200.slab import SlabAssembledOperator

    ctx = l3b.Context(0)
    x = np.linspace(0, 1, 3)
    slab = make_slab(x, x, x, 2, 2, 3)  # 2 element layers over 3 ranks
    assert slab.n_elems == 0
    for cond in (False, True):
        op = SlabAssembledOperator(ctx, slab, 4, "bench_diffusion3d", [1, 2, 3, 4, 5, 6], condensed=cond)
        assert op.n_local_dofs == 0 and op.n_owned_dofs == 0
        xs, res, its = op.solve(tol=1e-8)
        assert its == 0 and xs.numel() == 0
