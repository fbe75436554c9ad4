"""z-slab partition of the structured benchmark cube, one slab per rank.

The reference partitions with Metis and renumbers so that every rank owns a contiguous global node range, local ids
`[owned | ghost sorted by GID]` (mesh/PartitionMesh.hpp:411-440, dofs/NodeToDofMap.hpp:144-163). Metis is not available
here, so the partition is the structured one the benchmarks would get from any sane partitioner: rank r takes the element
layers z in [z0(r), z1(r)); the nodes of the plane shared with the slab below belong to the lower rank. Every rank builds
its slab alone (no communication at set-up): both sides of an interface enumerate the shared plane in the same (y, x)
lattice order, which is also the order of the ghost block — the stand-in for "sorted by GID".

Halo pattern of the matrix-free apply (algsys/MatrixFreeSystem.hpp:1046-1122, comm/ImportExport.hpp):
  Import  x: owner (rank r, top plane, packed) -> ghost block of rank r+1 (contiguous tail of its local vector)
  Export  y: ghost block of rank r+1 (contiguous) -> rank r, added into its top-plane dofs
Elements are numbered layer by layer, so the border elements (the ones that touch ghost nodes) are the first
`n_border_elems` of the slab and the interior ones the contiguous rest.
"""
import os
from dataclasses import dataclass

import numpy as np

import l3ster_b200 as l3b


def split_layers(n_layers, world):
    """Element layers [z0, z1) of every rank; ranks may get zero layers when world > n_layers (tests/EmptyPartitionTest.cpp)."""
    base, rem = divmod(n_layers, world)
    out, z = [], 0
    for r in range(world):
        k = base + (1 if r < rem else 0)
        out.append((z, z + k))
        z += k
    return out


@dataclass
class Slab:
    rank: int
    world: int
    order: int
    dim: int
    n_elems: int
    n_local_nodes: int
    n_owned_nodes: int
    nodes: np.ndarray            # (n_elems, nodes_per_elem) local ids in the [owned | ghost] numbering
    verts: np.ndarray            # (n_elems, 8, 3)
    side_boundaries: np.ndarray  # (n_elems, 6), 0xFFFF = not on a physical boundary
    lattice: np.ndarray          # (n_local_nodes, 3) global order-p lattice coordinates of the local nodes
    n_border_elems: int          # elements [0, n_border_elems) touch ghost nodes
    send_up_nodes: np.ndarray    # owned nodes of the top plane, (y, x) order: Import source / Export target
    lower: int                   # rank that owns our ghosts, or -1
    upper: int                   # rank whose ghosts we own, or -1

    @property
    def n_ghost_nodes(self):
        return self.n_local_nodes - self.n_owned_nodes

    def dirichlet_nodes(self, boundary_ids):
        """local nodes (owned and ghost) on the physical sides carrying one of `boundary_ids`"""
        sel = np.zeros(self.n_local_nodes, dtype=bool)
        for side in range(2 * self.dim):
            on = np.isin(self.side_boundaries[:, side], list(boundary_ids))
            if on.any():
                sel[self.nodes[on][:, l3b.side_node_inds(self.dim, self.order, side)].ravel()] = True
        return np.nonzero(sel)[0]


def make_slab(x, y, z, order, rank, world) -> Slab:
    """Slab `rank` of `world` of the cube mesh over the vertex coordinates x, y, z (benchmarks/Diffusion3D.hpp:8-24), cut into layers
    along z; with z = None: strip `rank` of the square mesh over x, y, cut along y (the 2-D configurations)."""
    dim = 2 if z is None else 3
    axes = [np.ascontiguousarray(a, dtype=np.float64) for a in ((x, y) if dim == 2 else (x, y, z))]
    cut = axes[-1]
    layers = split_layers(len(cut) - 1, world)
    z0, z1 = layers[rank]
    p = order
    nn, nv, ns = (p + 1) ** dim, 2**dim, 2 * dim
    if z1 == z0:
        e = np.zeros
        return Slab(rank, world, p, dim, 0, 0, 0, e((0, nn), np.uint32), e((0, nv, 3)), e((0, ns), np.uint16), e((0, dim), np.int64), 0,
                    e(0, np.int64), -1, -1)
    # neighbours: nearest non-empty slabs
    lower = next((r for r in range(rank - 1, -1, -1) if layers[r][1] > layers[r][0]), -1)
    upper = next((r for r in range(rank + 1, world) if layers[r][1] > layers[r][0]), -1)
    local_axes = axes[:-1] + [cut[z0:z1 + 1]]
    host = l3b.make_square_mesh(*local_axes, order=p) if dim == 2 else l3b.make_cube_mesh(*local_axes, order=p)
    nodes0, verts = np.array(host.nodes, dtype=np.int64), np.array(host.verts)
    sb = np.array(host.side_boundaries)
    n_elems, n_nodes = host.n_elems, host.n_nodes
    # lattice coordinates of every local node from the element's position and the node's place in it
    a = np.arange(nn)
    lat = np.zeros((n_nodes, dim), dtype=np.int64)
    el_pos = []
    for d in range(dim):
        ed = np.searchsorted(axes[d], verts[:, 0, d])
        el_pos.append(ed)
        lat[nodes0.ravel(), d] = (ed[:, None] * p + ((a // (p + 1) ** d) % (p + 1))[None, :]).ravel()
    ec = el_pos[-1]
    # ownership: the plane (line) shared with the slab below belongs to the lower rank
    ghost = (lat[:, -1] == z0 * p) if lower >= 0 else np.zeros(n_nodes, dtype=bool)
    stride = len(axes[0]) * p + 1
    plane_key = lat[:, 0] if dim == 2 else lat[:, 1] * stride + lat[:, 0]
    owned_ids = np.nonzero(~ghost)[0]
    ghost_ids = np.nonzero(ghost)[0]
    ghost_ids = ghost_ids[np.argsort(plane_key[ghost_ids], kind="stable")]
    perm = np.empty(n_nodes, dtype=np.int64)
    perm[owned_ids] = np.arange(len(owned_ids))
    perm[ghost_ids] = len(owned_ids) + np.arange(len(ghost_ids))
    nodes = perm[nodes0].astype(np.uint32)
    lat_new = np.empty_like(lat)
    lat_new[perm] = lat
    top = np.nonzero(lat_new[:, -1] == z1 * p)[0] if upper >= 0 else np.zeros(0, dtype=np.int64)
    top_key = lat_new[top, 0] if dim == 2 else lat_new[top, 1] * stride + lat_new[top, 0]
    top = top[np.argsort(top_key, kind="stable")]
    # the faces between slabs are not physical boundaries (sides 0 = low, 1 = high of the cut direction in both element types,
    # mesh/ElementTraits.hpp:88-93, 118-137)
    no_bnd = np.uint16(0xFFFF)
    if lower >= 0:
        sb[ec == z0, 0] = no_bnd
    if upper >= 0:
        sb[ec == z1 - 1, 1] = no_bnd
    touches_ghost = (nodes >= len(owned_ids)).any(axis=1)
    n_border = int(touches_ghost.sum())
    assert not touches_ghost[n_border:].any(), "border elements are expected to be the first layer of the slab"
    return Slab(rank, world, p, dim, n_elems, n_nodes, len(owned_ids), nodes, verts, sb, lat_new, n_border, top, lower, upper)


class _DevicePtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _device_view(ptr, n, device):
    """zero-copy torch view of n doubles of device memory owned by the library"""
    import torch

    if not ptr or n == 0:  # an empty rank: the library holds no storage
        return torch.empty(0, dtype=torch.float64, device=device)
    return torch.as_tensor(_DevicePtr(ptr, n), device=device)


class Halo:
    """Import / Export of the slab's interface dofs (comm/ImportExport.hpp:29-472) over torch.distributed point-to-point
    calls: NCCL for device tensors, gloo for host tensors (the CPU tests). Vectors cover the local dofs [owned | ghost]."""

    def __init__(self, slab: Slab, dofs_per_node, device="cpu", ctx=None):
        import torch

        self.torch, self.slab, self.ctx = torch, slab, ctx
        self.n_owned_dofs = slab.n_owned_nodes * dofs_per_node
        self.n_local_dofs = slab.n_local_nodes * dofs_per_node
        self.n_ghost_dofs = self.n_local_dofs - self.n_owned_dofs
        up = (slab.send_up_nodes[:, None] * dofs_per_node + np.arange(dofs_per_node)[None, :]).ravel().astype(np.int32)
        self.n_up = len(up)
        self.up_idx = torch.from_numpy(up).to(device)
        self.send_up = torch.empty(self.n_up, dtype=torch.float64, device=device)
        self.recv_up = torch.empty(self.n_up, dtype=torch.float64, device=device)

    # pack / unpack: the library's kernels on the device (asynchronous on the context stream), torch indexing on the host
    def pack(self, x):
        if self.n_up == 0:
            return 0
        if x.is_cuda:
            self.ctx.vec_gather(x.data_ptr(), self.n_local_dofs, self.up_idx.data_ptr(), self.n_up, self.send_up.data_ptr())
            return 1
        self.send_up.copy_(x[self.up_idx.long()])
        return 0

    def unpack_add(self, y):
        if self.n_up == 0:
            return 0
        if y.is_cuda:
            self.ctx.vec_scatter_add(y.data_ptr(), self.n_local_dofs, self.up_idx.data_ptr(), self.n_up, self.recv_up.data_ptr())
            return 1
        y.index_add_(0, self.up_idx.long(), self.recv_up)
        return 0

    def _p2p(self, send, dst, recv, src):
        import torch.distributed as dist

        ops = []
        if dst >= 0 and send is not None and send.numel() > 0:
            ops.append(dist.P2POp(dist.isend, send, dst))
        if src >= 0 and recv is not None and recv.numel() > 0:
            ops.append(dist.P2POp(dist.irecv, recv, src))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def import_x(self, x):
        """owner -> ghost copies: the packed top plane goes up, the ghost block (contiguous tail of x) is filled from below"""
        self._p2p(self.send_up, self.slab.upper, x[self.n_owned_dofs:] if self.n_ghost_dofs else None, self.slab.lower)

    def export_y(self, y):
        """ghost contributions -> owner: the ghost block goes down, the contributions from above land in recv_up"""
        self._p2p(y[self.n_owned_dofs:] if self.n_ghost_dofs else None, self.slab.lower, self.recv_up, self.slab.upper)


def comm_of(ctx):
    """the context's communicator over the torch.distributed ranks (created once; NCCL in the library, torch only ships the id)"""
    if getattr(ctx, "_slab_comm", None) is None:
        ctx._slab_comm = l3b.Comm.from_torch_distributed(ctx)
    return ctx._slab_comm


def device_halo(ctx, slab: Slab, dofs_per_node, n_owned_nodes=None, n_local_nodes=None, send_up_nodes=None):
    """l3b_halo of the slab's dof layout: the top plane is shared with the upper rank (packed in (y, x) order = the order of the upper
    rank's ghost block), the ghost block is owned by the lower rank"""
    if hasattr(slab, "owned_halo") and n_owned_nodes is None:  # a partition.RankView: general neighbour lists
        return slab.device_halo(comm_of(ctx), dofs_per_node)
    n_owned_nodes = slab.n_owned_nodes if n_owned_nodes is None else n_owned_nodes
    n_local_nodes = slab.n_local_nodes if n_local_nodes is None else n_local_nodes
    send_up_nodes = slab.send_up_nodes if send_up_nodes is None else send_up_nodes
    up = (np.asarray(send_up_nodes)[:, None] * dofs_per_node + np.arange(dofs_per_node)[None, :]).ravel().astype(np.int32)
    n_ghost = (n_local_nodes - n_owned_nodes) * dofs_per_node
    owned = [(slab.upper, up)] if slab.upper >= 0 and len(up) else []
    shared = [(slab.lower, 0, n_ghost)] if slab.lower >= 0 and n_ghost else []
    return l3b.DeviceHalo(comm_of(ctx), n_owned_nodes * dofs_per_node, n_ghost, owned, shared)


class SlabOperator:
    """Matrix-free operator of one rank (a Slab, or a partition.RankView of an imported partition) with the overlap of
    MatrixFreeSystem::applyImpl (:1046-1122) inside the library: Import behind the zeroing of y, border elements, Export behind the
    interior elements. Vectors are device tensors over the local dofs [owned | ghost]."""

    def __init__(self, ctx, slab: Slab, dofs_per_node, kernel, dirichlet_boundary_ids=()):
        import torch

        self.torch, self.ctx, self.slab, self.dpn = torch, ctx, slab, dofs_per_node
        dev = torch.device("cuda", torch.cuda.current_device())
        self.n_local_dofs, self.n_owned_dofs = slab.n_local_nodes * dofs_per_node, slab.n_owned_nodes * dofs_per_node
        self.mesh = self.sys = None
        if slab.n_elems > 0:
            self.mesh = l3b.Mesh(ctx, slab.dim, slab.order, slab.verts, slab.nodes, slab.side_boundaries, slab.n_local_nodes, slab.n_owned_nodes)
            mask = np.zeros(self.n_local_dofs, dtype=np.uint8)
            if dirichlet_boundary_ids:
                mask[slab.dirichlet_nodes(dirichlet_boundary_ids) * dofs_per_node] = 1  # dof 0 (T), benchmarks/Diffusion3D.hpp:102-104
            self.sys = l3b.MatrixFreeSystem(ctx, self.mesh, dofs_per_node, 1, mask, None)
            self.sys.assembleProblem(kernel)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        self.launches = 0
        self.comm = comm_of(ctx) if slab.world > 1 else None  # collective: every rank, also an empty one
        self.diag = self.rhs = None
        if self.sys is not None:
            if slab.world > 1:
                self.dev_halo = device_halo(ctx, slab, dofs_per_node)
                self.sys.set_halo(self.dev_halo, slab.n_border_elems)
            # computeDiagAndRhs (MatrixFreeSystem.hpp:877-941): element contributions, Export-sum of the ghost parts, Dirichlet dofs —
            # one library call (l3b_mf_end_assembly)
            self.sys.endAssembly()
            self.diag = _device_view(self.sys.device_diag, self.n_local_dofs, dev)
            self.rhs = _device_view(self.sys.device_rhs, self.n_local_dofs, dev)

    def solve(self, tol=1e-6, max_iters=10000, x0=None):
        """CG + native Jacobi over all ranks (benchmarks/Diffusion3D.hpp:115-118): l3b_mf_solve_device — the library's PCG driver with the
        halo'd apply and ncclAllReduce for the dot products. Returns (x over the local dofs, achieved residual norm, iterations)."""
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        x = torch.zeros(max(self.n_local_dofs, 1), dtype=torch.float64, device=dev) if x0 is None else x0
        torch.cuda.current_stream().synchronize()  # torch filled x on its own stream; the library's stream does not wait for it
        if self.sys is None:  # empty rank (tests/EmptyPartitionTest.cpp): zero dofs, it takes part in the reductions only
            def allreduce(sp, n):
                self.comm.allreduce_sum(sp, n)

            res, it = self.ctx.pcg(0, 0, lambda xp, yp, ep: False, allreduce if self.comm is not None else None, None, None, x.data_ptr(), tol, max_iters)
            return x[:0], res, it
        res, it = self.sys.solve_device(x.data_ptr(), "cg", tol, max_iters, x0_is_zero=x0 is None)
        return x, res, it

    def apply(self, x, y, alpha=1.0, beta=0.0, energy_ptr=None):
        """y[owned] = alpha (A x)[owned] + beta y[owned]; x[ghost] is overwritten by the Import. Asynchronous on the context stream: ONE
        library call (l3b_mf_apply_device with the system's halo: pack + Import behind the zeroing of y, border elements, Export behind
        the interior elements, unpack-add, Dirichlet rows). energy_ptr: device scalar that receives this rank's share of x^T A x."""
        if self.sys is None:
            return
        self.sys.apply_device(x.data_ptr(), y.data_ptr(), 1, alpha, beta, energy_ptr=energy_ptr)
        self.launches = self.sys.kernel_launches


class _SlabAsHost:
    """the slab's arrays under the names CondensedAssembledSystem reads from a HostMesh"""

    def __init__(self, slab: Slab):
        self.dim, self.order, self.nodes, self.verts, self.side_boundaries = slab.dim, slab.order, slab.nodes, slab.verts, slab.side_boundaries
        self.n_nodes, self.n_elems = slab.n_local_nodes, slab.n_elems


class SlabAssembledOperator:
    """Assembled system of one slab (BASELINE configs[1] on more than one GPU): every rank assembles its elements into the rows of
    its local nodes [owned | ghost] (`assembleProblem`, no exchange — as in the reference); the global operator is Import x, local
    sparse product, Export-sum of the ghost rows, the global diagonal and rhs the Export-sums of the local ones. The reference
    export-adds the shared ROWS to their owners at endAssembly instead (AssembledSystem.hpp:384-389): same operator, same halo."""

    def __init__(self, ctx, slab: Slab, dofs_per_node, kernel, dirichlet_boundary_ids=(), dirichlet_value=0.0, *, field_data=None,
                 dirichlet=None, condensed=False):
        """kernel: a name, or a list of dicts (name, boundary_ids, asm_opts, dof_inds, field_inds, time) assembled in turn (domain and
        boundary kernels of one problem); field_data: (n_fields, n_local_nodes) nodal values in the slab's numbering, ghosts included
        (post/FieldAccess.hpp: the values at ghost nodes must be current); dirichlet: (local dofs, values) instead of the
        (boundary ids -> dof 0 = value) short form; condensed: CondensationPolicy::ElementBoundary — the operator then lives on the
        primary (element-boundary) nodes of the slab, [owned | ghost] like the nodes (every ghost node is a primary one), and
        `recover` returns the nodal solution"""
        import torch

        self.torch, self.ctx, self.slab, self.dpn = torch, ctx, slab, dofs_per_node
        dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        kernels = [dict(name=kernel)] if isinstance(kernel, str) else kernel
        if dirichlet is not None:
            dofs, vals = np.asarray(dirichlet[0], dtype=np.int64), np.asarray(dirichlet[1], dtype=np.float64).reshape(-1, 1)
        else:
            dofs = slab.dirichlet_nodes(dirichlet_boundary_ids) * dofs_per_node if dirichlet_boundary_ids else np.zeros(0, dtype=np.int64)
            vals = np.full((len(dofs), 1), dirichlet_value)
        self.cs = None
        if condensed:
            import dataclasses

            from .condensation import CondensedAssembledSystem

            self.cs = CondensedAssembledSystem(ctx, _SlabAsHost(slab), dofs_per_node)
            n_owned_prim = int((self.cs.primary_nodes < slab.n_owned_nodes).sum())  # local ids are [owned | ghost]: owned primaries first
            assert (self.cs.prim_of[slab.n_owned_nodes:] >= 0).all(), "ghost nodes lie on element boundaries"
            send_up = self.cs.prim_of[slab.send_up_nodes] if len(slab.send_up_nodes) else slab.send_up_nodes
            pslab = dataclasses.replace(slab, n_local_nodes=len(self.cs.primary_nodes), n_owned_nodes=n_owned_prim, send_up_nodes=send_up)
            self.n_local_dofs, self.n_owned_dofs = pslab.n_local_nodes * dofs_per_node, pslab.n_owned_nodes * dofs_per_node
            self.dev_halo = device_halo(ctx, pslab, dofs_per_node) if slab.world > 1 else None
            self.cs.beginAssembly()
            for k in kernels:
                self.cs.assembleProblem(k["name"], k.get("boundary_ids", ()), field_data if l3b.kernel_info(k["name"])["n_fields"] else None,
                                        k.get("field_inds"), k.get("dof_inds"), k.get("asm_opts", l3b.AssemblyOptions()), k.get("time", 0.0))
            self.cs.endAssembly(dofs, vals, n_owned_primary_dofs=self.n_owned_dofs)
            self.mesh, self.sys = self.cs.local_mesh, self.cs.condensed
        else:
            self.n_local_dofs, self.n_owned_dofs = slab.n_local_nodes * dofs_per_node, slab.n_owned_nodes * dofs_per_node
            self.dev_halo = device_halo(ctx, slab, dofs_per_node) if slab.world > 1 else None
            self.mesh = l3b.Mesh(ctx, slab.dim, slab.order, slab.verts, slab.nodes, slab.side_boundaries, slab.n_local_nodes, slab.n_owned_nodes)
            # a partition.RankView in the extended numbering: the reference's row-complete owner matrix — the graph of the owned rows
            # holds every rank's contributions (SparsityGraph.hpp:83-278) and the shared rows are export-added at endAssembly
            # (AssembledSystem.hpp:384-389); otherwise every rank keeps the ghost rows its elements filled
            self.row_export = bool(getattr(slab, "extended", False))
            graph = plan = None
            if self.row_export:
                graph, plan = slab.part.rank_graph(slab.rank)
            self.sys = l3b.AssembledSystem(ctx, self.mesh, dofs_per_node, 1, graph)
            if self.dev_halo is not None:
                self.sys.set_halo(self.dev_halo)  # spmv_device, diag_device and the solvers now act over all ranks
            self.fields = ctx.upload_fields(field_data) if field_data is not None else None
            self.sys.beginAssembly()
            for k in kernels:
                self.sys.assembleProblem(k["name"], k.get("boundary_ids", ()), self.fields if l3b.kernel_info(k["name"])["n_fields"] else None,
                                         k.get("field_inds"), k.get("dof_inds"), k.get("asm_opts", l3b.AssemblyOptions()), k.get("time", 0.0))
            if self.row_export and self.dev_halo is not None:
                self.sys.export_shared_rows(*plan)
            self.sys.endAssemblyRanked(dofs.astype(np.int32), vals, self.n_owned_dofs)
        if self.dev_halo is not None and self.sys._halo is None:
            self.sys.set_halo(self.dev_halo)
        self.rhs = _device_view(self.sys.device_rhs, self.n_local_dofs, dev).clone()
        self.diag = torch.zeros(self.n_local_dofs, dtype=torch.float64, device=dev)
        torch.cuda.current_stream().synchronize()
        self.sys.diag_device(self.diag.data_ptr())  # global diagonal: Export-sum of the local ones (or the complete owned rows')
        if self.dev_halo is not None and not getattr(self, "row_export", False):
            self.dev_halo.export_add(self.rhs.data_ptr())  # global rhs on the owned rows (a copy; the system keeps its own)
        ctx.synchronize()

    def recover(self, x):
        """condensed operator only: nodal solution over the slab's local nodes (n_local_nodes * dofs_per_node, host) from the condensed
        solution x over the primary dofs; the ghost primaries are refreshed from their owners first"""
        if self.dev_halo is not None:
            self.dev_halo.import_(x.data_ptr())
        self.ctx.synchronize()
        return self.cs.recover(x.cpu().numpy())[:, 0]

    def apply(self, x, y):
        """y[owned] = (A x)[owned]; x[ghost] is overwritten by the Import. Asynchronous on the context stream (l3b_asm_spmv_device with
        the system's halo: Import x, local sparse product, Export-sum of the ghost rows)."""
        self.sys.spmv_device(x.data_ptr(), y.data_ptr())

    def solve(self, tol=1e-6, max_iters=10000, gmres=False, restart_length=250, max_restarts=39, callbacks=False):
        """CG (or, gmres=True, restarted GMRES) + Jacobi over all ranks on the assembled matrix (solve/BelosSolvers.hpp:116-131):
        l3b_asm_solve_device; callbacks=True drives the same iteration through l3b_pcg_device / l3b_gmres_device with this object's
        apply and the communicator's all-reduce (the entry points a user with an operator of their own binds to)."""
        torch = self.torch
        dev = torch.device("cuda", torch.cuda.current_device())
        x = torch.zeros(max(self.n_local_dofs, 1), dtype=torch.float64, device=dev)
        torch.cuda.current_stream().synchronize()  # torch filled x on its own stream; the library's stream does not wait for it
        if not callbacks:
            res, it = self.sys.solve_device(x.data_ptr(), "gmres" if gmres else "cg", tol, max_iters, restart_length, max_restarts)
            return x[:self.n_local_dofs], res, it
        comm = comm_of(self.ctx) if self.slab.world > 1 else None

        def apply(xp, yp, _energy_ptr):
            self.sys.spmv_device(xp, yp)
            return False

        def allreduce(sp, n):
            comm.allreduce_sum(sp, n)

        if gmres:
            res, it = self.ctx.gmres(self.n_local_dofs, self.n_owned_dofs, apply, allreduce if comm else None, self.diag.data_ptr(),
                                     self.rhs.data_ptr(), x.data_ptr(), tol, restart_length, max_restarts, max_iters)
        else:
            res, it = self.ctx.pcg(self.n_local_dofs, self.n_owned_dofs, apply, allreduce if comm else None, self.diag.data_ptr(),
                                   self.rhs.data_ptr(), x.data_ptr(), tol, max_iters)
        return x[:self.n_local_dofs], res, it
