"""Partition import (l3b_partition_*): rank views of a globally known order-p mesh for externally supplied partition vectors.

The reference partitions with METIS (mesh/PartitionMesh.hpp:142-183) and then assigns / renumbers nodes deterministically
(:322-440); METIS is third party and absent here, so `epart` (and optionally `npart`) come from outside — a file, another tool, or
the geometric bisection below — and the library does the rest the way the reference does (see csrc/partition_host.hpp)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

import l3ster_b200 as l3b
from l3ster_b200 import _p, lib


@dataclass
class RankView:
    """What one rank holds of a partitioned mesh; same attribute names as slab.Slab where they mean the same thing."""
    rank: int
    world: int
    dim: int
    order: int
    extended: bool
    n_elems: int
    n_border_elems: int          # elements [0, n_border_elems) touch ghost nodes
    n_owned_nodes: int
    n_local_nodes: int
    first_gid: int               # owned global ids are first_gid + local id
    elem_ids: np.ndarray         # global element ids of the local elements
    nodes: np.ndarray            # (n_elems, nodes_per_elem) local ids
    verts: np.ndarray
    side_boundaries: np.ndarray
    gids: np.ndarray             # local id -> global id (renumbered)
    owned_halo: list             # [(rank, local owned nodes shared with it, in its ghost order)]
    shared_halo: list            # [(rank, offset, size)] ranges of the ghost block
    part: "Partition" = None

    @property
    def lower(self):  # Slab compatibility: "has neighbours"
        return self.shared_halo[0][0] if self.shared_halo else -1

    @property
    def upper(self):
        return self.owned_halo[0][0] if self.owned_halo else -1

    def local_ids(self, gids):
        """local ids of the global ids that are local here (others dropped)"""
        gids = np.asarray(gids, dtype=np.int64)
        no = self.n_owned_nodes
        own = gids[(gids >= self.first_gid) & (gids < self.first_gid + no)] - self.first_gid
        gh = self.gids[no:]
        pos = np.searchsorted(gh, gids)
        ok = (pos < len(gh)) & (gh[np.minimum(pos, max(len(gh) - 1, 0))] == gids) if len(gh) else np.zeros(len(gids), dtype=bool)
        return np.sort(np.concatenate([own, no + pos[ok]]))

    def dirichlet_nodes(self, boundary_ids):
        """local nodes (owned and ghost) on sides carrying one of `boundary_ids` — looked up in the WHOLE mesh: a ghost node can lie on
        the boundary through an element of another rank only (bcs/LocalDirichletBC.hpp gets that through its own exchange)"""
        return self.local_ids(self.part.global_boundary_nodes(boundary_ids))

    def halo_dofs(self, dofs_per_node):
        owned = [(r, (np.asarray(n, dtype=np.int64)[:, None] * dofs_per_node + np.arange(dofs_per_node)[None, :]).ravel().astype(np.int32))
                 for r, n in self.owned_halo]
        shared = [(r, off * dofs_per_node, size * dofs_per_node) for r, off, size in self.shared_halo]
        return owned, shared

    def device_halo(self, comm, dofs_per_node):
        owned, shared = self.halo_dofs(dofs_per_node)
        return l3b.DeviceHalo(comm, self.n_owned_nodes * dofs_per_node, (self.n_local_nodes - self.n_owned_nodes) * dofs_per_node, owned, shared)


class Partition:
    def __init__(self, dim, order, n_nodes, nodes, verts, side_boundaries, n_parts, epart, npart=None):
        self.dim, self.order, self.n_nodes, self.n_parts = dim, order, int(n_nodes), int(n_parts)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        self.verts = np.asarray(verts, dtype=np.float64)
        self.side_boundaries = np.asarray(side_boundaries, dtype=np.uint16)
        self.epart = np.ascontiguousarray(epart, dtype=np.int32)
        npart_c = None if npart is None else np.ascontiguousarray(npart, dtype=np.int32)
        self._h = C.c_void_p()
        rc = lib().l3b_partition_create(dim, order, self.n_nodes, self.nodes.shape[0], _p(self.nodes), n_parts, _p(self.epart), _p(npart_c),
                                        C.byref(self._h))
        if rc != 0:
            raise l3b.L3BError(rc, lib().l3b_global_error().decode())
        self.new_id = np.zeros(self.n_nodes, dtype=np.int64)
        self.npart = np.zeros(self.n_nodes, dtype=np.int32)
        self.dist = np.zeros(n_parts + 1, dtype=np.int64)
        lib().l3b_partition_node_map(self._h, _p(self.new_id), _p(self.npart), _p(self.dist))

    @classmethod
    def from_host_mesh(cls, host, n_parts, epart, npart=None):
        return cls(host.dim, host.order, host.n_nodes, host.nodes, host.verts, host.side_boundaries, n_parts, epart, npart)

    def __del__(self):
        try:
            lib().l3b_partition_destroy(self._h)
        except Exception:
            pass

    def rank_view(self, rank, extended=False) -> RankView:
        info = np.zeros(8, dtype=np.int64)
        rc = lib().l3b_partition_rank_info(self._h, rank, int(extended), _p(info))
        if rc != 0:
            raise l3b.L3BError(rc, lib().l3b_global_error().decode())
        n_el, n_border, n_owned, n_local, n_on, n_sn, n_send, first = map(int, info)
        nn = self.nodes.shape[1]
        elem_ids = np.zeros(n_el, dtype=np.int64)
        nodes = np.zeros((n_el, nn), dtype=np.uint32)
        gids = np.zeros(n_local, dtype=np.int64)
        lib().l3b_partition_rank_mesh(self._h, rank, int(extended), _p(elem_ids), _p(nodes), _p(gids))
        o_ranks, o_ptr, o_nodes = np.zeros(max(n_on, 1), dtype=np.int32), np.zeros(n_on + 1, dtype=np.int64), np.zeros(max(n_send, 1), dtype=np.int32)
        s_ranks, s_off = np.zeros(max(n_sn, 1), dtype=np.int32), np.zeros(n_sn + 1, dtype=np.int64)
        lib().l3b_partition_rank_halo(self._h, rank, int(extended), _p(o_ranks), _p(o_ptr), _p(o_nodes), _p(s_ranks), _p(s_off))
        owned = [(int(o_ranks[k]), o_nodes[o_ptr[k]:o_ptr[k + 1]].copy()) for k in range(n_on)]
        shared = [(int(s_ranks[j]), int(s_off[j]), int(s_off[j + 1] - s_off[j])) for j in range(n_sn)]
        return RankView(rank, self.n_parts, self.dim, self.order, bool(extended), n_el, n_border, n_owned, n_local, first, elem_ids, nodes,
                        self.verts[elem_ids], self.side_boundaries[elem_ids], gids, owned, shared, self)

    def rank_graph(self, rank, with_export_plan=True):
        """extended view: (node_ptr, node_nbr) for AssembledSystem(graph=...), and the receive plan (entry_ptr, pos) of export_shared_rows"""
        ptr, nbr, ep, pos = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = lib().l3b_partition_rank_graph(self._h, rank, C.byref(ptr), C.byref(nbr), C.byref(ep) if with_export_plan else None,
                                            C.byref(pos) if with_export_plan else None)
        if rc != 0:
            raise l3b.L3BError(rc, lib().l3b_global_error().decode())
        info = np.zeros(8, dtype=np.int64)
        lib().l3b_partition_rank_info(self._h, rank, 1, _p(info))
        n_local, n_send = int(info[3]), int(info[6])

        def take(ptr_, ctype, n):
            try:
                return np.ctypeslib.as_array(C.cast(ptr_, C.POINTER(ctype)), shape=(max(n, 1),)).copy()[:n]
            finally:
                lib().l3b_free(ptr_)

        p = take(ptr, C.c_int64, n_local + 1)
        n = take(nbr, C.c_uint32, int(p[-1]))
        if not with_export_plan:
            return (p, n), None
        e = take(ep, C.c_int64, n_send + 1)
        q = take(pos, C.c_uint32, int(e[-1]))
        return (p, n), (e, q)

    def global_boundary_nodes(self, boundary_ids):
        """global ids (renumbered) of all nodes on sides carrying one of `boundary_ids` — over the whole mesh, so that a rank also
        knows the Dirichlet status of ghost nodes whose boundary side belongs to another rank's element"""
        sel = np.zeros(self.n_nodes, dtype=bool)
        for side in range(2 * self.dim):
            on = np.isin(self.side_boundaries[:, side], list(boundary_ids))
            if on.any():
                sel[self.nodes[on][:, l3b.side_node_inds(self.dim, self.order, side)].ravel()] = True
        return np.sort(self.new_id[np.nonzero(sel)[0]])


def bisection_epart(centroids, n_parts):
    """Recursive coordinate bisection of the element centroids into n_parts parts of (almost) equal element counts — a stand-in for the
    external partitioner (METIS in the reference): parts are contiguous boxes, cut along the longest extent."""
    centroids = np.asarray(centroids, dtype=np.float64)
    epart = np.zeros(len(centroids), dtype=np.int32)

    def split(idx, first, count):
        if count == 1:
            epart[idx] = first
            return
        left = count // 2
        ext = centroids[idx].max(axis=0) - centroids[idx].min(axis=0)
        axis = int(np.argmax(ext))
        order = idx[np.argsort(centroids[idx, axis], kind="stable")]
        cut = len(order) * left // count
        split(order[:cut], first, left)
        split(order[cut:], first + left, count - left)

    split(np.arange(len(centroids)), 0, n_parts)
    return epart
