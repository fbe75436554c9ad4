"""Static condensation of the element-interior dofs — the reference's `CondensationPolicy::ElementBoundary`
(algsys/StaticCondensationManager.hpp:135-535, what benchmarks/Diffusion3DBenchmark.cpp:6 runs) — on top of the C ABI.

`CondensedAssembledSystem` has the reference's AssembledSystem life cycle (beginAssembly / assembleProblem / endAssembly / solve); the
global matrix only holds the primary dofs (nodes on element boundaries), the interior ones are eliminated element by element on the
device (`csrc/condense.cuh`) and recovered after the solve. No arithmetic happens here: this file builds index lists."""
import ctypes as C

import numpy as np

from . import AssembledSystem, AssemblyOptions, Context, HostMesh, L3BError, Mesh, _p, lib, node_graph


def boundary_interior_split(dim, order):
    """local node indices (lexicographic, x fastest) on the element boundary / in its interior (mesh/ElementTraits.hpp: boundary_node_inds,
    internal_node_inds — both ascending)"""
    nb = order + 1
    idx = np.arange(nb**dim)
    on_bnd = np.zeros(nb**dim, dtype=bool)
    for d in range(dim):
        c = (idx // nb**d) % nb
        on_bnd |= (c == 0) | (c == order)
    return np.flatnonzero(on_bnd).astype(np.int32), np.flatnonzero(~on_bnd).astype(np.int32)


class CondensedAssembledSystem:
    def __init__(self, ctx: Context, host: HostMesh, dofs_per_node, n_rhs=1):
        self.ctx, self.host, self.dofs_per_node, self.n_rhs = ctx, host, dofs_per_node, n_rhs
        nn = host.nodes.shape[1]
        self.bnd_idx, self.int_idx = boundary_interior_split(host.dim, host.order)
        # primary nodes: on some element's boundary, numbered in node-id order (the dof numbering of dofs/NodeToDofMap.hpp:249-264 runs
        # over the nodes that carry dofs, in node-id order)
        self.primary_nodes = np.unique(host.nodes[:, self.bnd_idx])
        self.prim_of = np.full(host.n_nodes, -1, dtype=np.int64)
        self.prim_of[self.primary_nodes] = np.arange(len(self.primary_nodes))
        self.elem_prim = np.ascontiguousarray(self.prim_of[host.nodes[:, self.bnd_idx]], dtype=np.uint32)
        n_prim = len(self.primary_nodes)
        self.condensed = AssembledSystem.from_graph(ctx, n_prim, dofs_per_node, n_rhs, node_graph(n_prim, self.elem_prim))
        # element-local system: every element its own node ids, so its CRS rows are the rows of K_e
        n_loc = host.n_elems * nn
        local_nodes = np.arange(n_loc, dtype=np.uint32).reshape(host.n_elems, nn)
        self.local_mesh = Mesh(ctx, host.dim, host.order, host.verts, local_nodes, host.side_boundaries, n_loc, n_loc)
        node_ptr = np.arange(n_loc + 1, dtype=np.int64) * nn
        node_nbr = (np.repeat(np.arange(host.n_elems, dtype=np.uint32) * nn, nn * nn) + np.tile(np.arange(nn, dtype=np.uint32), n_loc))
        self.local = AssembledSystem(ctx, self.local_mesh, dofs_per_node, n_rhs, (node_ptr, np.ascontiguousarray(node_nbr, dtype=np.uint32)))
        self._h = C.c_void_p()
        elem_nodes = np.ascontiguousarray(host.nodes, dtype=np.uint32)
        ctx._chk(lib().l3b_cond_create(ctx._h, self.local._h, self.condensed._h, host.n_elems, nn, len(self.bnd_idx), _p(self.bnd_idx),
                                       len(self.int_idx), _p(self.int_idx), _p(self.elem_prim), _p(elem_nodes), C.byref(self._h)))
        self._fields = []

    def __del__(self):
        try:
            lib().l3b_cond_destroy(self._h)
        except Exception:
            pass

    @property
    def n_primary_dofs(self):
        return self.condensed.n_dofs

    def beginAssembly(self):
        self.local.beginAssembly()
        self.condensed.beginAssembly()
        self._fields = []

    def assembleProblem(self, kernel, boundary_ids=(), field_data=None, field_inds=None, dof_inds=None, asm_opts=AssemblyOptions(), time=0.0):
        """field_data: (n_fields, n_nodes) nodal values over the mesh's nodes (post/SolutionManager.hpp layout), re-indexed here to the
        element-local node ids"""
        fields = None
        if field_data is not None:
            fields = self.ctx.upload_fields(np.ascontiguousarray(np.asarray(field_data, dtype=np.float64)[:, self.host.nodes.ravel()]))
            self._fields.append(fields)
        self.local.assembleProblem(kernel, boundary_ids, fields, field_inds, dof_inds, asm_opts, time)

    def endAssembly(self, dirichlet_dofs=None, dirichlet_vals=None, n_owned_primary_dofs=None):
        """dirichlet_dofs in the numbering node * dofs_per_node + d of the mesh; they must sit on primary nodes (domain boundaries do).
        n_owned_primary_dofs: on a rank that also holds ghost rows (l3b_asm_end_assembly_ranked), in condensed numbering"""
        self.ctx._chk(lib().l3b_cond_condense(self._h))
        cd, vals = None, None
        if dirichlet_dofs is not None and len(dirichlet_dofs) > 0:
            d = np.asarray(dirichlet_dofs, dtype=np.int64)
            prim = self.prim_of[d // self.dofs_per_node]
            if (prim < 0).any():
                raise L3BError(2, "a Dirichlet dof sits on an element-interior node: it has no row in the condensed system")
            cd, vals = (prim * self.dofs_per_node + d % self.dofs_per_node).astype(np.int32), dirichlet_vals
        if n_owned_primary_dofs is None:
            self.condensed.endAssembly(cd, vals)
        else:
            self.condensed.endAssemblyRanked(cd, vals, n_owned_primary_dofs)

    def recover(self, x_condensed):
        """nodal solution over all mesh nodes, (n_nodes * dofs_per_node, n_rhs), from the solution of the condensed system"""
        xc = np.ascontiguousarray(np.asarray(x_condensed, dtype=np.float64).reshape(self.n_primary_dofs, -1).T)
        out = np.zeros((self.n_rhs, self.host.n_nodes * self.dofs_per_node))
        self.ctx._chk(lib().l3b_cond_recover(self._h, _p(xc), self.host.n_nodes, _p(out)))
        return out.T.copy()

    def solve(self, tol=1e-6, max_iters=10000, gmres=False):
        if self.n_rhs != 1:
            raise NotImplementedError("the Krylov drivers take one right-hand side: solve the columns of `condensed` yourself, then recover()")
        xc, achieved, iters = (self.condensed.solve_gmres if gmres else self.condensed.solve)(tol, max_iters=max_iters)
        return self.recover(xc)[:, 0], achieved, iters
