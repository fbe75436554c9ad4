"""l3ster_b200 — B200-native (sm_100a) implementation of L3STER's element-local least-squares assembly and matrix-free
operator hot path, behind the C ABI declared in ``include/l3ster_b200.h``.

This package is the thin Python host layer used by the tests and ``bench.py``: it loads ``libl3ster_b200.so`` with ctypes
and mirrors the reference's system API names (``beginAssembly / assembleProblem / endAssembly / solve`` of
``algsys/AssembledSystem.hpp:22-83`` and ``algsys/MatrixFreeSystem.hpp:47-120``).  There is no CPU fallback: creating a
:class:`Context` without a CUDA device raises, and a missing extension raises at import of :func:`lib`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
# L3B_LIB_PATH: load an experiment build (csrc/Makefile: EXTRA/BUILD/OUT) instead of the product library
LIB_PATH = os.environ.get("L3B_LIB_PATH") or os.path.join(_DIR, "libl3ster_b200.so")

QUAD, HEX = 2, 3
NO_BOUNDARY = 0xFFFF


class L3BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[l3b status {code}] {msg}")
        self.code, self.msg = code, msg


class _KernelInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("dimension", C.c_int), ("n_equations", C.c_int), ("n_unknowns", C.c_int), ("n_fields", C.c_int),
                ("n_rhs", C.c_int), ("is_boundary", C.c_int), ("n_instances", C.c_int), ("is_residual", C.c_int)]


class _AsmOpts(C.Structure):
    _fields_ = [("value_order", C.c_int), ("derivative_order", C.c_int), ("eval_strategy", C.c_int)]


@dataclass(frozen=True)
class AssemblyOptions:
    """algsys/AssembleLocalSystem.hpp:24-49"""
    value_order: int = 1
    derivative_order: int = 0
    eval_strategy: int = 0  # 0 Auto, 1 LocalElement, 2 SumFactorization, 3 SumFactorizationOddEvenDecomposition

    def _c(self):
        return _AsmOpts(self.value_order, self.derivative_order, self.eval_strategy)


# callbacks of l3b_pcg_device
APPLY_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
ALLREDUCE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
_lib = None
APPLY_INIT, APPLY_ELEMENTS, APPLY_FINISH, APPLY_BOUNDARY = 1, 2, 4, 8
COMM_ID_BYTES = 128

# every symbol include/l3ster_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "l3b_context_create", "l3b_context_destroy", "l3b_last_error", "l3b_global_error", "l3b_context_synchronize", "l3b_context_stream",
    "l3b_kernel_count", "l3b_kernel_find", "l3b_kernel_get_info", "l3b_kernel_get_instance",
    "l3b_tables_gll", "l3b_tables_gauss", "l3b_tables_1d", "l3b_tables_dense",
    "l3b_host_mesh_cube", "l3b_host_mesh_square", "l3b_host_mesh_destroy", "l3b_host_mesh_info", "l3b_host_mesh_nodes",
    "l3b_host_mesh_verts", "l3b_host_mesh_side_boundaries", "l3b_node_graph", "l3b_graph_expand", "l3b_free",
    "l3b_mesh_upload", "l3b_mesh_update_verts", "l3b_mesh_destroy", "l3b_fields_upload", "l3b_fields_update", "l3b_fields_destroy",
    "l3b_asm_create", "l3b_asm_destroy", "l3b_asm_nnz", "l3b_asm_begin_assembly", "l3b_asm_assemble", "l3b_asm_end_assembly",
    "l3b_asm_download", "l3b_asm_device_values", "l3b_asm_spmv", "l3b_asm_solve_cg", "l3b_asm_last_kernel_ms",
    "l3b_mf_create", "l3b_mf_destroy", "l3b_mf_assemble", "l3b_mf_end_assembly", "l3b_mf_download", "l3b_mf_apply_device", "l3b_mf_apply",
    "l3b_mf_solve_cg", "l3b_mf_num_dofs", "l3b_mf_kernel_launches", "l3b_microbench",
    "l3b_mf_apply_phase_device", "l3b_vec_gather", "l3b_vec_scatter_add",
    "l3b_mf_end_assembly_begin", "l3b_mf_end_assembly_finish", "l3b_mf_device_diag", "l3b_mf_device_rhs", "l3b_pcg_device",
    "l3b_asm_spmv_device", "l3b_asm_diag_device", "l3b_asm_device_rhs", "l3b_asm_end_assembly_ranked",
    "l3b_gmres_device", "l3b_asm_solve_gmres", "l3b_mf_solve_gmres",
    "l3b_compute_integral", "l3b_compute_norm_l2",
    "l3b_crs_create", "l3b_cond_create", "l3b_cond_destroy", "l3b_cond_condense", "l3b_cond_recover",
    "l3b_comm_unique_id", "l3b_comm_create", "l3b_comm_attach", "l3b_comm_destroy", "l3b_comm_rank", "l3b_comm_size", "l3b_comm_allreduce_sum",
    "l3b_halo_create", "l3b_halo_destroy", "l3b_halo_import_begin", "l3b_halo_import_end", "l3b_halo_export_begin", "l3b_halo_export_end",
    "l3b_mf_set_halo", "l3b_asm_set_halo", "l3b_mf_solve_device", "l3b_asm_solve_device", "l3b_mf_apply_energy_device",
    "l3b_partition_create", "l3b_partition_destroy", "l3b_partition_node_map", "l3b_partition_rank_info", "l3b_partition_rank_mesh",
    "l3b_partition_rank_halo", "l3b_partition_rank_graph", "l3b_asm_export_shared_rows",
    "l3b_mesh_set_element_domains", "l3b_dofmap_create", "l3b_dofmap_destroy", "l3b_dofmap_info", "l3b_dofmap_get", "l3b_asm_set_dofmap",
    "l3b_asm_download_compact", "l3b_compute_values_at_nodes", "l3b_update_solution", "l3b_fields_device",
    "l3b_mf_set_host_apply", "l3b_mf_host_apply_info", "l3b_host_apply_plan",
]


def lib():
    """Load the CUDA extension. Fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C l3ster_b200/csrc`). l3ster_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.l3b_last_error.restype = C.c_char_p
    L.l3b_last_error.argtypes = [vp]
    L.l3b_global_error.restype = C.c_char_p
    L.l3b_context_create.argtypes = [i32, C.POINTER(vp)]
    L.l3b_context_destroy.argtypes = [vp]
    L.l3b_context_destroy.restype = None
    L.l3b_context_synchronize.argtypes = [vp]
    L.l3b_context_stream.argtypes = [vp]
    L.l3b_context_stream.restype = vp
    L.l3b_kernel_find.argtypes = [C.c_char_p]
    L.l3b_kernel_get_info.argtypes = [i32, C.POINTER(_KernelInfo)]
    L.l3b_kernel_get_instance.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.l3b_tables_gll.argtypes = [i32, vp]
    L.l3b_tables_gauss.argtypes = [i32, vp, vp]
    L.l3b_tables_1d.argtypes = [i32, i32, vp, vp, vp]
    L.l3b_tables_dense.argtypes = [i32, i32, i32, i32, C.POINTER(i32), vp, vp, vp, vp]
    L.l3b_host_mesh_cube.argtypes = [i32, vp, i32, vp, i32, vp, i32, C.POINTER(vp)]
    L.l3b_host_mesh_square.argtypes = [i32, vp, i32, vp, i32, C.POINTER(vp)]
    L.l3b_host_mesh_destroy.argtypes = [vp]
    L.l3b_host_mesh_destroy.restype = None
    L.l3b_host_mesh_info.argtypes = [vp, vp]
    for f in ("l3b_host_mesh_nodes", "l3b_host_mesh_verts", "l3b_host_mesh_side_boundaries"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = vp
    L.l3b_node_graph.argtypes = [i64, i64, i32, vp, C.POINTER(vp), C.POINTER(vp)]
    L.l3b_graph_expand.argtypes = [i64, vp, vp, i32, vp, vp]
    L.l3b_free.argtypes = [vp]
    L.l3b_free.restype = None
    L.l3b_mesh_upload.argtypes = [vp, i32, i32, i64, vp, vp, vp, i64, i64, C.POINTER(vp)]
    L.l3b_mesh_update_verts.argtypes = [vp, vp]
    L.l3b_mesh_destroy.argtypes = [vp]
    L.l3b_mesh_destroy.restype = None
    L.l3b_fields_upload.argtypes = [vp, i64, i32, vp, C.POINTER(vp)]
    L.l3b_fields_update.argtypes = [vp, vp]
    L.l3b_fields_destroy.argtypes = [vp]
    L.l3b_fields_destroy.restype = None
    L.l3b_asm_create.argtypes = [vp, vp, i32, i32, vp, vp, C.POINTER(vp)]
    L.l3b_asm_destroy.argtypes = [vp]
    L.l3b_asm_destroy.restype = None
    L.l3b_asm_nnz.argtypes = [vp]
    L.l3b_asm_nnz.restype = i64
    L.l3b_asm_begin_assembly.argtypes = [vp]
    L.l3b_asm_assemble.argtypes = [vp, i32, _AsmOpts, dbl, vp, vp, vp, vp, i32]
    L.l3b_asm_end_assembly.argtypes = [vp, i64, vp, vp]
    L.l3b_asm_download.argtypes = [vp, vp, vp]
    L.l3b_asm_device_values.argtypes = [vp]
    L.l3b_asm_device_values.restype = vp
    L.l3b_asm_spmv.argtypes = [vp, vp, vp]
    L.l3b_asm_solve_cg.argtypes = [vp, dbl, i32, vp, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_asm_last_kernel_ms.argtypes = [vp]
    L.l3b_asm_last_kernel_ms.restype = dbl
    L.l3b_mf_create.argtypes = [vp, vp, i32, i32, vp, vp, C.POINTER(vp)]
    L.l3b_mf_destroy.argtypes = [vp]
    L.l3b_mf_destroy.restype = None
    L.l3b_mf_assemble.argtypes = [vp, i32, _AsmOpts, dbl, vp, vp, vp, vp, i32]
    L.l3b_crs_create.argtypes = [vp, i64, i32, i32, vp, vp, C.POINTER(vp)]
    L.l3b_cond_create.argtypes = [vp, vp, vp, i64, i32, i32, vp, i32, vp, vp, vp, C.POINTER(vp)]
    L.l3b_cond_destroy.argtypes = [vp]
    L.l3b_cond_condense.argtypes = [vp]
    L.l3b_cond_recover.argtypes = [vp, vp, i64, vp]
    for f in ("l3b_compute_integral", "l3b_compute_norm_l2"):
        getattr(L, f).argtypes = [vp, vp, i32, _AsmOpts, dbl, vp, vp, vp, i32, vp]
    L.l3b_mf_end_assembly.argtypes = [vp]
    L.l3b_mf_download.argtypes = [vp, vp, vp]
    L.l3b_mf_apply_device.argtypes = [vp, vp, vp, i32, dbl, dbl]
    L.l3b_mf_apply_phase_device.argtypes = [vp, vp, vp, i32, dbl, dbl, i32, i64, i64, vp]
    L.l3b_vec_gather.argtypes = [vp, vp, i64, vp, i64, i32, vp]
    L.l3b_mf_end_assembly_begin.argtypes = [vp]
    L.l3b_asm_solve_gmres.argtypes = [vp, dbl, i32, i32, i32, vp, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_mf_solve_gmres.argtypes = [vp, dbl, i32, i32, i32, vp, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_gmres_device.argtypes = [vp, i64, i64, APPLY_CB, ALLREDUCE_CB, vp, vp, vp, vp, dbl, i32, i32, i32, i32, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_asm_spmv_device.argtypes = [vp, vp, vp]
    L.l3b_asm_diag_device.argtypes = [vp, vp]
    L.l3b_asm_device_rhs.argtypes = [vp]
    L.l3b_asm_device_rhs.restype = vp
    L.l3b_asm_end_assembly_ranked.argtypes = [vp, i64, vp, vp, i64]
    L.l3b_mf_end_assembly_finish.argtypes = [vp]
    L.l3b_mf_device_diag.argtypes = [vp]
    L.l3b_mf_device_diag.restype = vp
    L.l3b_mf_device_rhs.argtypes = [vp]
    L.l3b_mf_device_rhs.restype = vp
    L.l3b_pcg_device.argtypes = [vp, i64, i64, APPLY_CB, ALLREDUCE_CB, vp, vp, vp, vp, dbl, i32, i32, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_comm_unique_id.argtypes = [vp]
    L.l3b_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.l3b_comm_attach.argtypes = [vp, vp, C.POINTER(vp)]
    L.l3b_comm_destroy.argtypes = [vp]
    L.l3b_comm_destroy.restype = None
    L.l3b_comm_rank.argtypes = [vp]
    L.l3b_comm_size.argtypes = [vp]
    L.l3b_comm_allreduce_sum.argtypes = [vp, vp, i32]
    L.l3b_halo_create.argtypes = [vp, i64, i64, i32, vp, vp, vp, i32, vp, vp, C.POINTER(vp)]
    L.l3b_halo_destroy.argtypes = [vp]
    L.l3b_halo_destroy.restype = None
    L.l3b_halo_import_begin.argtypes = [vp, vp, i32]
    L.l3b_halo_import_end.argtypes = [vp]
    L.l3b_halo_export_begin.argtypes = [vp, vp, i32]
    L.l3b_halo_export_end.argtypes = [vp, vp, i32]
    L.l3b_mf_set_halo.argtypes = [vp, vp, i64]
    L.l3b_asm_set_halo.argtypes = [vp, vp]
    L.l3b_mf_apply_energy_device.argtypes = [vp, vp, vp, dbl, dbl, vp]
    L.l3b_partition_create.argtypes = [i32, i32, i64, i64, vp, i32, vp, vp, C.POINTER(vp)]
    L.l3b_partition_destroy.argtypes = [vp]
    L.l3b_partition_destroy.restype = None
    L.l3b_partition_node_map.argtypes = [vp, vp, vp, vp]
    L.l3b_partition_rank_info.argtypes = [vp, i32, i32, vp]
    L.l3b_partition_rank_mesh.argtypes = [vp, i32, i32, vp, vp, vp]
    L.l3b_partition_rank_halo.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.l3b_partition_rank_graph.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.l3b_asm_export_shared_rows.argtypes = [vp, vp, vp]
    L.l3b_mesh_set_element_domains.argtypes = [vp, vp]
    L.l3b_dofmap_create.argtypes = [i32, i32, i64, i64, vp, vp, vp, i32, i32, vp, vp, vp, i64, C.POINTER(vp)]
    L.l3b_dofmap_destroy.argtypes = [vp]
    L.l3b_dofmap_destroy.restype = None
    L.l3b_dofmap_info.argtypes = [vp, vp]
    L.l3b_dofmap_get.argtypes = [vp, vp, vp, vp, vp]
    L.l3b_asm_set_dofmap.argtypes = [vp, vp]
    L.l3b_asm_download_compact.argtypes = [vp, vp, vp, vp]
    L.l3b_compute_values_at_nodes.argtypes = [vp, vp, i32, dbl, vp, vp, vp, i32, i32, vp, vp, i64, vp]
    L.l3b_update_solution.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.l3b_fields_device.argtypes = [vp]
    L.l3b_fields_device.restype = vp
    for f in ("l3b_mf_solve_device", "l3b_asm_solve_device"):
        getattr(L, f).argtypes = [vp, i32, dbl, i32, i32, i32, vp, i32, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_vec_scatter_add.argtypes = [vp, vp, i64, vp, i64, i32, vp]
    L.l3b_mf_apply.argtypes = [vp, vp, vp, i32, dbl, dbl]
    L.l3b_mf_set_host_apply.argtypes = [vp, i32, i32, i64]
    L.l3b_mf_host_apply_info.argtypes = [vp, vp]
    L.l3b_host_apply_plan.argtypes = [i64, i64, i64, i32, vp, i64, vp, i64, i64, i64, C.POINTER(i32)] + [C.POINTER(vp)] * 5
    L.l3b_mf_solve_cg.argtypes = [vp, dbl, i32, vp, C.POINTER(dbl), C.POINTER(i32)]
    L.l3b_mf_num_dofs.argtypes = [vp]
    L.l3b_mf_num_dofs.restype = i64
    L.l3b_mf_kernel_launches.argtypes = [vp]
    L.l3b_microbench.argtypes = [vp, i32, C.POINTER(dbl)]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _iarr(v, n=None):
    if v is None:
        return None
    a = np.ascontiguousarray(v, dtype=np.int32)
    if n is not None and len(a) != n:
        raise ValueError(f"expected {n} indices, got {len(a)}")
    return a


# ---------------------------------------------------------------------------------------------------------------------
def kernel_id(name: str) -> int:
    kid = lib().l3b_kernel_find(name.encode())
    if kid < 0:
        raise KeyError(f"kernel '{name}' is not registered")
    return kid


def kernel_info(name_or_id):
    kid = kernel_id(name_or_id) if isinstance(name_or_id, str) else name_or_id
    info = _KernelInfo()
    if lib().l3b_kernel_get_info(kid, C.byref(info)) != 0:
        raise KeyError(kid)
    inst = []
    for i in range(info.n_instances):
        o, q = C.c_int(), C.c_int()
        lib().l3b_kernel_get_instance(kid, i, C.byref(o), C.byref(q))
        inst.append((o.value, q.value))
    return dict(id=kid, name=info.name.decode(), dimension=info.dimension, n_equations=info.n_equations, n_unknowns=info.n_unknowns,
                n_fields=info.n_fields, n_rhs=info.n_rhs, is_boundary=bool(info.is_boundary), is_residual=bool(info.is_residual),
                instances=inst)


def list_kernels():
    return [kernel_info(i) for i in range(lib().l3b_kernel_count())]


# ---- tables (host-only entry points; usable without a GPU)
def tables_gll(n):
    out = np.zeros(n)
    lib().l3b_tables_gll(n, _p(out))
    return out


def tables_gauss(n):
    p, w = np.zeros(n), np.zeros(n)
    lib().l3b_tables_gauss(n, _p(p), _p(w))
    return p, w


def tables_1d(order, nq):
    a, b, c = np.zeros((order + 1, nq)), np.zeros((order + 1, nq)), np.zeros((nq, nq))
    lib().l3b_tables_1d(order, nq, _p(a), _p(b), _p(c))
    return a, b, c


def tables_dense(dim, order, nq, side=-1):
    nb = (order + 1) ** dim
    q = nq**dim if side < 0 else nq ** (dim - 1)
    pts, wts, vals, ders = np.zeros((q, dim)), np.zeros(q), np.zeros((q, nb)), np.zeros((q, dim, nb))
    n = C.c_int()
    rc = lib().l3b_tables_dense(dim, order, nq, side, C.byref(n), _p(pts), _p(wts), _p(vals), _p(ders))
    if rc != 0 or n.value != q:
        raise L3BError(rc, lib().l3b_global_error().decode())
    return pts, wts, vals, ders


# ---- host mesh front end
class HostMesh:
    """Structured quad/hex mesh with the reference's order-p node numbering (mesh/primitives/*, ConvertMeshToOrder.hpp)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        info = np.zeros(6, dtype=np.int64)
        lib().l3b_host_mesh_info(self._h, _p(info))
        self.dim, self.order, self.n_nodes, self.n_elems, self.nodes_per_elem, self.n_sides = map(int, info)
        L = lib()

        def view(ptr, dtype, shape):
            n = int(np.prod(shape))
            buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

        self.nodes = view(L.l3b_host_mesh_nodes(self._h), np.uint32, (self.n_elems, self.nodes_per_elem))
        self.verts = view(L.l3b_host_mesh_verts(self._h), np.float64, (self.n_elems, 2**self.dim, 3))
        self.side_boundaries = view(L.l3b_host_mesh_side_boundaries(self._h), np.uint16, (self.n_elems, self.n_sides))

    def __del__(self):
        try:
            lib().l3b_host_mesh_destroy(self._h)
        except Exception:
            pass

    def node_graph(self):
        return node_graph(self.n_nodes, self.nodes)

    def boundary_nodes(self, boundary_ids):
        """Local node ids on the sides carrying one of `boundary_ids` (bcs/LocalDirichletBC.hpp:86-105)."""
        nb = self.order + 1
        sel = np.zeros(self.n_nodes, dtype=bool)
        for side in range(self.n_sides):
            on = np.isin(self.side_boundaries[:, side], list(boundary_ids))
            if on.any():
                sel[self.nodes[on][:, side_node_inds(self.dim, self.order, side)].ravel()] = True
        return np.nonzero(sel)[0]


def side_node_inds(dim, order, side):
    """mesh/ElementTraits.hpp:72-98, 118-137"""
    n = order + 1
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    i, j = i.ravel(), j.ravel()
    if dim == 3:
        nps = n * n
        return [i * n + j, i * n + j + nps * (n - 1), i * nps + j, i * nps + j + n * (n - 1), i * nps + j * n, i * nps + j * n + n - 1][side]
    k = np.arange(n)
    return [k, k + n * (n - 1), k * n, k * n + n - 1][side]


def make_cube_mesh(x, y=None, z=None, order=1) -> HostMesh:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = x if y is None else np.ascontiguousarray(y, dtype=np.float64)
    z = x if z is None else np.ascontiguousarray(z, dtype=np.float64)
    h = C.c_void_p()
    rc = lib().l3b_host_mesh_cube(len(x), _p(x), len(y), _p(y), len(z), _p(z), order, C.byref(h))
    if rc != 0:
        raise L3BError(rc, lib().l3b_global_error().decode())
    return HostMesh(h.value)


def make_square_mesh(x, y=None, order=1) -> HostMesh:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = x if y is None else np.ascontiguousarray(y, dtype=np.float64)
    h = C.c_void_p()
    rc = lib().l3b_host_mesh_square(len(x), _p(x), len(y), _p(y), order, C.byref(h))
    if rc != 0:
        raise L3BError(rc, lib().l3b_global_error().decode())
    return HostMesh(h.value)


def node_graph(n_nodes, nodes):
    """Node-level sparsity graph (algsys/SparsityGraph.hpp:25-81): returns (ptr, nbr) numpy arrays."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    ptr, nbr = C.c_void_p(), C.c_void_p()
    rc = lib().l3b_node_graph(n_nodes, nodes.shape[0], nodes.shape[1], _p(nodes), C.byref(ptr), C.byref(nbr))
    if rc != 0:
        raise L3BError(rc, lib().l3b_global_error().decode())
    try:
        p = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(n_nodes + 1,)).copy()
        n = np.ctypeslib.as_array(C.cast(nbr, C.POINTER(C.c_uint32)), shape=(max(int(p[-1]), 1),)).copy()[: int(p[-1])]
    finally:
        lib().l3b_free(ptr)
        lib().l3b_free(nbr)
    return p, n


def host_apply_plan(n_nodes, nodes, chunk_elems, block_nodes, n_owned_nodes=None, n_border_elems=0, halo_nodes=None):
    """Schedule of the streamed host-buffer apply (csrc/apply_plan_host.hpp): returns (item_elems [n_items, 2], up, down) with up / down
    = per item the list of node ranges (begin, end) to copy in before / out after the item."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, np.shape(nodes)[-1] if np.ndim(nodes) == 2 else 1)
    halo = np.zeros(0, dtype=np.int32) if halo_nodes is None else np.ascontiguousarray(halo_nodes, dtype=np.int32)
    n_items = C.c_int()
    out = [C.c_void_p() for _ in range(5)]
    rc = lib().l3b_host_apply_plan(n_nodes, n_nodes if n_owned_nodes is None else n_owned_nodes, nodes.shape[0], nodes.shape[1], _p(nodes),
                                   n_border_elems, _p(halo), len(halo), chunk_elems, block_nodes, C.byref(n_items), *[C.byref(o) for o in out])
    if rc != 0:
        raise L3BError(rc, lib().l3b_global_error().decode())
    try:
        k = n_items.value
        arr = lambda ptr, n: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(max(n, 1),)).copy()[:n]  # noqa: E731
        items = arr(out[0], 2 * k).reshape(k, 2)
        up_ptr, down_ptr = arr(out[1], k + 1), arr(out[3], k + 1)
        up_r, down_r = arr(out[2], 2 * int(up_ptr[-1])).reshape(-1, 2), arr(out[4], 2 * int(down_ptr[-1])).reshape(-1, 2)
    finally:
        for o in out:
            lib().l3b_free(o)
    up = [[tuple(map(int, r)) for r in up_r[up_ptr[i]:up_ptr[i + 1]]] for i in range(k)]
    down = [[tuple(map(int, r)) for r in down_r[down_ptr[i]:down_ptr[i + 1]]] for i in range(k)]
    return items, up, down


def expand_graph(ptr, nbr, dofs_per_node, with_cols=True):
    """dof-level CRS (row_ptr, col_ind) of the reference's Tpetra graph (SparsityGraph.hpp:254-278)."""
    n_nodes = len(ptr) - 1
    row_ptr = np.zeros(n_nodes * dofs_per_node + 1, dtype=np.int64)
    col_ind = np.zeros(int(ptr[-1]) * dofs_per_node**2, dtype=np.int32) if with_cols else None
    lib().l3b_graph_expand(n_nodes, _p(ptr), _p(nbr), dofs_per_node, _p(row_ptr), _p(col_ind))
    return row_ptr, col_ind


# ---------------------------------------------------------------------------------------------------------------------
class Context:
    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().l3b_context_create(device, C.byref(self._h))
        if rc != 0:
            raise L3BError(rc, lib().l3b_global_error().decode())

    def _chk(self, rc):
        if rc != 0:
            raise L3BError(rc, lib().l3b_last_error(self._h).decode())

    def synchronize(self):
        self._chk(lib().l3b_context_synchronize(self._h))

    @property
    def stream(self):
        return lib().l3b_context_stream(self._h)

    def __del__(self):
        try:
            lib().l3b_context_destroy(self._h)
        except Exception:
            pass

    def microbench(self, mode):
        """0: fp64 FMA TFLOP/s, 1: fp64 DMMA TFLOP/s, 2: HBM copy GB/s"""
        out = C.c_double()
        self._chk(lib().l3b_microbench(self._h, mode, C.byref(out)))
        return out.value

    def _krylov_callbacks(self, apply, allreduce):
        err = []

        def _apply(_, x, y, e):
            try:
                return 1 if apply(x, y, e) is True else 0
            except Exception as exc:  # surfaces as an L3BError with the message kept
                err.append(exc)
                return 2

        def _reduce(_, s, n):
            try:
                allreduce(s, n)
                return 0
            except Exception as exc:
                err.append(exc)
                return 1

        return APPLY_CB(_apply), (ALLREDUCE_CB(_reduce) if allreduce is not None else C.cast(None, ALLREDUCE_CB)), err

    def pcg(self, n_local, n_owned, apply, allreduce, diag_ptr, b_ptr, x_ptr, tol=1e-6, max_iters=10000, x0_is_zero=True):
        """l3b_pcg_device: apply(x_ptr, y_ptr, energy_ptr) and allreduce(scalars_ptr, n) are Python callables working on device pointers;
        apply returns True when it added this rank's share of x^T A x to the device scalar at energy_ptr (see l3b_apply_callback);
        x0_is_zero=False: the vector at x_ptr is the initial guess"""
        a_cb, r_cb, err = self._krylov_callbacks(apply, allreduce)
        at, it = C.c_double(), C.c_int()
        rc = lib().l3b_pcg_device(self._h, n_local, n_owned, a_cb, r_cb, None, diag_ptr, b_ptr, x_ptr, tol, max_iters, int(x0_is_zero),
                                  C.byref(at), C.byref(it))
        if rc != 0:
            raise L3BError(rc, lib().l3b_last_error(self._h).decode() + (f" ({err[0]!r})" if err else ""))
        return at.value, it.value

    def gmres(self, n_local, n_owned, apply, allreduce, diag_ptr, b_ptr, x_ptr, tol=1e-6, restart_length=250, max_restarts=39, max_iters=10000,
              x0_is_zero=True):
        """l3b_gmres_device with the same callbacks as pcg (the energy pointer is always null here)"""
        a_cb, r_cb, err = self._krylov_callbacks(apply, allreduce)
        at, it = C.c_double(), C.c_int()
        rc = lib().l3b_gmres_device(self._h, n_local, n_owned, a_cb, r_cb, None, diag_ptr, b_ptr, x_ptr, tol, restart_length, max_restarts,
                                    max_iters, int(x0_is_zero), C.byref(at), C.byref(it))
        if rc != 0:
            raise L3BError(rc, lib().l3b_last_error(self._h).decode() + (f" ({err[0]!r})" if err else ""))
        return at.value, it.value

    def vec_gather(self, src_ptr, ld, idx_ptr, n, dst_ptr, n_cols=1):
        self._chk(lib().l3b_vec_gather(self._h, src_ptr, ld, idx_ptr, n, n_cols, dst_ptr))

    def vec_scatter_add(self, dst_ptr, ld, idx_ptr, n, src_ptr, n_cols=1):
        self._chk(lib().l3b_vec_scatter_add(self._h, dst_ptr, ld, idx_ptr, n, n_cols, src_ptr))

    def upload_mesh(self, mesh: HostMesh, n_owned_nodes=None):
        return Mesh(self, mesh.dim, mesh.order, mesh.verts, mesh.nodes, mesh.side_boundaries, mesh.n_nodes,
                    mesh.n_nodes if n_owned_nodes is None else n_owned_nodes)

    def upload_fields(self, data):
        """data: (n_fields, n_local_nodes) — post/SolutionManager.hpp:83-101 layout"""
        return Fields(self, data)


class DofMap:
    """ProblemDefinition -> NodeToGlobalDofMap -> sparsity graph (l3b_dofmap_*): `definitions` is a list of (domain ids, dof indices) —
    ProblemDefinition::define(domains, dofs). Attributes: active (n_nodes, dpn) bool, dof (n_nodes, dpn) compact ids or -1, n_dofs,
    row_ptr / col_ind (the reference's compact CRS graph)."""

    def __init__(self, dim, order, n_nodes, nodes, elem_domains, side_boundaries, dofs_per_node, definitions, base_dof=0):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        ed = None if elem_domains is None else np.ascontiguousarray(elem_domains, dtype=np.int32)
        sb = None if side_boundaries is None else np.ascontiguousarray(side_boundaries, dtype=np.uint16)
        ptr = np.zeros(len(definitions) + 1, dtype=np.int32)
        ids, masks = [], []
        for i, (doms, dofs) in enumerate(definitions):
            ids += list(doms)
            ptr[i + 1] = len(ids)
            masks.append(sum(1 << int(d) for d in dofs))
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        masks = np.ascontiguousarray(masks, dtype=np.uint32)
        self._h = C.c_void_p()
        rc = lib().l3b_dofmap_create(dim, order, n_nodes, nodes.shape[0], _p(nodes), _p(ed), _p(sb), dofs_per_node, len(definitions), _p(ptr),
                                     _p(ids), _p(masks), base_dof, C.byref(self._h))
        if rc != 0:
            raise L3BError(rc, lib().l3b_global_error().decode())
        info = np.zeros(4, dtype=np.int64)
        lib().l3b_dofmap_info(self._h, _p(info))
        self.n_dofs, nnz, self.n_nodes, self.dofs_per_node = map(int, info)
        active = np.zeros((self.n_nodes, dofs_per_node), dtype=np.uint8)
        self.dof = np.zeros((self.n_nodes, dofs_per_node), dtype=np.int64)
        self.row_ptr = np.zeros(self.n_dofs + 1, dtype=np.int64)
        self.col_ind = np.zeros(nnz, dtype=np.int32)
        lib().l3b_dofmap_get(self._h, _p(active), _p(self.dof), _p(self.row_ptr), _p(self.col_ind))
        self.active = active.astype(bool)
        self.base_dof = base_dof

    def padded_dirichlet(self, mask=None, vals=None, n_rhs=1):
        """Dirichlet mask / values over the PADDED dofs with the inactive pairs closed (mask 1, value 0): what MatrixFreeSystem takes"""
        m = np.zeros(self.n_nodes * self.dofs_per_node, dtype=np.uint8) if mask is None else np.array(mask, dtype=np.uint8).ravel()
        v = np.zeros((len(m), n_rhs)) if vals is None else np.array(vals, dtype=np.float64).reshape(len(m), -1)
        inactive = ~self.active.ravel()
        m[inactive] = 1
        v[inactive] = 0.0
        return m, v

    def compact(self, padded):
        """padded vector(s) (n_nodes * dpn [, n_cols]) -> compact (n_dofs [, n_cols])"""
        padded = np.asarray(padded)
        return padded.reshape(self.n_nodes * self.dofs_per_node, -1)[self.active.ravel()].reshape((self.n_dofs,) + padded.shape[1:])

    def __del__(self):
        try:
            lib().l3b_dofmap_destroy(self._h)
        except Exception:
            pass


class Comm:
    """The rank's communicator (l3b_comm_*): NCCL, one rank per GPU. `Comm(ctx)` is the single-rank communicator;
    `Comm.from_torch_distributed(ctx)` creates one over the ranks of the initialised torch.distributed group (the unique id is
    drawn by rank 0 and broadcast through torch.distributed — the only thing torch carries)."""

    def __init__(self, ctx: Context, rank=0, world=1, unique_id=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        if unique_id is None:
            if world != 1:
                raise ValueError("more than one rank needs the unique id drawn by rank 0 (Comm.unique_id())")
            unique_id = Comm.unique_id()
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._h = C.c_void_p()
        ctx._chk(lib().l3b_comm_create(ctx._h, rank, world, buf, C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = lib().l3b_comm_unique_id(buf)
        if rc != 0:
            raise L3BError(rc, lib().l3b_global_error().decode())
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, ctx: Context):
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(ctx)
        box = [Comm.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, dist.get_rank(), dist.get_world_size(), box[0])

    def allreduce_sum(self, scalars_ptr, n):
        self.ctx._chk(lib().l3b_comm_allreduce_sum(self._h, scalars_ptr, n))

    def __del__(self):
        try:
            lib().l3b_comm_destroy(self._h)
        except Exception:
            pass


class DeviceHalo:
    """comm::ImportExportContext + Import + Export on the device (l3b_halo_*). owned: list of (rank, local owned dofs shared with it, in
    the neighbour's ghost order); shared: list of (rank, offset, size) ranges of the ghost block owned by that rank."""

    def __init__(self, comm: Comm, n_owned, n_ghost, owned, shared):
        self.comm, self.n_owned, self.n_ghost = comm, int(n_owned), int(n_ghost)
        o_ranks = np.array([r for r, _ in owned], dtype=np.int32)
        o_ptr = np.zeros(len(owned) + 1, dtype=np.int64)
        for k, (_, idx) in enumerate(owned):
            o_ptr[k + 1] = o_ptr[k] + len(idx)
        o_idx = np.concatenate([np.asarray(i, dtype=np.int32) for _, i in owned]) if owned else np.zeros(0, dtype=np.int32)
        s_ranks = np.array([r for r, _, _ in shared], dtype=np.int32)
        s_off = np.zeros(len(shared) + 1, dtype=np.int64)
        for j, (_, off, size) in enumerate(shared):
            if off != s_off[j]:
                raise ValueError("the ghost ranges of the shared neighbours must tile the ghost block in order")
            s_off[j + 1] = off + size
        self._h = C.c_void_p()
        comm.ctx._chk(lib().l3b_halo_create(comm._h, self.n_owned, self.n_ghost, len(owned), _p(o_ranks), _p(o_ptr), _p(np.ascontiguousarray(o_idx)),
                                            len(shared), _p(s_ranks), _p(s_off), C.byref(self._h)))

    def import_begin(self, x_ptr, n_cols=1):
        self.comm.ctx._chk(lib().l3b_halo_import_begin(self._h, x_ptr, n_cols))

    def import_end(self):
        self.comm.ctx._chk(lib().l3b_halo_import_end(self._h))

    def export_begin(self, y_ptr, n_cols=1):
        self.comm.ctx._chk(lib().l3b_halo_export_begin(self._h, y_ptr, n_cols))

    def export_end(self, y_ptr, n_cols=1):
        self.comm.ctx._chk(lib().l3b_halo_export_end(self._h, y_ptr, n_cols))

    def import_(self, x_ptr, n_cols=1):
        self.import_begin(x_ptr, n_cols)
        self.import_end()

    def export_add(self, y_ptr, n_cols=1):
        self.export_begin(y_ptr, n_cols)
        self.export_end(y_ptr, n_cols)

    def __del__(self):
        try:
            lib().l3b_halo_destroy(self._h)
        except Exception:
            pass


class Mesh:
    def _integrate(self, fn, kernel, boundary_ids, fields, field_inds, asm_opts, time):
        kid, _, fh, fi, bi, nb = _kernel_args(kernel, None, fields, field_inds, boundary_ids)
        info = kernel_info(kid)
        out = np.zeros(info["n_equations"] * info["n_rhs"])
        self.ctx._chk(fn(self.ctx._h, self._h, kid, asm_opts._c(), time, fh, _p(fi), _p(bi), nb, _p(out)))
        return out

    def computeIntegral(self, kernel, boundary_ids=(), fields=None, field_inds=None, asm_opts=AssemblyOptions(), time=0.0):
        """computeIntegral (post/Integral.hpp:102-121) of a residual kernel over this rank's elements (domain kernel) or over the sides
        carrying `boundary_ids` (boundary kernel); the caller sums over ranks"""
        return self._integrate(lib().l3b_compute_integral, kernel, boundary_ids, fields, field_inds, asm_opts, time)

    def computeNormL2(self, kernel, boundary_ids=(), fields=None, field_inds=None, asm_opts=AssemblyOptions(), time=0.0):
        """computeNormL2 (post/NormL2.hpp:31-60); on more than one rank sum the squares over the ranks"""
        return self._integrate(lib().l3b_compute_norm_l2, kernel, boundary_ids, fields, field_inds, asm_opts, time)

    def computeValuesAtNodes(self, kernel, values_ptr, dofs_per_node, ids=(), dof_inds=None, fields=None, field_inds=None, time=0.0, ld=None,
                             halo=None):
        """l3b_compute_values_at_nodes: nodal values of a residual kernel into the (device, padded) vector at values_ptr"""
        info = kernel_info(kernel)
        di = _iarr(dof_inds, info["n_equations"])
        fi = _iarr(field_inds, info["n_fields"])
        bi = _iarr(list(ids))
        ld = self.n_local_nodes * dofs_per_node if ld is None else ld
        self.ctx._chk(lib().l3b_compute_values_at_nodes(self.ctx._h, self._h, info["id"], time, fields._h if fields is not None else None, _p(fi),
                                                        _p(bi), 0 if bi is None else len(bi), dofs_per_node, _p(di), values_ptr, ld,
                                                        halo._h if halo is not None else None))

    def set_element_domains(self, domain_ids):
        """element -> domain id: domain kernels can then be restricted to domains (assembleProblem(kernel, domain_ids, ...))"""
        d = None if domain_ids is None else np.ascontiguousarray(domain_ids, dtype=np.int32)
        self.ctx._chk(lib().l3b_mesh_set_element_domains(self._h, _p(d)))

    def update_verts(self, verts):
        """new vertex coordinates, same connectivity (H2D copy, no reallocation)"""
        self.ctx._chk(lib().l3b_mesh_update_verts(self._h, verts.ctypes.data))

    def __init__(self, ctx: Context, dim, order, verts, nodes, side_boundaries, n_local_nodes, n_owned_nodes):
        self.ctx, self.dim, self.order = ctx, dim, order
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        sb = None if side_boundaries is None else np.ascontiguousarray(side_boundaries, dtype=np.uint16)
        self.n_elems, self.nodes_per_elem = nodes.shape
        self.n_local_nodes = int(n_local_nodes)
        self.nodes = nodes
        self._h = C.c_void_p()
        ctx._chk(lib().l3b_mesh_upload(ctx._h, dim, order, self.n_elems, _p(verts), _p(nodes), _p(sb), n_local_nodes, n_owned_nodes,
                                       C.byref(self._h)))

    def __del__(self):
        try:
            lib().l3b_mesh_destroy(self._h)
        except Exception:
            pass


class Fields:
    def __init__(self, ctx: Context, data):
        data = np.ascontiguousarray(data, dtype=np.float64)
        self.ctx, self.n_fields, self.n_nodes = ctx, data.shape[0], data.shape[1]
        self._h = C.c_void_p()
        ctx._chk(lib().l3b_fields_upload(ctx._h, self.n_nodes, self.n_fields, _p(data), C.byref(self._h)))

    def update(self, data):
        data = np.ascontiguousarray(data, dtype=np.float64)
        self.ctx._chk(lib().l3b_fields_update(self._h, _p(data)))

    @property
    def device_ptr(self):
        return lib().l3b_fields_device(self._h)

    def update_from_solution(self, x_ptr, dofs_per_node, dof_inds, field_inds):
        """updateSolution on the device: field field_inds[i] <- dof dof_inds[i] of the solution vector at x_ptr"""
        di, fi = _iarr(dof_inds), _iarr(field_inds, len(dof_inds))
        self.ctx._chk(lib().l3b_update_solution(self.ctx._h, x_ptr, dofs_per_node, _p(di), len(di), self._h, _p(fi)))

    def __del__(self):
        try:
            lib().l3b_fields_destroy(self._h)
        except Exception:
            pass


def _kernel_args(kernel, dof_inds, fields, field_inds, boundary_ids):
    info = kernel_info(kernel)
    di = _iarr(dof_inds, info["n_unknowns"])
    fi = _iarr(field_inds, info["n_fields"])
    bi = _iarr(list(boundary_ids))
    return info["id"], di, (fields._h if fields is not None else None), fi, bi, (0 if bi is None else len(bi))


class AssembledSystem:
    """algsys/AssembledSystem.hpp:22-83 on the device: node-block CRS values + rhs."""

    def __init__(self, ctx: Context, mesh: Mesh, dofs_per_node, n_rhs=1, graph=None):
        self.ctx, self.mesh, self.dofs_per_node, self.n_rhs = ctx, mesh, dofs_per_node, n_rhs
        self.node_ptr, self.node_nbr = graph if graph is not None else node_graph(mesh.n_local_nodes, mesh.nodes)
        self.n_dofs = mesh.n_local_nodes * dofs_per_node
        self._halo = None
        self._h = C.c_void_p()
        ctx._chk(lib().l3b_asm_create(ctx._h, mesh._h, dofs_per_node, n_rhs, _p(self.node_ptr), _p(self.node_nbr), C.byref(self._h)))
        self.nnz = int(lib().l3b_asm_nnz(self._h))

    @classmethod
    def from_graph(cls, ctx: Context, n_nodes, dofs_per_node, n_rhs, graph):
        """Storage over a node graph without a mesh (l3b_crs_create): the condensed system of static condensation"""
        self = cls.__new__(cls)
        self.ctx, self.mesh, self.dofs_per_node, self.n_rhs = ctx, None, dofs_per_node, n_rhs
        self.node_ptr, self.node_nbr = graph
        self.n_dofs = n_nodes * dofs_per_node
        self._halo = None
        self._h = C.c_void_p()
        ctx._chk(lib().l3b_crs_create(ctx._h, n_nodes, dofs_per_node, n_rhs, _p(self.node_ptr), _p(self.node_nbr), C.byref(self._h)))
        self.nnz = int(lib().l3b_asm_nnz(self._h))
        return self

    def __del__(self):
        try:
            lib().l3b_asm_destroy(self._h)
        except Exception:
            pass

    def beginAssembly(self):
        self.ctx._chk(lib().l3b_asm_begin_assembly(self._h))

    def assembleProblem(self, kernel, boundary_ids=(), fields=None, field_inds=None, dof_inds=None, asm_opts=AssemblyOptions(), time=0.0):
        kid, di, fh, fi, bi, nb = _kernel_args(kernel, dof_inds, fields, field_inds, boundary_ids)
        self._keepalive = fields  # the launch is asynchronous w.r.t. Python object lifetimes
        self.ctx._chk(lib().l3b_asm_assemble(self._h, kid, asm_opts._c(), time, _p(di), fh, _p(fi), _p(bi), nb))

    def endAssembly(self, dirichlet_dofs=None, dirichlet_vals=None):
        n = 0 if dirichlet_dofs is None else len(dirichlet_dofs)
        d = None if n == 0 else np.ascontiguousarray(dirichlet_dofs, dtype=np.int32)
        v = None if n == 0 else np.ascontiguousarray(np.asarray(dirichlet_vals, dtype=np.float64).reshape(n, -1).T)
        self.ctx._chk(lib().l3b_asm_end_assembly(self._h, n, _p(d), _p(v)))

    def solve_device(self, x_ptr, method="cg", tol=1e-6, max_iters=10000, restart_length=250, max_restarts=39, x0_is_zero=True):
        """l3b_asm_solve_device: solution on the device over the local rows"""
        at, it = C.c_double(), C.c_int()
        self.ctx._chk(lib().l3b_asm_solve_device(self._h, {"cg": 0, "gmres": 1}[method], tol, max_iters, restart_length, max_restarts, x_ptr,
                                                 int(x0_is_zero), C.byref(at), C.byref(it)))
        return at.value, it.value

    def set_dofmap(self, dofmap: "DofMap | None"):
        """inactive (node, dof) pairs are closed as identity rows at endAssembly"""
        self._dofmap = dofmap
        self.ctx._chk(lib().l3b_asm_set_dofmap(self._h, dofmap._h if dofmap is not None else None))

    def download_compact(self, dofmap: "DofMap"):
        """(values over dofmap's compact CRS graph, rhs (n_dofs, n_rhs)): the reference's matrix"""
        vals = np.zeros(len(dofmap.col_ind))
        rhs = np.zeros((self.n_rhs, dofmap.n_dofs))
        self.ctx._chk(lib().l3b_asm_download_compact(self._h, dofmap._h, _p(vals), _p(rhs)))
        return vals, rhs.T.copy()

    def export_shared_rows(self, recv_entry_ptr, recv_pos):
        """l3b_asm_export_shared_rows: ghost-row values and the ghost block of the rhs go to their owners (AssembledSystem.hpp:384-389)"""
        ep = np.ascontiguousarray(recv_entry_ptr, dtype=np.int64)
        rp = np.ascontiguousarray(recv_pos, dtype=np.uint32)
        self.ctx._chk(lib().l3b_asm_export_shared_rows(self._h, _p(ep), _p(rp)))

    def set_halo(self, halo: "DeviceHalo | None"):
        """rows [owned | ghost] over more than one rank: spmv_device, diag_device and the solvers become the global operator"""
        self._halo = halo
        self.ctx._chk(lib().l3b_asm_set_halo(self._h, halo._h if halo is not None else None))

    def endAssemblyRanked(self, dirichlet_dofs, dirichlet_vals, n_owned_dofs):
        """endAssembly on a rank that also holds ghost rows: ghost copies of Dirichlet rows become zero rows"""
        n = 0 if dirichlet_dofs is None else len(dirichlet_dofs)
        d = None if n == 0 else np.ascontiguousarray(dirichlet_dofs, dtype=np.int32)
        v = None if n == 0 else np.ascontiguousarray(np.asarray(dirichlet_vals, dtype=np.float64).reshape(n, -1).T)
        self.ctx._chk(lib().l3b_asm_end_assembly_ranked(self._h, n, _p(d), _p(v), n_owned_dofs))

    def spmv_device(self, x_ptr, y_ptr):
        self.ctx._chk(lib().l3b_asm_spmv_device(self._h, x_ptr, y_ptr))

    def diag_device(self, diag_ptr):
        self.ctx._chk(lib().l3b_asm_diag_device(self._h, diag_ptr))

    @property
    def device_rhs(self):
        return lib().l3b_asm_device_rhs(self._h)

    def graph(self):
        return expand_graph(self.node_ptr, self.node_nbr, self.dofs_per_node)

    def download(self, values=True):
        vals = np.zeros(self.nnz) if values else None
        rhs = np.zeros((self.n_rhs, self.n_dofs))
        self.ctx._chk(lib().l3b_asm_download(self._h, _p(vals), _p(rhs)))
        return vals, rhs.T.copy()

    def download_rhs_into(self, host_ptr):
        """rhs (n_rhs x n_dofs, C ABI layout) into a caller-owned (e.g. pinned) host buffer"""
        self.ctx._chk(lib().l3b_asm_download(self._h, None, host_ptr))

    def getMatrix(self):
        import scipy.sparse as sp
        row_ptr, col_ind = self.graph()
        vals, _ = self.download()
        return sp.csr_matrix((vals, col_ind, row_ptr), shape=(self.n_dofs, self.n_dofs))

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self.ctx._chk(lib().l3b_asm_spmv(self._h, _p(x), _p(y)))
        return y

    def solve(self, tol=1e-6, max_iters=10000, x0=None):
        """x0: initial guess (default zero); the reference starts a repeated solve from its previous solution"""
        x = np.zeros(max(self.n_dofs, 1)) if x0 is None else np.array(x0, dtype=np.float64, copy=True).reshape(-1)
        at, it = C.c_double(), C.c_int()
        self.ctx._chk(lib().l3b_asm_solve_cg(self._h, tol, max_iters, _p(x), C.byref(at), C.byref(it)))
        return x[:self.n_dofs], at.value, it.value

    def solve_gmres(self, tol=1e-6, restart_length=250, max_restarts=39, max_iters=10000, x0=None):
        """lstr::Gmres with the native Jacobi preconditioner (solve/BelosSolvers.hpp:125-131, SolverInterface.hpp:26-37 defaults)"""
        x = np.zeros(max(self.n_dofs, 1)) if x0 is None else np.array(x0, dtype=np.float64, copy=True).reshape(-1)
        at, it = C.c_double(), C.c_int()
        self.ctx._chk(lib().l3b_asm_solve_gmres(self._h, tol, restart_length, max_restarts, max_iters, _p(x), C.byref(at), C.byref(it)))
        return x[:self.n_dofs], at.value, it.value

    @property
    def last_kernel_ms(self):
        return lib().l3b_asm_last_kernel_ms(self._h)


class MatrixFreeSystem:
    """algsys/MatrixFreeSystem.hpp:47-120 on the device."""

    def __init__(self, ctx: Context, mesh: Mesh, dofs_per_node, n_rhs=1, dirichlet_mask=None, dirichlet_vals=None):
        self.ctx, self.mesh, self.dofs_per_node, self.n_rhs = ctx, mesh, dofs_per_node, n_rhs
        self.n_dofs = mesh.n_local_nodes * dofs_per_node
        m = None if dirichlet_mask is None else np.ascontiguousarray(dirichlet_mask, dtype=np.uint8)
        v = None
        if dirichlet_vals is not None:
            v = np.ascontiguousarray(np.asarray(dirichlet_vals, dtype=np.float64).reshape(self.n_dofs, -1).T)
        self._h = C.c_void_p()
        self._keepalive = []
        ctx._chk(lib().l3b_mf_create(ctx._h, mesh._h, dofs_per_node, n_rhs, _p(m), _p(v), C.byref(self._h)))

    def __del__(self):
        try:
            lib().l3b_mf_destroy(self._h)
        except Exception:
            pass

    def assembleProblem(self, kernel, boundary_ids=(), fields=None, field_inds=None, dof_inds=None, asm_opts=AssemblyOptions(), time=0.0):
        kid, di, fh, fi, bi, nb = _kernel_args(kernel, dof_inds, fields, field_inds, boundary_ids)
        self._keepalive.append(fields)  # the system stores the device pointer (like FieldAccess references SolutionManager)
        self.ctx._chk(lib().l3b_mf_assemble(self._h, kid, asm_opts._c(), time, _p(di), fh, _p(fi), _p(bi), nb))

    def solve_device(self, x_ptr, method="cg", tol=1e-6, max_iters=10000, restart_length=250, max_restarts=39, x0_is_zero=True):
        """l3b_mf_solve_device: solution on the device over the local dofs"""
        at, it = C.c_double(), C.c_int()
        self.ctx._chk(lib().l3b_mf_solve_device(self._h, {"cg": 0, "gmres": 1}[method], tol, max_iters, restart_length, max_restarts, x_ptr,
                                                int(x0_is_zero), C.byref(at), C.byref(it)))
        return at.value, it.value

    def set_halo(self, halo: "DeviceHalo | None", n_border_elems=0):
        """more than one rank (l3b_mf_set_halo): endAssembly, apply and the solvers then work over all ranks"""
        self._halo = halo
        self.ctx._chk(lib().l3b_mf_set_halo(self._h, halo._h if halo is not None else None, n_border_elems))

    def endAssembly(self):
        self.ctx._chk(lib().l3b_mf_end_assembly(self._h))

    def endAssemblyBegin(self):
        """element contributions to diag / rhs over [owned | ghost]; export-add the ghost parts, then endAssemblyFinish()"""
        self.ctx._chk(lib().l3b_mf_end_assembly_begin(self._h))

    def endAssemblyFinish(self):
        self.ctx._chk(lib().l3b_mf_end_assembly_finish(self._h))

    @property
    def device_diag(self):
        return lib().l3b_mf_device_diag(self._h)

    @property
    def device_rhs(self):
        return lib().l3b_mf_device_rhs(self._h)

    def download(self):
        diag = np.zeros(self.n_dofs)
        rhs = np.zeros((self.n_rhs, self.n_dofs))
        self.ctx._chk(lib().l3b_mf_download(self._h, _p(diag), _p(rhs)))
        return diag, rhs.T.copy()

    def apply(self, x, y=None, alpha=1.0, beta=0.0):
        """y = alpha A x + beta y with host buffers (x, y: n_dofs x n_cols)."""
        x = np.asarray(x, dtype=np.float64)
        nc = 1 if x.ndim == 1 else x.shape[1]
        x = x.reshape(self.n_dofs, nc)
        xf = np.ascontiguousarray(x.T)
        yf = np.zeros_like(xf) if y is None else np.array(np.asarray(y, dtype=np.float64).reshape(self.n_dofs, nc).T, order="C", copy=True)
        self.ctx._chk(lib().l3b_mf_apply(self._h, _p(xf), _p(yf), nc, alpha, beta))
        return yf.T.copy()

    def apply_raw(self, xf, yf, n_cols=1, alpha=1.0, beta=0.0):
        """host buffers already in the C ABI layout (column-major, contiguous): no numpy copies in the timed region"""
        self.ctx._chk(lib().l3b_mf_apply(self._h, xf.ctypes.data, yf.ctypes.data, n_cols, alpha, beta))

    def set_host_apply(self, mode=1, n_chunks=0, block_nodes=65536):
        """how `apply` / `apply_raw` move their host vectors (l3b_mf_set_host_apply): 0 serial, 1 streamed when it pays, 2 streamed
        whenever legal"""
        self.ctx._chk(lib().l3b_mf_set_host_apply(self._h, mode, n_chunks, block_nodes))

    def host_apply_info(self):
        info = np.zeros(4, dtype=np.int64)
        self.ctx._chk(lib().l3b_mf_host_apply_info(self._h, _p(info)))
        return {"streamed": bool(info[0]), "items": int(info[1]), "up_ranges": int(info[2]), "down_ranges": int(info[3])}

    def apply_device(self, x_ptr, y_ptr, n_cols=1, alpha=1.0, beta=0.0, energy_ptr=None):
        if energy_ptr is None:
            self.ctx._chk(lib().l3b_mf_apply_device(self._h, x_ptr, y_ptr, n_cols, alpha, beta))
        else:
            if n_cols != 1:
                raise ValueError("the energy is collected for single-column applies")
            self.ctx._chk(lib().l3b_mf_apply_energy_device(self._h, x_ptr, y_ptr, alpha, beta, energy_ptr))

    def apply_phase_device(self, x_ptr, y_ptr, phases, elem_begin=0, elem_end=None, n_cols=1, alpha=1.0, beta=0.0, energy_ptr=None):
        """One phase of the apply (APPLY_INIT | APPLY_ELEMENTS | APPLY_FINISH) on device pointers over [owned | ghost] dofs; energy_ptr:
        device scalar that collects x^T A x of column 0 (elements of this call, owned Dirichlet dofs in the FINISH phase)."""
        end = self.mesh.n_elems if elem_end is None else elem_end
        self.ctx._chk(lib().l3b_mf_apply_phase_device(self._h, x_ptr, y_ptr, n_cols, alpha, beta, phases, elem_begin, end, energy_ptr))

    def solve(self, tol=1e-6, max_iters=10000, x0=None):
        """x0: initial guess over the local dofs (default zero, the benchmark's first solve)"""
        x = np.zeros(max(self.n_dofs, 1)) if x0 is None else np.array(x0, dtype=np.float64, copy=True).reshape(-1)
        at, it = C.c_double(), C.c_int()
        self.ctx._chk(lib().l3b_mf_solve_cg(self._h, tol, max_iters, _p(x), C.byref(at), C.byref(it)))
        return x[:self.n_dofs], at.value, it.value

    def solve_gmres(self, tol=1e-6, restart_length=250, max_restarts=39, max_iters=10000, x0=None):
        x = np.zeros(max(self.n_dofs, 1)) if x0 is None else np.array(x0, dtype=np.float64, copy=True).reshape(-1)
        at, it = C.c_double(), C.c_int()
        self.ctx._chk(lib().l3b_mf_solve_gmres(self._h, tol, restart_length, max_restarts, max_iters, _p(x), C.byref(at), C.byref(it)))
        return x[:self.n_dofs], at.value, it.value

    @property
    def kernel_launches(self):
        return lib().l3b_mf_kernel_launches(self._h)
