"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/liboracle.so`` (the plain C++ restatement of the reference's hot path, see
``oracle/l3ster_oracle.hpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package ``l3ster_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LINE, QUAD, HEX = 1, 2, 3

_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_dp = C.POINTER(C.c_double)


_NATIVE_BUILT = False


def build(native: bool = False) -> str:
    """Compile the oracle. ``native=True`` builds a -march=native copy (for CPU-baseline timing on the current host)."""
    global _NATIVE_BUILT
    out = "liboracle_native.so" if native else "liboracle.so"
    if native:
        objdir = os.path.join(_DIR, "_native")
        if _NATIVE_BUILT:  # once per process: the copy on disk may come from another host (another ISA), so the first call always compiles
            return os.path.join(objdir, out)
        os.makedirs(objdir, exist_ok=True)
        srcs = ["math", "element", "sumfact", "kernels", "mesh", "system", "capi"]
        # 512-bit vectors where the host has them: GCC prefers 256-bit ones on most AVX-512 cores, which costs the SYRK micro-kernel a third
        cmd = ["g++", "-std=c++20", "-O3", "-march=native", "-mprefer-vector-width=512", "-fopenmp-simd", "-fPIC", "-pthread", "-shared", "-o",
               os.path.join(objdir, out)]
        cmd += [os.path.join(_DIR, s + ".cpp") for s in srcs]
        subprocess.run(cmd, check=True, cwd=_DIR)
        _NATIVE_BUILT = True
        return os.path.join(objdir, out)
    subprocess.run(["make", "-s", "-j8"], check=True, cwd=_DIR)
    return os.path.join(_DIR, out)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, native: bool = False):
        path = os.path.join(_DIR, "_native", "liboracle_native.so") if native else os.path.join(_DIR, "liboracle.so")
        if not os.path.exists(path):
            path = build(native)
        self.lib = lib = C.CDLL(path)
        lib.orc_last_error.restype = C.c_char_p
        lib.orc_ref_basis_value.restype = C.c_double
        lib.orc_ref_basis_der.restype = C.c_double
        lib.orc_boundary_jacobian.restype = C.c_double
        for name in ("orc_mesh_cube", "orc_mesh_square", "orc_mesh_single", "orc_mesh_from_arrays", "orc_mesh_from_nodes", "orc_asm_create", "orc_mf_create"):
            getattr(lib, name).restype = C.c_void_p
        lib.orc_asm_nnz.restype = C.c_longlong

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())

    # ---- tables
    def legendre(self, n):
        out = np.zeros(n + 1)
        self._chk(self.lib.orc_legendre(n, _ptr(out)))
        return out

    def lobatto(self, n):
        out = np.zeros(n)
        self._chk(self.lib.orc_lobatto(n, _ptr(out)))
        return out

    def gauss(self, n):
        p, w = np.zeros(n), np.zeros(n)
        self._chk(self.lib.orc_gauss(n, _ptr(p), _ptr(w)))
        return p, w

    def lagrange_interp(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        out = np.zeros(len(x))
        self._chk(self.lib.orc_lagrange_interp(len(x), _ptr(x), _ptr(y), _ptr(out)))
        return out

    def quadrature(self, et, qo):
        n = self.lib.orc_quadrature_size(et, qo)
        p, w = np.zeros((n, et)), np.zeros(n)
        self._chk(self.lib.orc_quadrature(et, qo, _ptr(p), _ptr(w)))
        return p, w

    def ref_basis_value(self, et, order, i, pt):
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        return self.lib.orc_ref_basis_value(et, order, i, _ptr(pt))

    def ref_basis_der(self, et, order, i, d, pt):
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        return self.lib.orc_ref_basis_der(et, order, i, d, _ptr(pt))

    def ref_basis_at_quad(self, et, order, qo, side=-1):
        nb = (order + 1) ** et
        nq1 = qo // 2 + 1
        q = nq1**et if side < 0 else max(1, nq1 ** (et - 1))
        pts, wts = np.zeros((q, et)), np.zeros(q)
        vals, ders = np.zeros((q, nb)), np.zeros((q, et, nb))
        n = self.lib.orc_ref_basis_at_quad(et, order, qo, side, _ptr(pts), _ptr(wts), _ptr(vals), _ptr(ders))
        if n != q:
            raise RuntimeError(self.lib.orc_last_error().decode() or "quadrature size mismatch")
        return pts, wts, vals, ders

    # ---- mapping
    def jacobi_mat(self, et, verts, pt):
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        J = np.zeros((et, et))
        self._chk(self.lib.orc_jacobi_mat(et, _ptr(verts), _ptr(pt), _ptr(J)))
        return J

    def map_to_physical(self, et, verts, pt):
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        out = np.zeros(3)
        self._chk(self.lib.orc_map_to_physical(et, _ptr(verts), _ptr(pt), _ptr(out)))
        return out

    def boundary_normal(self, et, side, J):
        J = np.ascontiguousarray(J, dtype=np.float64)
        n = np.zeros(et)
        self._chk(self.lib.orc_boundary_normal(et, side, _ptr(J), _ptr(n)))
        return n

    def boundary_jacobian(self, et, side, J):
        J = np.ascontiguousarray(J, dtype=np.float64)
        return self.lib.orc_boundary_jacobian(et, side, _ptr(J))

    def phys_basis_ders(self, et, order, verts, pt):
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        out = np.zeros((et, (order + 1) ** et))
        self._chk(self.lib.orc_phys_basis_ders(et, order, _ptr(verts), _ptr(pt), _ptr(out)))
        return out

    # ---- element level
    def kernel_params(self, name):
        out = np.zeros(5, dtype=np.int32)
        self._chk(self.lib.orc_kernel_params(name.encode(), _ptr(out)))
        return dict(zip(("dimension", "n_equations", "n_unknowns", "n_fields", "n_rhs"), map(int, out)))

    def _nv(self, kernel, et, order, node_vals):
        nf = self.kernel_params(kernel)["n_fields"]
        nn = (order + 1) ** et
        if nf == 0:
            return np.zeros(1)
        nv = np.ascontiguousarray(node_vals, dtype=np.float64)
        assert nv.shape == (nn, nf)
        return nv

    def assemble_local(self, kernel, et, order, verts, node_vals=None, n_rhs=1, value_order=1, der_order=0, time=0.0, side=-1):
        kp = self.kernel_params(kernel)
        L = (order + 1) ** et * kp["n_unknowns"]
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        nv = self._nv(kernel, et, order, node_vals)
        K = np.zeros((L, L))
        F = np.zeros((n_rhs, L))  # col-major L x n_rhs
        self._chk(self.lib.orc_assemble_local(kernel.encode(), n_rhs, et, order, _ptr(verts), _ptr(nv), value_order, der_order,
                                              C.c_double(time), side, _ptr(K), _ptr(F)))
        return K, F.T.copy()

    def eval_local_operator(self, kernel, et, order, verts, x, node_vals=None, value_order=1, der_order=0, time=0.0, side=-1):
        x = np.asarray(x, dtype=np.float64)
        L, nc = x.shape
        xf = np.ascontiguousarray(x.T)
        yf = np.zeros_like(xf)
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        nv = self._nv(kernel, et, order, node_vals)
        self._chk(self.lib.orc_eval_local_operator(kernel.encode(), nc, et, order, _ptr(verts), _ptr(nv), value_order, der_order,
                                                   C.c_double(time), side, nc, _ptr(xf), _ptr(yf)))
        return yf.T.copy()

    def precompute_diag_rhs(self, kernel, et, order, verts, node_vals=None, n_rhs=1, value_order=1, der_order=0, time=0.0, side=-1,
                            dir_inds=None, dir_vals=None):
        kp = self.kernel_params(kernel)
        L = (order + 1) ** et * kp["n_unknowns"]
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        nv = self._nv(kernel, et, order, node_vals)
        nd = 0 if dir_inds is None else len(dir_inds)
        di = np.ascontiguousarray(dir_inds if nd else [0], dtype=np.int32)
        dv = np.ascontiguousarray(np.asarray(dir_vals, dtype=np.float64).T) if nd else np.zeros(1)
        diag = np.zeros(L)
        rhs = np.zeros((n_rhs, L))
        self._chk(self.lib.orc_precompute_diag_rhs(kernel.encode(), n_rhs, et, order, _ptr(verts), _ptr(nv), value_order, der_order,
                                                   C.c_double(time), side, nd, _ptr(di), _ptr(dv), _ptr(diag), _ptr(rhs)))
        return diag, rhs.T.copy()

    def eval_sumfact(self, kernel, et, order, verts, x, node_vals=None, value_order=1, der_order=0, eval_strategy=0, time=0.0):
        x = np.asarray(x, dtype=np.float64)
        L, nc = x.shape
        xf = np.ascontiguousarray(x.T)
        yf = np.zeros_like(xf)
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        nv = self._nv(kernel, et, order, node_vals)
        self._chk(self.lib.orc_eval_sumfact(kernel.encode(), nc, et, order, _ptr(verts), _ptr(nv), value_order, der_order,
                                            eval_strategy, C.c_double(time), nc, _ptr(xf), _ptr(yf)))
        return yf.T.copy()

    def sumfact_sweep(self, basis_order, quad_order, kind, odd_even, x, y0=None):
        """x: (n_in, cols) matrix (column = one line); returns (cols, n_out)."""
        x = np.asarray(x, dtype=np.float64)
        n_in, cols = x.shape
        nb, nq = basis_order + 1, quad_order // 2 + 1
        n_out = nq if kind < 2 else nb
        xin = np.ascontiguousarray(x.T)  # col-major n_in x cols
        out = np.zeros((n_out, cols)) if y0 is None else np.ascontiguousarray(np.asarray(y0, dtype=np.float64).T)
        self._chk(self.lib.orc_sumfact_sweep(basis_order, quad_order, kind, int(odd_even), cols, _ptr(xin), _ptr(out)))
        return out.T.copy()

    def sumfact_tables(self, basis_order, quad_order):
        nb, nq = basis_order + 1, quad_order // 2 + 1
        a, b = np.zeros((nb, nq)), np.zeros((nb, nq))
        self._chk(self.lib.orc_sumfact_tables(basis_order, quad_order, _ptr(a), _ptr(b)))
        return a, b

    # ---- mesh
    def mesh_cube(self, dx, dy=None, dz=None, order=1):
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        dy = dx if dy is None else np.ascontiguousarray(dy, dtype=np.float64)
        dz = dx if dz is None else np.ascontiguousarray(dz, dtype=np.float64)
        h = self.lib.orc_mesh_cube(len(dx), _ptr(dx), len(dy), _ptr(dy), len(dz), _ptr(dz), order)
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return OracleMesh(self, h)

    def mesh_square(self, dx, dy=None, order=1):
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        dy = dx if dy is None else np.ascontiguousarray(dy, dtype=np.float64)
        h = self.lib.orc_mesh_square(len(dx), _ptr(dx), len(dy), _ptr(dy), order)
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return OracleMesh(self, h)

    def mesh_from_arrays(self, et, coords, elems, bnd_elems, bnd_domains, bnd_ids, order):
        """order-1 mesh by arrays -> convertMeshToOrder(order) + matchBoundaries (the general, geometric-matching path)"""
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        elems = np.ascontiguousarray(elems, dtype=np.int64)
        be = np.ascontiguousarray(bnd_elems, dtype=np.int64)
        bd = np.ascontiguousarray(bnd_domains, dtype=np.int32)
        bi = np.ascontiguousarray(bnd_ids, dtype=np.int64)
        h = self.lib.orc_mesh_from_arrays(et, len(coords), _ptr(coords), len(elems), _ptr(elems), len(be), _ptr(be), _ptr(bd), _ptr(bi), order)
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return OracleMesh(self, h)

    def mesh_from_nodes(self, et, order, n_nodes, elem_nodes, elem_verts):
        """mesh of order `order` straight from element node lists + vertices (no conversion, no boundary elements)"""
        en = np.ascontiguousarray(elem_nodes, dtype=np.uint64)
        ev = np.ascontiguousarray(elem_verts, dtype=np.float64)
        h = self.lib.orc_mesh_from_nodes(et, order, C.c_longlong(n_nodes), C.c_longlong(en.shape[0]), _ptr(en), _ptr(ev))
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return OracleMesh(self, h)

    def mesh_single(self, et, order, verts):
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        h = self.lib.orc_mesh_single(et, order, _ptr(verts))
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        return OracleMesh(self, h)

    def side_node_inds(self, et, order, side):
        out = np.zeros((order + 1) ** (et - 1), dtype=np.int32)
        n = self.lib.orc_side_node_inds(et, order, side, _ptr(out))
        return out[:n]

    def boundary_node_inds(self, et, order):
        n = self.lib.orc_boundary_node_inds(et, order, None)
        out = np.zeros(n, dtype=np.int32)
        self.lib.orc_boundary_node_inds(et, order, _ptr(out))
        return out

    def crs_cg(self, row_ptr, col_ind, values, b, tol=1e-6, max_iters=10000):
        n = len(row_ptr) - 1
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col_ind = np.ascontiguousarray(col_ind, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(n)
        at, it = C.c_double(), C.c_int()
        self._chk(self.lib.orc_crs_cg(C.c_longlong(n), _ptr(row_ptr), _ptr(col_ind), _ptr(values), _ptr(b), C.c_double(tol), max_iters,
                                      _ptr(x), C.byref(at), C.byref(it)))
        return x, at.value, it.value


class OracleMesh:
    def __init__(self, orc: Oracle, handle):
        self.orc, self.h = orc, C.c_void_p(handle)
        info = np.zeros(6, dtype=np.int64)
        orc.lib.orc_mesh_info(self.h, _ptr(info))
        self.et, self.order, self.n_nodes, self.n_elems, self.nodes_per_elem, self.n_boundary = map(int, info)
        self.elem_nodes = np.zeros((self.n_elems, self.nodes_per_elem), dtype=np.uint64)
        self.elem_verts = np.zeros((self.n_elems, 2**self.et, 3))
        orc.lib.orc_mesh_get(self.h, _ptr(self.elem_nodes), _ptr(self.elem_verts))
        nb = self.n_boundary
        self.bnd_domain = np.zeros(nb, dtype=np.int32)
        self.bnd_parent = np.zeros(nb, dtype=np.int64)
        self.bnd_side = np.zeros(nb, dtype=np.int32)
        npb = (self.order + 1) ** (self.et - 1)
        self.bnd_nodes = np.zeros((nb, npb), dtype=np.uint64)
        if nb:
            orc.lib.orc_mesh_boundary(self.h, _ptr(self.bnd_domain), _ptr(self.bnd_parent), _ptr(self.bnd_side), _ptr(self.bnd_nodes))

    def __del__(self):
        try:
            self.orc.lib.orc_mesh_free(self.h)
        except Exception:
            pass

    def set_verts(self, verts):
        verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(self.elem_verts.shape)
        self.orc.lib.orc_mesh_set_verts(self.h, _ptr(verts))
        self.elem_verts = verts.copy()

    def assembled_system(self, U, n_rhs=1):
        return OracleAssembled(self, U, n_rhs)

    def compute_integral(self, kernel, boundary_ids=(), fields=None, field_inds=None, value_order=1, der_order=0, time=0.0, norm_l2=False):
        """computeIntegral / computeNormL2 (post/Integral.hpp:102-121, post/NormL2.hpp:31-60) of a residual kernel"""
        f = None if fields is None else np.ascontiguousarray(fields, dtype=np.float64)
        b = np.ascontiguousarray(list(boundary_ids) or [0], dtype=np.int32)
        fi = None if field_inds is None else np.ascontiguousarray(field_inds, dtype=np.int32)
        out = np.zeros(64)
        n = C.c_int(0)
        self.orc._chk(self.orc.lib.orc_compute_integral(self.h, kernel.encode(), value_order, der_order, C.c_double(time), _ptr(f), _ptr(fi),
                                                        len(boundary_ids), _ptr(b), int(norm_l2), _ptr(out), C.byref(n)))
        return out[: n.value].copy()

    def matrix_free_system(self, U, n_rhs=1, is_dirichlet=None, dirichlet_vals=None):
        return OracleMatrixFree(self, U, n_rhs, is_dirichlet, dirichlet_vals)

    def values_at_nodes(self, kernel, values, dofs_per_node, boundary_ids=(), dof_inds=None, fields=None, field_inds=None, time=0.0):
        """computeValuesAtNodes (algsys/ComputeValuesAtNodes.hpp:316-506): `values` (n_nodes * dpn, n_rhs) is updated in place and returned"""
        v = np.array(np.asarray(values, dtype=np.float64).reshape(self.n_nodes * dofs_per_node, -1).T, order="C", copy=True)
        f = None if fields is None else np.ascontiguousarray(fields, dtype=np.float64)
        b = np.ascontiguousarray(list(boundary_ids) or [0], dtype=np.int32)
        fi = None if field_inds is None else np.ascontiguousarray(field_inds, dtype=np.int32)
        di = None if dof_inds is None else np.ascontiguousarray(dof_inds, dtype=np.int32)
        self.orc._chk(self.orc.lib.orc_values_at_nodes(self.h, kernel.encode(), C.c_double(time), _ptr(f), _ptr(fi), len(boundary_ids), _ptr(b),
                                                       dofs_per_node, _ptr(di), _ptr(v)))
        return v.T.copy()


class OracleAssembled:
    def __init__(self, mesh: OracleMesh, U, n_rhs):
        self.mesh, self.U, self.n_rhs = mesh, U, n_rhs
        self.lib = mesh.orc.lib
        h = self.lib.orc_asm_create(mesh.h, U, n_rhs)
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        self.h = C.c_void_p(h)
        self.n_dofs = mesh.n_nodes * U
        self.nnz = int(self.lib.orc_asm_nnz(self.h))
        self.row_ptr = np.zeros(self.n_dofs + 1, dtype=np.int64)
        self.col_ind = np.zeros(self.nnz, dtype=np.int32)
        self.lib.orc_asm_graph(self.h, _ptr(self.row_ptr), _ptr(self.col_ind))

    def __del__(self):
        try:
            self.lib.orc_asm_free(self.h)
        except Exception:
            pass

    def zero(self):
        self.lib.orc_asm_zero(self.h)

    def assemble(self, kernel, value_order=1, der_order=0, time=0.0, fields=None, n_threads=1, boundary_ids=()):
        f = None if fields is None else np.ascontiguousarray(fields, dtype=np.float64)
        b = np.ascontiguousarray(list(boundary_ids) or [0], dtype=np.int32)
        secs = C.c_double()
        self.mesh.orc._chk(self.lib.orc_asm_assemble(self.h, kernel.encode(), value_order, der_order, C.c_double(time), _ptr(f), n_threads,
                                                     len(boundary_ids), _ptr(b), C.byref(secs)))
        return secs.value

    def assemble_ex(self, kernel, value_order=1, der_order=0, time=0.0, fields=None, n_threads=1, boundary_ids=(), dof_inds=None, field_inds=None):
        """assemble with the kernel's unknowns mapped to system dofs / its fields to storage columns"""
        f = None if fields is None else np.ascontiguousarray(fields, dtype=np.float64)
        b = np.ascontiguousarray(list(boundary_ids) or [0], dtype=np.int32)
        di = None if dof_inds is None else np.ascontiguousarray(dof_inds, dtype=np.int32)
        fi = None if field_inds is None else np.ascontiguousarray(field_inds, dtype=np.int32)
        self.mesh.orc._chk(self.lib.orc_asm_assemble_ex(self.h, kernel.encode(), value_order, der_order, C.c_double(time), _ptr(f), n_threads,
                                                        len(boundary_ids), _ptr(b), _ptr(di), _ptr(fi)))

    def apply_dirichlet(self, dofs, vals):
        dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        vals = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(len(dofs), -1).T)
        self.mesh.orc._chk(self.lib.orc_asm_dirichlet(self.h, len(dofs), _ptr(dofs), _ptr(vals)))

    def get(self):
        values = np.zeros(self.nnz)
        rhs = np.zeros((self.n_rhs, self.n_dofs))
        self.lib.orc_asm_get(self.h, _ptr(values), _ptr(rhs))
        return values, rhs.T.copy()


class OracleMatrixFree:
    def __init__(self, mesh: OracleMesh, U, n_rhs, is_dirichlet, dirichlet_vals):
        self.mesh, self.U, self.n_rhs = mesh, U, n_rhs
        self.lib = mesh.orc.lib
        self.n_dofs = mesh.n_nodes * U
        isd = None if is_dirichlet is None else np.ascontiguousarray(is_dirichlet, dtype=np.uint8)
        dv = None
        if dirichlet_vals is not None:
            dv = np.ascontiguousarray(np.asarray(dirichlet_vals, dtype=np.float64).reshape(self.n_dofs, -1).T)
        h = self.lib.orc_mf_create(mesh.h, U, n_rhs, _ptr(isd), _ptr(dv))
        if not h:
            raise RuntimeError(self.lib.orc_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            self.lib.orc_mf_free(self.h)
        except Exception:
            pass

    def add_kernel(self, kernel, value_order=1, der_order=0, eval_strategy=0, time=0.0, fields=None, boundary_ids=()):
        f = None if fields is None else np.ascontiguousarray(fields, dtype=np.float64)
        b = np.ascontiguousarray(list(boundary_ids) or [0], dtype=np.int32)
        self.mesh.orc._chk(self.lib.orc_mf_add_kernel(self.h, kernel.encode(), value_order, der_order, eval_strategy, C.c_double(time),
                                                      _ptr(f), len(boundary_ids), _ptr(b)))

    def init(self, n_threads=1):
        diag = np.zeros(self.n_dofs)
        rhs = np.zeros((self.n_rhs, self.n_dofs))
        self.mesh.orc._chk(self.lib.orc_mf_init(self.h, n_threads, _ptr(diag), _ptr(rhs)))
        return diag, rhs.T.copy()

    def apply(self, x, y=None, alpha=1.0, beta=0.0, n_threads=1, repeats=1):
        x = np.asarray(x, dtype=np.float64).reshape(self.n_dofs, -1)
        nc = x.shape[1]
        xf = np.ascontiguousarray(x.T)
        yf = np.zeros_like(xf) if y is None else np.array(np.asarray(y, dtype=np.float64).reshape(self.n_dofs, -1).T, order="C", copy=True)
        secs = C.c_double()
        self.mesh.orc._chk(self.lib.orc_mf_apply(self.h, _ptr(xf), _ptr(yf), nc, C.c_double(alpha), C.c_double(beta), n_threads, repeats,
                                                 C.byref(secs)))
        self.last_secs = secs.value
        return yf.T.copy()

    def cg(self, tol=1e-6, max_iters=10000, n_threads=1):
        x = np.zeros(self.n_dofs)
        at, it = C.c_double(), C.c_int()
        self.mesh.orc._chk(self.lib.orc_mf_cg(self.h, C.c_double(tol), max_iters, n_threads, _ptr(x), C.byref(at), C.byref(it)))
        return x, at.value, it.value


def condense_element_boundary(K, F, elem_nodes, bnd_idx, int_idx, U):
    """CondensationPolicy::ElementBoundary restated on an assembled system (algsys/StaticCondensationManager.hpp:330-418): K dense
    (n_dofs x n_dofs, all nodes, dof = node * U + u), F (n_dofs x n_rhs). Per element: K_ii = rows/cols of its interior nodes (only
    this element contributes there, :391-403), K_pi = its boundary rows x interior cols (:405-416), and the primary block receives
    -K_pi K_ii^-1 K_ip, the primary rhs -K_pi K_ii^-1 f_i (:340-347; Eigen's dynamic-size inverse() is a partially pivoted LU, as
    numpy.linalg.inv). Returns (primary nodes, S, F_c, recover) with recover(x_c) -> nodal solution over all dofs (:420-535)."""
    import numpy as np

    K = np.array(K, dtype=np.float64)
    F = np.array(F, dtype=np.float64).reshape(K.shape[0], -1)
    elem_nodes = np.asarray(elem_nodes, dtype=np.int64)
    prim_nodes = np.unique(elem_nodes[:, bnd_idx])
    dofs = lambda nodes: (np.asarray(nodes)[:, None] * U + np.arange(U)[None, :]).ravel()
    P = dofs(prim_nodes)
    S, Fc = K.copy(), F.copy()
    saved = []
    for en in elem_nodes:
        p, i = dofs(en[bnd_idx]), dofs(en[int_idx])
        Kii_inv = np.linalg.inv(K[np.ix_(i, i)])
        Kpi = K[np.ix_(p, i)]
        S[np.ix_(p, p)] -= Kpi @ Kii_inv @ Kpi.T
        Fc[p] -= Kpi @ Kii_inv @ F[i]
        saved.append((p, i, Kii_inv, Kpi))

    def recover(x_c):
        x = np.zeros_like(F)
        x[P] = np.asarray(x_c).reshape(len(P), -1)
        for p, i, Kii_inv, Kpi in saved:
            x[i] = Kii_inv @ (F[i] - Kpi.T @ x[p])
        return x

    return prim_nodes, S[np.ix_(P, P)], Fc[P], recover
