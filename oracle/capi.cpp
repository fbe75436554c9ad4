// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp). extern "C" surface for ctypes (tests/, bench.py's
// cpu_baseline and --impl reference legs, __graft_entry__.smoke()). Never linked into the product.
#include "l3ster_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>

using namespace orc;

namespace
{
thread_local std::string g_err;
template < typename F >
int guarded(F&& f)
{
    try
    {
        f();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_err = e.what();
        return 1;
    }
}
Kernel kernelWithRhs(const char* name, int n_rhs)
{
    Kernel k = getKernel(name);
    if (n_rhs > 0)
        k.params.n_rhs = n_rhs;
    return k;
}
AssemblyOptions mkOpts(int value_order, int derivative_order, int eval_strategy)
{
    AssemblyOptions o;
    o.value_order      = value_order;
    o.derivative_order = derivative_order;
    o.eval_strategy    = eval_strategy;
    return o;
}
struct MeshHandle
{
    Mesh mesh;
};
struct AsmHandle
{
    AssembledSystem sys;
};
struct MfHandle
{
    MatrixFreeSystem                         sys;
    std::vector< std::unique_ptr< Kernel > > kernels;
    std::vector< std::vector< val_t > >      fields;
};
} // namespace

extern "C"
{
const char* orc_last_error()
{
    return g_err.c_str();
}

// ---- tables
int orc_legendre(int n, double* out)
{
    return guarded([&] {
        const auto c = legendreCoefs(n);
        std::copy(c.begin(), c.end(), out);
    });
}
int orc_lobatto(int n_points, double* out)
{
    return guarded([&] {
        const auto& a = lobattoAbsc(n_points);
        std::copy(a.begin(), a.end(), out);
    });
}
int orc_gauss(int n_points, double* pts, double* wts)
{
    return guarded([&] {
        const auto& r = gaussLegendre(n_points);
        std::copy(r.points.begin(), r.points.end(), pts);
        std::copy(r.weights.begin(), r.weights.end(), wts);
    });
}
int orc_lagrange_interp(int n, const double* x, const double* y, double* coefs)
{
    return guarded([&] {
        const auto c = lagrangeInterp({x, x + n}, {y, y + n});
        std::copy(c.begin(), c.end(), coefs);
    });
}
int orc_quadrature_size(int et, int quad_order)
{
    return ipow(refQuadSize(quad_order), et);
}
int orc_quadrature(int et, int quad_order, double* pts, double* wts)
{
    return guarded([&] {
        const auto q = makeQuadrature(static_cast< ElementType >(et), quad_order);
        std::copy(q.points.begin(), q.points.end(), pts);
        std::copy(q.weights.begin(), q.weights.end(), wts);
    });
}
double orc_ref_basis_value(int et, int order, int I, const double* pt)
{
    return refBasisValue(static_cast< ElementType >(et), order, I, pt);
}
double orc_ref_basis_der(int et, int order, int I, int d, const double* pt)
{
    return refBasisDer(static_cast< ElementType >(et), order, I, d, pt);
}
// side < 0: domain table. Outputs may be null. Returns the number of quadrature points (or -1).
int orc_ref_basis_at_quad(int et, int order, int quad_order, int side, double* pts, double* wts, double* vals, double* ders)
{
    int n = -1;
    guarded([&] {
        const auto r = side < 0 ? makeRefBasisAtDomainQuad(static_cast< ElementType >(et), order, quad_order)
                                : makeRefBasisAtBoundaryQuad(static_cast< ElementType >(et), order, quad_order, side);
        if (pts)
            std::copy(r.quad.points.begin(), r.quad.points.end(), pts);
        if (wts)
            std::copy(r.quad.weights.begin(), r.quad.weights.end(), wts);
        if (vals)
            std::copy(r.values.begin(), r.values.end(), vals);
        if (ders)
            std::copy(r.derivatives.begin(), r.derivatives.end(), ders);
        n = r.quad.size;
    });
    return n;
}

// ---- mapping
int orc_jacobi_mat(int et, const double* verts, const double* pt, double* J)
{
    return guarded([&] { jacobiMat(static_cast< ElementType >(et), verts, pt, J); });
}
int orc_map_to_physical(int et, const double* verts, const double* pt, double* out3)
{
    return guarded([&] { mapToPhysicalSpace(static_cast< ElementType >(et), verts, pt, out3); });
}
int orc_boundary_normal(int et, int side, const double* J, double* n)
{
    return guarded([&] { boundaryNormal(static_cast< ElementType >(et), side, J, n); });
}
double orc_boundary_jacobian(int et, int side, const double* J)
{
    return boundaryIntegralJacobian(static_cast< ElementType >(et), side, J);
}
// phys ders (dim x n_bases row-major) = J^-1 * ref ders at a point
int orc_phys_basis_ders(int et, int order, const double* verts, const double* pt, double* out)
{
    return guarded([&] {
        const auto t   = static_cast< ElementType >(et);
        const int  dim = nativeDim(t), nb = numNodes(t, order);
        double     J[9], Ji[9];
        jacobiMat(t, verts, pt, J);
        inverse(dim, J, Ji);
        for (int r = 0; r < dim; ++r)
            for (int a = 0; a < nb; ++a)
            {
                double acc = 0.;
                for (int k = 0; k < dim; ++k)
                    acc += Ji[r * dim + k] * refBasisDer(t, order, a, k, pt);
                out[r * nb + a] = acc;
            }
    });
}

// ---- element level
int orc_kernel_params(const char* name, int* out5)
{
    return guarded([&] {
        const Kernel k = getKernel(name);
        out5[0]       = k.params.dimension;
        out5[1]       = k.params.n_equations;
        out5[2]       = k.params.n_unknowns;
        out5[3]       = k.params.n_fields;
        out5[4]       = k.params.n_rhs;
    });
}
int orc_assemble_local(const char* kernel, int n_rhs, int et, int order, const double* verts, const double* node_vals, int value_order,
                       int der_order, double time, int side, double* K, double* F)
{
    return guarded([&] {
        const auto k  = kernelWithRhs(kernel, n_rhs);
        const auto t  = static_cast< ElementType >(et);
        const int  qo = 2 * mkOpts(value_order, der_order, 0).order(order);
        const auto rbq = side < 0 ? makeRefBasisAtDomainQuad(t, order, qo) : makeRefBasisAtBoundaryQuad(t, order, qo, side);
        assembleLocalSystem(k, t, order, verts, node_vals, rbq, time, side, K, F);
    });
}
int orc_eval_local_operator(const char* kernel, int n_rhs, int et, int order, const double* verts, const double* node_vals, int value_order,
                            int der_order, double time, int side, int n_cols, const double* x, double* y)
{
    return guarded([&] {
        const auto k  = kernelWithRhs(kernel, n_rhs);
        const auto t  = static_cast< ElementType >(et);
        const int  qo = 2 * mkOpts(value_order, der_order, 0).order(order);
        const auto rbq = side < 0 ? makeRefBasisAtDomainQuad(t, order, qo) : makeRefBasisAtBoundaryQuad(t, order, qo, side);
        evaluateLocalOperator(k, t, order, verts, node_vals, rbq, time, side, n_cols, x, y);
    });
}
int orc_precompute_diag_rhs(const char* kernel, int n_rhs, int et, int order, const double* verts, const double* node_vals, int value_order,
                            int der_order, double time, int side, int n_dir, const int* dir_inds, const double* dir_vals, double* diag,
                            double* rhs)
{
    return guarded([&] {
        const auto k  = kernelWithRhs(kernel, n_rhs);
        const auto t  = static_cast< ElementType >(et);
        const int  qo = 2 * mkOpts(value_order, der_order, 0).order(order);
        const auto rbq = side < 0 ? makeRefBasisAtDomainQuad(t, order, qo) : makeRefBasisAtBoundaryQuad(t, order, qo, side);
        precomputeOperatorDiagonalAndRhs(k, t, order, verts, node_vals, rbq, time, side, n_dir, dir_inds, dir_vals, diag, rhs);
    });
}
// x: L x n_cols col-major in the element's (node*U + u) ordering; node_vals row-major n_nodes x n_fields. Wraps the
// reference layouts of gatherSumFact / scatterSumFact exactly like tests/LocalOperatorCommon.hpp:150-188 does.
int orc_eval_sumfact(const char* kernel, int n_rhs, int et, int order, const double* verts, const double* node_vals, int value_order,
                     int der_order, int eval_strategy, double time, int n_cols, const double* x, double* y)
{
    return guarded([&] {
        const auto k  = kernelWithRhs(kernel, n_rhs);
        const auto t  = static_cast< ElementType >(et);
        const int  nn = numNodes(t, order), U = k.params.n_unknowns, NF = k.params.n_fields, L = nn * U;
        std::vector< double > X(static_cast< std::size_t >(nn) * (U * n_cols + NF)), Y(static_cast< std::size_t >(nn) * U * n_cols);
        for (int n = 0; n < nn; ++n)
            for (int r = 0; r < n_cols; ++r)
                for (int u = 0; u < U; ++u)
                    X[static_cast< std::size_t >(r * U + u) * nn + n] = x[n * U + u + static_cast< std::size_t >(r) * L];
        for (int f = 0; f < NF; ++f)
            for (int n = 0; n < nn; ++n)
                X[static_cast< std::size_t >(U * n_cols + f) * nn + n] = node_vals[static_cast< std::size_t >(n) * NF + f];
        evalLocalOperatorSumFact(k, t, order, verts, mkOpts(value_order, der_order, eval_strategy), time, n_cols, X.data(), Y.data());
        for (int n = 0; n < nn; ++n)
            for (int r = 0; r < n_cols; ++r)
                for (int u = 0; u < U; ++u)
                    y[n * U + u + static_cast< std::size_t >(r) * L] = Y[static_cast< std::size_t >(n) * (U * n_cols) + r * U + u];
    });
}
// single sweep, for the odd-even == standard test (tests/SumFactorizationTests.cpp:58-129)
// kind: 0 back interp, 1 back der, 2 forward interp assign, 3 forward der accumulate
int orc_sumfact_sweep(int basis_order, int quad_order, int kind, int odd_even, int cols, const double* in, double* out)
{
    return guarded([&] {
        const auto t = makeSumFactTables(basis_order, quad_order, odd_even != 0);
        switch (kind)
        {
        case 0:
            sumFactSweep(in, out, t.nb, t.nq, cols, t.interp.data(), false, false, odd_even != 0);
            break;
        case 1:
            sumFactSweep(in, out, t.nb, t.nq, cols, t.der.data(), true, false, odd_even != 0);
            break;
        case 2:
            sumFactSweep(in, out, t.nq, t.nb, cols, t.interp_t.data(), false, false, odd_even != 0);
            break;
        case 3:
            sumFactSweep(in, out, t.nq, t.nb, cols, t.der_t.data(), true, true, odd_even != 0);
            break;
        default:
            throw std::invalid_argument{"sweep kind"};
        }
    });
}
int orc_sumfact_tables(int basis_order, int quad_order, double* interp, double* der)
{
    return guarded([&] {
        const auto t = makeSumFactTables(basis_order, quad_order, false);
        std::copy(t.interp.begin(), t.interp.end(), interp);
        std::copy(t.der.begin(), t.der.end(), der);
    });
}

// ---- mesh
void* orc_mesh_cube(int nx, const double* dx, int ny, const double* dy, int nz, const double* dz, int order)
{
    void* ret = nullptr;
    guarded([&] {
        auto m = makeCubeMesh({dx, dx + nx}, {dy, dy + ny}, {dz, dz + nz});
        if (order > 1)
            m = convertMeshToOrder(m, order);
        matchBoundaries(m);
        ret = new MeshHandle{std::move(m)};
    });
    return ret;
}
void* orc_mesh_square(int nx, const double* dx, int ny, const double* dy, int order)
{
    void* ret = nullptr;
    guarded([&] {
        auto m = makeSquareMesh({dx, dx + nx}, {dy, dy + ny});
        if (order > 1)
            m = convertMeshToOrder(m, order);
        matchBoundaries(m);
        ret = new MeshHandle{std::move(m)};
    });
    return ret;
}
// single element mesh from explicit vertices (tests/LocalOperatorCommon.hpp:17-59 fixtures), nodes numbered
// boundary-first as the fixtures do
void* orc_mesh_single(int et, int order, const double* verts)
{
    void* ret = nullptr;
    guarded([&] {
        Mesh       m;
        const auto t = static_cast< ElementType >(et);
        m.et         = t;
        m.order      = order;
        m.n_elems    = 1;
        m.n_nodes    = numNodes(t, order);
        m.elem_nodes.resize(m.n_nodes);
        n_id_t n = 0;
        for (int i : boundaryNodeInds(t, order))
            m.elem_nodes[i] = n++;
        for (int i : internalNodeInds(t, order))
            m.elem_nodes[i] = n++;
        m.elem_verts.assign(verts, verts + (1 << et) * 3);
        m.elem_ids = {0};
        ret        = new MeshHandle{std::move(m)};
    });
    return ret;
}
// a mesh of order `order` given directly by its element node lists (lexicographic local order, numbering as mesh::convertMeshToOrder
// would produce it) and element vertices — no conversion and no boundary elements: for parity tests at sizes where the geometric node
// matching of the conversion (restated faithfully, and quadratic-ish in practice) would take minutes; the numbering itself is pinned at
// small sizes by comparing the two generators
void* orc_mesh_from_nodes(int et, int order, long long n_nodes, long long n_elems, const unsigned long long* elem_nodes, const double* elem_verts)
{
    void* ret = nullptr;
    guarded([&] {
        Mesh       m;
        const auto t = static_cast< ElementType >(et);
        m.et         = t;
        m.order      = order;
        m.n_nodes    = static_cast< std::size_t >(n_nodes);
        m.n_elems    = static_cast< std::size_t >(n_elems);
        const auto npe = static_cast< std::size_t >(numNodes(t, order));
        m.elem_nodes.assign(elem_nodes, elem_nodes + m.n_elems * npe);
        m.elem_verts.assign(elem_verts, elem_verts + m.n_elems * (1u << et) * 3);
        m.elem_ids.resize(m.n_elems);
        for (std::size_t e = 0; e < m.n_elems; ++e)
            m.elem_ids[e] = static_cast< n_id_t >(e);
        ret = new MeshHandle{std::move(m)};
    });
    return ret;
}
// an order-1 mesh given by arrays (what mesh::readMesh produces: node coordinates, volume elements with vertex lists in lexicographic
// order, boundary elements with their domain ids; ids: volume elements 0.., boundary elements as given), converted to `order` and
// boundary-matched — the general path of mesh/ConvertMeshToOrder.hpp, for unstructured parity tests
void* orc_mesh_from_arrays(int et, long long n_nodes, const double* coords, long long n_elems, const long long* elems, long long n_bnd,
                           const long long* bnd_elems, const int* bnd_domains, const long long* bnd_ids, int order)
{
    void* ret = nullptr;
    guarded([&] {
        Mesh       m;
        const auto t  = static_cast< ElementType >(et);
        const int  nv = 1 << et, nbv = 1 << (et - 1);
        m.et          = t;
        m.order       = 1;
        m.n_nodes     = static_cast< std::size_t >(n_nodes);
        m.n_elems     = static_cast< std::size_t >(n_elems);
        for (long long e = 0; e < n_elems; ++e)
        {
            m.elem_ids.push_back(static_cast< n_id_t >(e));
            for (int v = 0; v < nv; ++v)
            {
                const auto n = elems[e * nv + v];
                m.elem_nodes.push_back(static_cast< n_id_t >(n));
                m.elem_verts.insert(m.elem_verts.end(), coords + 3 * n, coords + 3 * n + 3);
            }
        }
        for (long long b = 0; b < n_bnd; ++b)
        {
            Mesh::BoundaryElem be;
            be.domain_id = bnd_domains[b];
            be.id        = static_cast< n_id_t >(bnd_ids[b]);
            for (int v = 0; v < nbv; ++v)
            {
                const auto n = bnd_elems[b * nbv + v];
                be.nodes.push_back(static_cast< n_id_t >(n));
                be.verts.insert(be.verts.end(), coords + 3 * n, coords + 3 * n + 3);
            }
            m.boundary.push_back(std::move(be));
        }
        if (order > 1)
            m = convertMeshToOrder(m, order);
        matchBoundaries(m);
        ret = new MeshHandle{std::move(m)};
    });
    return ret;
}
void orc_mesh_free(void* h)
{
    delete static_cast< MeshHandle* >(h);
}
// info: [et, order, n_nodes, n_elems, nodes_per_elem, n_boundary_elems]
void orc_mesh_info(void* h, long long* info)
{
    const auto& m = static_cast< MeshHandle* >(h)->mesh;
    info[0]       = m.et;
    info[1]       = m.order;
    info[2]       = static_cast< long long >(m.n_nodes);
    info[3]       = static_cast< long long >(m.n_elems);
    info[4]       = m.nodesPerElem();
    info[5]       = static_cast< long long >(m.boundary.size());
}
void orc_mesh_get(void* h, unsigned long long* elem_nodes, double* elem_verts)
{
    const auto& m = static_cast< MeshHandle* >(h)->mesh;
    if (elem_nodes)
        std::copy(m.elem_nodes.begin(), m.elem_nodes.end(), elem_nodes);
    if (elem_verts)
        std::copy(m.elem_verts.begin(), m.elem_verts.end(), elem_verts);
}
// per boundary element: domain id, parent element, side; nodes (n_b x nodes_per_boundary_elem)
void orc_mesh_boundary(void* h, int* domain, long long* parent, int* side, unsigned long long* nodes)
{
    const auto& m = static_cast< MeshHandle* >(h)->mesh;
    std::size_t k = 0;
    for (std::size_t i = 0; i < m.boundary.size(); ++i)
    {
        domain[i] = m.boundary[i].domain_id;
        parent[i] = static_cast< long long >(m.boundary[i].parent);
        side[i]   = m.boundary[i].side;
        if (nodes)
            for (auto n : m.boundary[i].nodes)
                nodes[k++] = n;
    }
}
// overwrite the element vertices (tests distort the structured meshes); boundary matching is node-based and unaffected
void orc_mesh_set_verts(void* h, const double* elem_verts)
{
    auto& m = static_cast< MeshHandle* >(h)->mesh;
    std::copy(elem_verts, elem_verts + m.elem_verts.size(), m.elem_verts.begin());
}
int orc_side_node_inds(int et, int order, int side, int* out)
{
    const auto v = sideNodeInds(static_cast< ElementType >(et), order, side);
    std::copy(v.begin(), v.end(), out);
    return static_cast< int >(v.size());
}
int orc_boundary_node_inds(int et, int order, int* out)
{
    const auto& v = boundaryNodeInds(static_cast< ElementType >(et), order);
    if (out)
        std::copy(v.begin(), v.end(), out);
    return static_cast< int >(v.size());
}

// ---- assembled system
void* orc_asm_create(void* mesh_h, int U, int n_rhs)
{
    void* ret = nullptr;
    guarded([&] {
        auto h       = std::make_unique< AsmHandle >();
        h->sys.mesh  = &static_cast< MeshHandle* >(mesh_h)->mesh;
        h->sys.U     = U;
        h->sys.n_rhs = n_rhs;
        h->sys.graph = makeSparsityGraph(*h->sys.mesh, U);
        h->sys.values.assign(h->sys.graph.col_ind.size(), 0.);
        h->sys.rhs.assign(h->sys.mesh->n_nodes * U * n_rhs, 0.);
        ret = h.release();
    });
    return ret;
}
void orc_asm_free(void* h)
{
    delete static_cast< AsmHandle* >(h);
}
long long orc_asm_nnz(void* h)
{
    return static_cast< long long >(static_cast< AsmHandle* >(h)->sys.graph.col_ind.size());
}
void orc_asm_graph(void* h, long long* row_ptr, int* col_ind)
{
    const auto& g = static_cast< AsmHandle* >(h)->sys.graph;
    std::copy(g.row_ptr.begin(), g.row_ptr.end(), row_ptr);
    std::copy(g.col_ind.begin(), g.col_ind.end(), col_ind);
}
void orc_asm_zero(void* h)
{
    auto& s = static_cast< AsmHandle* >(h)->sys;
    std::fill(s.values.begin(), s.values.end(), 0.);
    std::fill(s.rhs.begin(), s.rhs.end(), 0.);
}
// returns elapsed seconds in *secs (assembly + scatter only — the reference's assembleGlobalSystem region)
int orc_asm_assemble(void* h, const char* kernel, int value_order, int der_order, double time, const double* fields, int n_threads,
                     int n_bnd_ids, const int* bnd_ids, double* secs)
{
    return guarded([&] {
        auto&      s  = static_cast< AsmHandle* >(h)->sys;
        const auto k  = kernelWithRhs(kernel, s.n_rhs);
        const auto t0 = std::chrono::steady_clock::now();
        assembleGlobalSystem(s, k, mkOpts(value_order, der_order, 0), time, fields, n_threads, {bnd_ids, bnd_ids + n_bnd_ids});
        if (secs)
            *secs = std::chrono::duration< double >(std::chrono::steady_clock::now() - t0).count();
    });
}
// the same with the kernel's unknowns mapped to system dofs and its fields to storage columns (either may be null = identity)
int orc_asm_assemble_ex(void* h, const char* kernel, int value_order, int der_order, double time, const double* fields, int n_threads,
                        int n_bnd_ids, const int* bnd_ids, const int* dof_inds, const int* field_inds)
{
    return guarded([&] {
        auto&              s = static_cast< AsmHandle* >(h)->sys;
        const auto         k = kernelWithRhs(kernel, s.n_rhs);
        std::vector< int > di, fi;
        if (dof_inds)
            di.assign(dof_inds, dof_inds + k.params.n_unknowns);
        if (field_inds)
            fi.assign(field_inds, field_inds + k.params.n_fields);
        assembleGlobalSystem(s, k, mkOpts(value_order, der_order, 0), time, fields, n_threads, {bnd_ids, bnd_ids + n_bnd_ids}, di, fi);
    });
}
int orc_asm_dirichlet(void* h, int n, const int* dofs, const double* vals)
{
    return guarded([&] {
        auto& s = static_cast< AsmHandle* >(h)->sys;
        applyDirichletAlgebraic(s, {dofs, dofs + n}, {vals, vals + static_cast< std::size_t >(n) * s.n_rhs});
    });
}
void orc_asm_get(void* h, double* values, double* rhs)
{
    const auto& s = static_cast< AsmHandle* >(h)->sys;
    if (values)
        std::copy(s.values.begin(), s.values.end(), values);
    if (rhs)
        std::copy(s.rhs.begin(), s.rhs.end(), rhs);
}

// computeIntegral / computeNormL2 of a residual kernel on the mesh (one rank); out: n_equations * n_rhs doubles
int orc_compute_integral(void* mesh_h, const char* kernel, int value_order, int der_order, double time, const double* fields,
                         const int* field_inds, int n_bnd_ids, const int* bnd_ids, int norm_l2, double* out, int* n_out)
{
    return guarded([&] {
        const auto&        mesh = static_cast< MeshHandle* >(mesh_h)->mesh;
        const auto&        k    = getKernel(kernel);
        std::vector< int > fi;
        if (field_inds)
            fi.assign(field_inds, field_inds + k.params.n_fields);
        const auto r = computeIntegral(mesh, k, mkOpts(value_order, der_order, 0), time, fields, {bnd_ids, bnd_ids + n_bnd_ids}, fi, norm_l2 != 0);
        std::copy(r.begin(), r.end(), out);
        *n_out = static_cast< int >(r.size());
    });
}

int orc_values_at_nodes(void* mesh_h, const char* kernel, double time, const double* fields, const int* field_inds, int n_bnd_ids,
                        const int* bnd_ids, int dpn, const int* dof_inds, double* values)
{
    return guarded([&] {
        const auto&        mesh = static_cast< MeshHandle* >(mesh_h)->mesh;
        const auto&        k    = getKernel(kernel);
        std::vector< int > fi, di(k.params.n_equations);
        if (field_inds)
            fi.assign(field_inds, field_inds + k.params.n_fields);
        for (int eq = 0; eq < k.params.n_equations; ++eq)
            di[eq] = dof_inds ? dof_inds[eq] : eq;
        computeValuesAtNodes(mesh, k, time, fields, fi, {bnd_ids, bnd_ids + n_bnd_ids}, dpn, di, values);
    });
}

// ---- matrix-free system
void* orc_mf_create(void* mesh_h, int U, int n_rhs, const unsigned char* is_dirichlet, const double* dirichlet_vals)
{
    void* ret = nullptr;
    guarded([&] {
        auto h             = std::make_unique< MfHandle >();
        h->sys.mesh        = &static_cast< MeshHandle* >(mesh_h)->mesh;
        h->sys.U           = U;
        h->sys.n_rhs       = n_rhs;
        const std::size_t n = h->sys.mesh->n_nodes * U;
        if (is_dirichlet)
        {
            h->sys.is_dirichlet.assign(is_dirichlet, is_dirichlet + n);
            h->sys.dirichlet_vals.assign(n * n_rhs, 0.);
            if (dirichlet_vals)
                std::copy(dirichlet_vals, dirichlet_vals + n * n_rhs, h->sys.dirichlet_vals.begin());
        }
        ret = h.release();
    });
    return ret;
}
void orc_mf_free(void* h)
{
    delete static_cast< MfHandle* >(h);
}
int orc_mf_add_kernel(void* hv, const char* kernel, int value_order, int der_order, int eval_strategy, double time, const double* fields,
                      int n_bnd_ids, const int* bnd_ids)
{
    return guarded([&] {
        auto* h = static_cast< MfHandle* >(hv);
        h->kernels.push_back(std::make_unique< Kernel >(kernelWithRhs(kernel, h->sys.n_rhs)));
        MatrixFreeSystem::Entry en{};
        en.kernel = h->kernels.back().get();
        en.opts   = mkOpts(value_order, der_order, eval_strategy);
        en.time   = time;
        en.boundary_ids.assign(bnd_ids, bnd_ids + n_bnd_ids);
        if (en.kernel->params.n_fields > 0)
        {
            h->fields.emplace_back(fields, fields + h->sys.mesh->n_nodes * en.kernel->params.n_fields);
            en.fields = h->fields.back().data();
        }
        h->sys.kernels.push_back(std::move(en));
    });
}
int orc_mf_init(void* hv, int n_threads, double* diag, double* rhs)
{
    return guarded([&] {
        auto& s = static_cast< MfHandle* >(hv)->sys;
        mfComputeDiagAndRhs(s, n_threads);
        if (diag)
            std::copy(s.diag.begin(), s.diag.end(), diag);
        if (rhs)
            std::copy(s.rhs.begin(), s.rhs.end(), rhs);
    });
}
int orc_mf_apply(void* hv, const double* x, double* y, int n_cols, double alpha, double beta, int n_threads, int repeats, double* secs)
{
    return guarded([&] {
        const auto& s  = static_cast< MfHandle* >(hv)->sys;
        const auto  t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < std::max(1, repeats); ++i)
            mfApply(s, x, y, n_cols, alpha, beta, n_threads);
        if (secs)
            *secs = std::chrono::duration< double >(std::chrono::steady_clock::now() - t0).count();
    });
}
// CG + native Jacobi on the matrix-free operator; b = system rhs (column 0), x0 = 0
int orc_mf_cg(void* hv, double tol, int max_iters, int n_threads, double* x, double* achieved_tol, int* iters)
{
    return guarded([&] {
        auto&             s = static_cast< MfHandle* >(hv)->sys;
        const std::size_t n = s.mesh->n_nodes * s.U;
        const auto        r = cgJacobi([&](const val_t* in, val_t* out) { mfApply(s, in, out, 1, 1., 0., n_threads); }, s.diag.data(),
                                s.rhs.data(), x, n, tol, max_iters);
        *achieved_tol       = r.tol;
        *iters              = r.iters;
    });
}
// CG + Jacobi on an explicit CRS matrix (assembled path)
int orc_crs_cg(long long n, const long long* row_ptr, const int* col_ind, const double* values, const double* b, double tol, int max_iters,
               double* x, double* achieved_tol, int* iters)
{
    return guarded([&] {
        std::vector< double > diag(n, 0.);
        for (long long r = 0; r < n; ++r)
            for (auto k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
                if (col_ind[k] == r)
                    diag[r] = values[k];
        const auto res = cgJacobi(
            [&](const val_t* in, val_t* out) {
                for (long long r = 0; r < n; ++r)
                {
                    double acc = 0.;
                    for (auto k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
                        acc += values[k] * in[col_ind[k]];
                    out[r] = acc;
                }
            },
            diag.data(), b, x, static_cast< std::size_t >(n), tol, max_iters);
        *achieved_tol = res.tol;
        *iters        = res.iters;
    });
}
} // extern "C"
