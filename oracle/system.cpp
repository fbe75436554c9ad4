// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp).
// Single-rank restatement of the reference's system level: DOF numbering, sparsity graph, global assembly + scatter,
// algebraic Dirichlet BCs, the matrix-free system (diag/rhs init and operator apply) and Jacobi-preconditioned CG.
// Threading follows the reference's model (parallel loop over elements, thread-local element scratch, relaxed atomic
// adds into shared global arrays) so the same code doubles as the timed CPU baseline.
#include "l3ster_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <exception>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>

namespace orc
{
namespace
{
template < typename F >
void parallelFor(std::size_t n, int n_threads, F&& f)
{
    n_threads = std::max(1, std::min< int >(n_threads, static_cast< int >(std::max< std::size_t >(n, 1))));
    if (n_threads == 1)
    {
        for (std::size_t i = 0; i < n; ++i)
            f(i, 0);
        return;
    }
    std::atomic< std::size_t > next{0};
    std::exception_ptr         err;
    std::mutex                 err_mtx;
    std::vector< std::thread > pool;
    constexpr std::size_t      chunk = 4;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t] {
            try
            {
                for (;;)
                {
                    const std::size_t b = next.fetch_add(chunk, std::memory_order_relaxed);
                    if (b >= n)
                        break;
                    for (std::size_t i = b; i < std::min(n, b + chunk); ++i)
                        f(i, t);
                }
            }
            catch (...)
            {
                std::lock_guard lock{err_mtx};
                err = std::current_exception();
            }
        });
    for (auto& th : pool)
        th.join();
    if (err)
        std::rethrow_exception(err);
}

inline void atomicAdd(val_t& dst, val_t v)
{
    std::atomic_ref< val_t >{dst}.fetch_add(v, std::memory_order_relaxed);
}

// post/FieldAccess.hpp:40-47 — node_vals(row, col) = data[node + field*n_nodes_total]
void gatherFields(const val_t* fields, std::size_t n_nodes_total, const n_id_t* el_nodes, int nn, int NF, val_t* node_vals)
{
    for (int a = 0; a < nn; ++a)
        for (int f = 0; f < NF; ++f)
            node_vals[static_cast< std::size_t >(a) * NF + f] = fields[el_nodes[a] + static_cast< std::size_t >(f) * n_nodes_total];
}

int quadOrder(const AssemblyOptions& o, int EO)
{
    return 2 * o.order(EO);
}
} // namespace

// post/Integral.hpp:55-121, post/NormL2.hpp:10-60
std::vector< val_t > computeIntegral(const Mesh& mesh, const Kernel& kernel, const AssemblyOptions& opts_in, val_t time, const val_t* fields,
                                     const std::vector< int >& boundary_ids, const std::vector< int >& field_inds, bool norm_l2)
{
    AssemblyOptions opts = opts_in;
    if (norm_l2)
    {
        opts.value_order *= 2;
        opts.derivative_order *= 2;
    }
    const int qo = opts.order(mesh.order);
    const int nn = mesh.nodesPerElem(), NF = kernel.params.n_fields, nv = kernel.params.n_equations * kernel.params.n_rhs;
    std::vector< val_t > total(nv, 0.), part(nv), node_vals(static_cast< std::size_t >(nn) * std::max(NF, 1));
    const auto one = [&](std::size_t e, int side, const RefBasisAtQuad& rbq) {
        const n_id_t* el_nodes = &mesh.elem_nodes[e * nn];
        for (int a = 0; a < nn; ++a)
            for (int f = 0; f < NF; ++f)
                node_vals[static_cast< std::size_t >(a) * NF + f] =
                    fields[el_nodes[a] + static_cast< std::size_t >(field_inds.empty() ? f : field_inds[f]) * mesh.n_nodes];
        evalElementIntegral(kernel, mesh.et, mesh.order, &mesh.elem_verts[e * (1u << nativeDim(mesh.et)) * 3], node_vals.data(), rbq, time,
                            side, norm_l2, part.data());
        for (int i = 0; i < nv; ++i)
            total[i] += part[i];
    };
    if (not kernel.is_boundary)
    {
        const auto rbq = makeRefBasisAtDomainQuad(mesh.et, mesh.order, qo);
        for (std::size_t e = 0; e < mesh.n_elems; ++e)
            one(e, -1, rbq);
    }
    else
    {
        std::vector< RefBasisAtQuad > tables;
        for (int s = 0; s < numSides(mesh.et); ++s)
            tables.push_back(makeRefBasisAtBoundaryQuad(mesh.et, mesh.order, qo, s));
        for (const auto& b : mesh.boundary)
            if (std::find(boundary_ids.begin(), boundary_ids.end(), b.domain_id) != boundary_ids.end())
                one(b.parent, b.side, tables[b.side]);
    }
    if (norm_l2)
        for (auto& v : total)
            v = std::sqrt(v);
    return total;
}

// algsys/SparsityGraph.hpp:25-81 (count with over-allocation → fill with duplicates → sort + unique per row) and
// :254-278 (rows inserted with sorted local column ids). Single rank, all U unknowns active on every node:
// global dof of (node, u) = node*U + u   (dofs/NodeToDofMap.hpp:249-264), local id == global id.
CrsGraph makeSparsityGraph(const Mesh& mesh, int U)
{
    // The reference over-allocates L entries per (row, element) pair, fills with duplicates and sort-uniques each row.
    // The result is, per row, the sorted union of the DOFs of all elements containing the row's node; it is built here
    // node by node (all U rows of a node share one column set) to keep the oracle's memory bounded.
    const int                                 nn = mesh.nodesPerElem();
    std::vector< std::vector< std::size_t > > node2el(mesh.n_nodes);
    for (std::size_t e = 0; e < mesh.n_elems; ++e)
        for (int a = 0; a < nn; ++a)
            node2el[mesh.elem_nodes[e * nn + a]].push_back(e);
    CrsGraph g;
    g.row_ptr.assign(mesh.n_nodes * U + 1, 0);
    std::vector< n_id_t > nbr_nodes;
    for (std::size_t node = 0; node < mesh.n_nodes; ++node)
    {
        std::size_t cnt = 0;
        nbr_nodes.clear();
        for (auto e : node2el[node])
            nbr_nodes.insert(nbr_nodes.end(), &mesh.elem_nodes[e * nn], &mesh.elem_nodes[e * nn] + nn);
        std::sort(nbr_nodes.begin(), nbr_nodes.end());
        nbr_nodes.erase(std::unique(nbr_nodes.begin(), nbr_nodes.end()), nbr_nodes.end());
        cnt = nbr_nodes.size() * U;
        for (int u = 0; u < U; ++u)
            g.row_ptr[node * U + u + 1] = static_cast< std::int64_t >(cnt);
    }
    for (std::size_t r = 0; r < mesh.n_nodes * U; ++r)
        g.row_ptr[r + 1] += g.row_ptr[r];
    g.col_ind.resize(g.row_ptr.back());
    for (std::size_t node = 0; node < mesh.n_nodes; ++node)
    {
        nbr_nodes.clear();
        for (auto e : node2el[node])
            nbr_nodes.insert(nbr_nodes.end(), &mesh.elem_nodes[e * nn], &mesh.elem_nodes[e * nn] + nn);
        std::sort(nbr_nodes.begin(), nbr_nodes.end());
        nbr_nodes.erase(std::unique(nbr_nodes.begin(), nbr_nodes.end()), nbr_nodes.end());
        for (int u = 0; u < U; ++u)
        {
            auto* dst = &g.col_ind[g.row_ptr[node * U + u]];
            for (auto nb : nbr_nodes)
                for (int v = 0; v < U; ++v)
                    *dst++ = static_cast< local_dof_t >(nb * U + v);
        }
    }
    return g;
}

namespace
{
// Tpetra::CrsMatrix::sumIntoLocalValues as used by algsys/ScatterLocalSystem.hpp:38-46: per entry, locate the column in
// the (sorted) row and add atomically
void sumIntoLocalValues(AssembledSystem& sys, local_dof_t row, const local_dof_t* cols, const val_t* vals, int n)
{
    const auto  rb   = sys.graph.row_ptr[row], re = sys.graph.row_ptr[row + 1];
    const auto* cbeg = sys.graph.col_ind.data() + rb;
    const auto* cend = sys.graph.col_ind.data() + re;
    for (int i = 0; i < n; ++i)
    {
        const auto it = std::lower_bound(cbeg, cend, cols[i]);
        if (it == cend or *it != cols[i])
            throw std::logic_error{"column not present in the sparsity graph"};
        atomicAdd(sys.values[rb + (it - cbeg)], vals[i]);
    }
}
} // namespace

// algsys/AssembleGlobalSystem.hpp:13-96 → assembleLocalSystem → StaticCondensationManager<None>::condenseSystem
// (StaticCondensationManager.hpp:110-121) → getDofsFromNodes (dofs/DofsFromNodes.hpp:72-87) → scatterLocalSystem
// dof_inds: kernel unknown u -> system dof (getDofsFromNodes(nodes, map, dof_inds), dofs/DofsFromNodes.hpp:72-87); field_inds: kernel
// field f -> column of the field storage (post/FieldAccess.hpp:10-53). Empty = identity.
void assembleGlobalSystem(AssembledSystem& sys, const Kernel& kernel, const AssemblyOptions& opts, val_t time, const val_t* fields,
                          int n_threads, const std::vector< int >& boundary_ids, const std::vector< int >& dof_inds,
                          const std::vector< int >& field_inds)
{
    const Mesh& mesh = *sys.mesh;
    const int   nn = mesh.nodesPerElem(), kU = kernel.params.n_unknowns, NF = kernel.params.n_fields, NRHS = kernel.params.n_rhs;
    if ((dof_inds.empty() ? kU != sys.U : static_cast< int >(dof_inds.size()) != kU) or NRHS != sys.n_rhs)
        throw std::invalid_argument{"oracle assembled system: kernel unknowns / rhs must match the system"};
    const int         L      = nn * kU;
    const std::size_t n_dofs = mesh.n_nodes * sys.U;
    const int         qo     = quadOrder(opts, mesh.order);

    struct Work
    {
        std::size_t elem;
        int         side;
    };
    std::vector< Work >           work;
    std::vector< RefBasisAtQuad > tables; // index 0: domain; 1 + side: boundary
    if (not kernel.is_boundary)
    {
        tables.push_back(makeRefBasisAtDomainQuad(mesh.et, mesh.order, qo));
        for (std::size_t e = 0; e < mesh.n_elems; ++e)
            work.push_back({e, -1});
    }
    else
    {
        tables.emplace_back();
        for (int s = 0; s < numSides(mesh.et); ++s)
            tables.push_back(makeRefBasisAtBoundaryQuad(mesh.et, mesh.order, qo, s));
        for (const auto& b : mesh.boundary)
            if (std::find(boundary_ids.begin(), boundary_ids.end(), b.domain_id) != boundary_ids.end())
            {
                if (b.side < 0)
                    throw std::logic_error{"matchBoundaries() was not called"};
                work.push_back({b.parent, b.side});
            }
    }
    struct Scratch
    {
        std::vector< val_t >       K, F, node_vals;
        std::vector< local_dof_t > dofs;
    };
    std::vector< Scratch > scratch(std::max(1, n_threads));
    for (auto& s : scratch)
    {
        s.K.resize(static_cast< std::size_t >(L) * L);
        s.F.resize(static_cast< std::size_t >(L) * NRHS);
        s.node_vals.resize(static_cast< std::size_t >(nn) * std::max(NF, 1));
        s.dofs.resize(L);
    }
    parallelFor(work.size(), n_threads, [&](std::size_t wi, int tid) {
        auto&         s        = scratch[tid];
        const auto    [e, side] = work[wi];
        const n_id_t* el_nodes = &mesh.elem_nodes[e * nn];
        if (NF > 0)
        {
            if (field_inds.empty())
                gatherFields(fields, mesh.n_nodes, el_nodes, nn, NF, s.node_vals.data());
            else
                for (int a = 0; a < nn; ++a)
                    for (int f = 0; f < NF; ++f)
                        s.node_vals[static_cast< std::size_t >(a) * NF + f] = fields[el_nodes[a] + static_cast< std::size_t >(field_inds[f]) * mesh.n_nodes];
        }
        const auto& rbq = tables[side < 0 ? 0 : 1 + side];
        assembleLocalSystem(kernel, mesh.et, mesh.order, &mesh.elem_verts[e * (1u << nativeDim(mesh.et)) * 3], s.node_vals.data(), rbq,
                            time, side, s.K.data(), s.F.data());
        for (int a = 0; a < nn; ++a)
            for (int u = 0; u < kU; ++u)
                s.dofs[a * kU + u] = static_cast< local_dof_t >(el_nodes[a] * sys.U + (dof_inds.empty() ? u : dof_inds[u]));
        for (int r = 0; r < L; ++r)
        {
            sumIntoLocalValues(sys, s.dofs[r], s.dofs.data(), &s.K[static_cast< std::size_t >(r) * L], L);
            for (int c = 0; c < NRHS; ++c)
                atomicAdd(sys.rhs[s.dofs[r] + static_cast< std::size_t >(c) * n_dofs], s.F[r + static_cast< std::size_t >(c) * L]);
        }
    });
}

// bcs/DirichletBC.hpp:82-150: BC rows become identity rows with rhs = prescribed value; in all other rows the BC
// columns are moved to a side matrix, zeroed in the system, and rhs -= side_matrix * bc_vals
void applyDirichletAlgebraic(AssembledSystem& sys, const std::vector< local_dof_t >& dofs, const std::vector< val_t >& vals)
{
    const std::size_t   n_dofs = sys.graph.row_ptr.size() - 1;
    std::vector< char > is_bc(n_dofs, 0);
    std::vector< val_t > bc_vals(n_dofs * sys.n_rhs, 0.);
    for (std::size_t i = 0; i < dofs.size(); ++i)
    {
        is_bc[dofs[i]] = 1;
        for (int r = 0; r < sys.n_rhs; ++r)
            bc_vals[dofs[i] + static_cast< std::size_t >(r) * n_dofs] = vals[i + static_cast< std::size_t >(r) * dofs.size()];
    }
    for (std::size_t row = 0; row < n_dofs; ++row)
    {
        const auto rb = sys.graph.row_ptr[row], re = sys.graph.row_ptr[row + 1];
        if (is_bc[row])
        {
            for (auto k = rb; k < re; ++k)
                sys.values[k] = sys.graph.col_ind[k] == static_cast< local_dof_t >(row) ? 1. : 0.;
            for (int r = 0; r < sys.n_rhs; ++r)
                sys.rhs[row + static_cast< std::size_t >(r) * n_dofs] = bc_vals[row + static_cast< std::size_t >(r) * n_dofs];
        }
        else
            for (auto k = rb; k < re; ++k)
            {
                const auto col = sys.graph.col_ind[k];
                if (is_bc[col])
                {
                    for (int r = 0; r < sys.n_rhs; ++r)
                        sys.rhs[row + static_cast< std::size_t >(r) * n_dofs] -=
                            sys.values[k] * bc_vals[col + static_cast< std::size_t >(r) * n_dofs];
                    sys.values[k] = 0.;
                }
            }
    }
}

namespace
{
struct MfWork
{
    std::size_t elem;
    int         side;
};
std::vector< MfWork > mfWorkList(const MatrixFreeSystem& sys, const MatrixFreeSystem::Entry& en)
{
    const Mesh&           mesh = *sys.mesh;
    std::vector< MfWork > work;
    if (not en.kernel->is_boundary)
        for (std::size_t e = 0; e < mesh.n_elems; ++e)
            work.push_back({e, -1});
    else
        for (const auto& b : mesh.boundary)
            if (std::find(en.boundary_ids.begin(), en.boundary_ids.end(), b.domain_id) != en.boundary_ids.end())
            {
                if (b.side < 0)
                    throw std::logic_error{"matchBoundaries() was not called"};
                work.push_back({b.parent, b.side});
            }
    return work;
}
std::vector< RefBasisAtQuad > mfTables(const Mesh& mesh, const MatrixFreeSystem::Entry& en)
{
    const int                     qo = quadOrder(en.opts, mesh.order);
    std::vector< RefBasisAtQuad > tables;
    if (not en.kernel->is_boundary)
        tables.push_back(makeRefBasisAtDomainQuad(mesh.et, mesh.order, qo));
    else
    {
        tables.emplace_back();
        for (int s = 0; s < numSides(mesh.et); ++s)
            tables.push_back(makeRefBasisAtBoundaryQuad(mesh.et, mesh.order, qo, s));
    }
    return tables;
}
} // namespace

// algsys/MatrixFreeSystem.hpp:887-941 with makeInitKernel (:585-631) and detail::scatterInit (:377-390)
void mfComputeDiagAndRhs(MatrixFreeSystem& sys, int n_threads)
{
    const Mesh&       mesh   = *sys.mesh;
    const std::size_t n_dofs = mesh.n_nodes * sys.U;
    const int         nn = mesh.nodesPerElem(), nvert = 1 << nativeDim(mesh.et);
    sys.diag.assign(n_dofs, 0.);
    sys.rhs.assign(n_dofs * sys.n_rhs, 0.);
    for (const auto& en : sys.kernels)
    {
        const Kernel& kernel = *en.kernel;
        const int     kU = kernel.params.n_unknowns, NF = kernel.params.n_fields, NRHS = kernel.params.n_rhs;
        if (kU != sys.U or NRHS != sys.n_rhs)
            throw std::invalid_argument{"oracle matrix-free system: kernel unknowns / rhs must match the system"};
        const int  L      = nn * kU;
        const auto work   = mfWorkList(sys, en);
        const auto tables = mfTables(mesh, en);
        parallelFor(work.size(), n_threads, [&](std::size_t wi, int) {
            const auto    [e, side] = work[wi];
            const n_id_t* el_nodes = &mesh.elem_nodes[e * nn];
            std::vector< val_t >       node_vals(static_cast< std::size_t >(nn) * std::max(NF, 1)), diag(L), rhs(static_cast< std::size_t >(L) * NRHS);
            std::vector< local_dof_t > dofs(L);
            if (NF > 0)
                gatherFields(en.fields, mesh.n_nodes, el_nodes, nn, NF, node_vals.data());
            for (int a = 0; a < nn; ++a)
                for (int u = 0; u < kU; ++u)
                    dofs[a * kU + u] = static_cast< local_dof_t >(el_nodes[a] * sys.U + u);
            // getDirichletDofInds / gatherDirichletVals (:313-375)
            std::vector< int > dir_inds;
            for (int i = 0; i < L; ++i)
                if (not sys.is_dirichlet.empty() and sys.is_dirichlet[dofs[i]])
                    dir_inds.push_back(i);
            std::vector< val_t > dir_vals(dir_inds.size() * NRHS);
            for (std::size_t i = 0; i < dir_inds.size(); ++i)
                for (int r = 0; r < NRHS; ++r)
                    dir_vals[i + r * dir_inds.size()] = sys.dirichlet_vals[dofs[dir_inds[i]] + static_cast< std::size_t >(r) * n_dofs];
            precomputeOperatorDiagonalAndRhs(kernel, mesh.et, mesh.order, &mesh.elem_verts[e * nvert * 3], node_vals.data(),
                                             tables[side < 0 ? 0 : 1 + side], en.time, side, static_cast< int >(dir_inds.size()),
                                             dir_inds.data(), dir_vals.data(), diag.data(), rhs.data());
            for (int i = 0; i < L; ++i)
            {
                atomicAdd(sys.diag[dofs[i]], diag[i]);
                for (int r = 0; r < NRHS; ++r)
                    atomicAdd(sys.rhs[dofs[i] + static_cast< std::size_t >(r) * n_dofs], rhs[i + static_cast< std::size_t >(r) * L]);
            }
        });
    }
    // handle_dirichlet_dof (:911-915)
    if (not sys.is_dirichlet.empty())
        for (std::size_t d = 0; d < n_dofs; ++d)
            if (sys.is_dirichlet[d])
            {
                sys.diag[d] = 1.;
                for (int r = 0; r < sys.n_rhs; ++r)
                    sys.rhs[d + static_cast< std::size_t >(r) * n_dofs] = sys.dirichlet_vals[d + static_cast< std::size_t >(r) * n_dofs];
            }
}

// algsys/MatrixFreeSystem.hpp:1019-1140 with makeEvalKernel (:633-712), gather/scatter (:393-537)
void mfApply(const MatrixFreeSystem& sys, const val_t* x, val_t* y, int n_cols, val_t alpha, val_t beta, int n_threads)
{
    const Mesh&       mesh   = *sys.mesh;
    const std::size_t n_dofs = mesh.n_nodes * sys.U;
    const int         nn = mesh.nodesPerElem(), nvert = 1 << nativeDim(mesh.et);
    const bool        has_bc = not sys.is_dirichlet.empty();
    // beta == 0 ? y.putScalar(0) : y.scale(beta)   (:1038)
    for (std::size_t i = 0; i < n_dofs * n_cols; ++i)
        y[i] = beta == 0. ? 0. : beta * y[i];
    for (const auto& en : sys.kernels)
    {
        const Kernel& kernel = *en.kernel;
        const int     kU = kernel.params.n_unknowns, NF = kernel.params.n_fields;
        const int     L      = nn * kU;
        const auto    work   = mfWorkList(sys, en);
        const bool    use_sf = not kernel.is_boundary and en.opts.eval_strategy != 1 and (mesh.et == Quad or mesh.et == Hex);
        const auto    tables = use_sf ? std::vector< RefBasisAtQuad >{} : mfTables(mesh, en);
        const int     F_tot  = kU * n_cols + NF;
        parallelFor(work.size(), n_threads, [&](std::size_t wi, int) {
            const auto    [e, side] = work[wi];
            const n_id_t* el_nodes = &mesh.elem_nodes[e * nn];
            const val_t*  verts    = &mesh.elem_verts[e * nvert * 3];
            if (use_sf)
            {
                // gatherSumFact (:421-467) + FieldAccess::fill (SumFactorization.hpp:901-909)
                std::vector< val_t > X(static_cast< std::size_t >(nn) * F_tot), Y(static_cast< std::size_t >(nn) * kU * n_cols);
                for (int node = 0; node < nn; ++node)
                    for (int dof = 0; dof < kU; ++dof)
                    {
                        const std::size_t id = el_nodes[node] * sys.U + dof;
                        for (int rhs = 0; rhs < n_cols; ++rhs)
                            X[static_cast< std::size_t >(rhs) * (kU * nn) + dof * nn + node] =
                                (has_bc and sys.is_dirichlet[id]) ? 0. : x[id + static_cast< std::size_t >(rhs) * n_dofs];
                    }
                for (int f = 0; f < NF; ++f)
                    for (int node = 0; node < nn; ++node)
                        X[static_cast< std::size_t >(kU * n_cols + f) * nn + node] = en.fields[el_nodes[node] + static_cast< std::size_t >(f) * mesh.n_nodes];
                evalLocalOperatorSumFact(kernel, mesh.et, mesh.order, verts, en.opts, en.time, n_cols, X.data(), Y.data());
                // scatterSumFact (:494-537)
                for (int node = 0; node < nn; ++node)
                    for (int dof = 0; dof < kU; ++dof)
                    {
                        const std::size_t id = el_nodes[node] * sys.U + dof;
                        if (has_bc and sys.is_dirichlet[id])
                            continue;
                        for (int rhs = 0; rhs < n_cols; ++rhs)
                            atomicAdd(y[id + static_cast< std::size_t >(rhs) * n_dofs],
                                      Y[static_cast< std::size_t >(node) * (kU * n_cols) + rhs * kU + dof] * alpha);
                    }
            }
            else
            {
                std::vector< val_t > node_vals(static_cast< std::size_t >(nn) * std::max(NF, 1)), xl(static_cast< std::size_t >(L) * n_cols),
                    yl(static_cast< std::size_t >(L) * n_cols);
                if (NF > 0)
                    gatherFields(en.fields, mesh.n_nodes, el_nodes, nn, NF, node_vals.data());
                // gather (:393-418)
                for (int a = 0; a < nn; ++a)
                    for (int u = 0; u < kU; ++u)
                    {
                        const std::size_t id = el_nodes[a] * sys.U + u;
                        for (int c = 0; c < n_cols; ++c)
                            xl[a * kU + u + static_cast< std::size_t >(c) * L] =
                                (has_bc and sys.is_dirichlet[id]) ? 0. : x[id + static_cast< std::size_t >(c) * n_dofs];
                    }
                evaluateLocalOperator(kernel, mesh.et, mesh.order, verts, node_vals.data(), tables[side < 0 ? 0 : 1 + side], en.time, side,
                                      n_cols, xl.data(), yl.data());
                // scatter (:469-492)
                for (int a = 0; a < nn; ++a)
                    for (int u = 0; u < kU; ++u)
                    {
                        const std::size_t id = el_nodes[a] * sys.U + u;
                        if (has_bc and sys.is_dirichlet[id])
                            continue;
                        for (int c = 0; c < n_cols; ++c)
                            atomicAdd(y[id + static_cast< std::size_t >(c) * n_dofs], yl[a * kU + u + static_cast< std::size_t >(c) * L] * alpha);
                    }
            }
        });
    }
    // Dirichlet rows are identity rows (:1087-1103)
    if (has_bc)
        for (std::size_t d = 0; d < n_dofs; ++d)
            if (sys.is_dirichlet[d])
                for (int c = 0; c < n_cols; ++c)
                    y[d + static_cast< std::size_t >(c) * n_dofs] += x[d + static_cast< std::size_t >(c) * n_dofs] * alpha;
}

// Belos "Block CG" (block size 1) with a left preconditioner (solve/BelosSolvers.hpp:76-89), native Jacobi
// M^-1 = damping * sign(d) / max(|d|, threshold) with the defaults damping 1, threshold 0 (NativePreconditioners.hpp:86-100),
// x0 = 0, absolute 2-norm residual test (IterSolverOpts defaults, solve/SolverInterface.hpp:26-37). Belos itself is a
// third-party dependency absent from the reference tree; iteration counts are not asserted anywhere by the reference
// (tests/SolverTests.cpp:181-185) — parity on them is unpinned.
SolveResult cgJacobi(const std::function< void(const val_t*, val_t*) >& apply, const val_t* diag, const val_t* b, val_t* x, std::size_t n,
                     val_t tol, int max_iters)
{
    std::vector< val_t > r(b, b + n), z(n), p(n), Ap(n), minv(n);
    for (std::size_t i = 0; i < n; ++i)
    {
        const val_t sign = diag[i] < 0 ? -1. : 1.;
        minv[i]          = sign / std::max(std::fabs(diag[i]), 0.);
    }
    std::fill_n(x, n, 0.);
    const auto dot = [&](const std::vector< val_t >& a, const std::vector< val_t >& c) {
        long double s = 0.;
        for (std::size_t i = 0; i < n; ++i)
            s += static_cast< long double >(a[i]) * c[i];
        return static_cast< val_t >(s);
    };
    val_t rnorm = std::sqrt(dot(r, r));
    int   iters = 0;
    if (rnorm <= tol)
        return {rnorm, 0};
    for (std::size_t i = 0; i < n; ++i)
        p[i] = z[i] = minv[i] * r[i];
    val_t rz = dot(r, z);
    while (iters < max_iters)
    {
        apply(p.data(), Ap.data());
        const val_t a = rz / dot(p, Ap);
        for (std::size_t i = 0; i < n; ++i)
        {
            x[i] += a * p[i];
            r[i] -= a * Ap[i];
        }
        ++iters;
        rnorm = std::sqrt(dot(r, r));
        if (rnorm <= tol)
            break;
        for (std::size_t i = 0; i < n; ++i)
            z[i] = minv[i] * r[i];
        const val_t rz_new = dot(r, z);
        const val_t bt     = rz_new / rz;
        rz                 = rz_new;
        for (std::size_t i = 0; i < n; ++i)
            p[i] = z[i] + bt * p[i];
    }
    return {rnorm, iters};
}
} // namespace orc
