// DOF maps for problem definitions with inactive (node, dof) pairs and several domains: the reference's ProblemDefinition ->
// NodeToGlobalDofMap -> sparsity graph chain (dofs/NodeToDofMap.hpp:144-163, 188-264, 335-357; algsys/SparsityGraph.hpp:25-81),
// pinned by its tests/MultiDomainTest.cpp and tests/SparsityGraphTest.cpp:99-146.
//
// Device storage stays PADDED: every node carries dofs_per_node slots (local dof = node * dofs_per_node + d) whether or not the problem
// activates them — the kernels index dofs without a table, and a (node, dof) pair no definition activates receives no contribution from
// any kernel, so its row and column stay empty. The systems close such pairs as identity rows with a zero right-hand side (the same
// mechanism as a homogeneous Dirichlet dof), which keeps the solvers well posed and the active part of the solution untouched.
// What the reference numbers and stores is the COMPACT form: only active pairs, numbered node-major, and a graph that holds, per
// definition, the clique of (element nodes) x (dofs of that definition). This header computes the compact numbering and graph and the
// translation between the two forms, so that everything handed across the C ABI (dof ids, CRS graph, values, vectors) is bit-exact with
// the reference's objects while the hot path keeps its table-free indexing.
#ifndef L3B_DOFMAP_HOST_HPP
#define L3B_DOFMAP_HOST_HPP

#include "mesh_host.hpp"

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace l3b::host
{
struct ProblemDefinition
{
    // definition i applies to the elements whose domain id is in ids[ptr[i] .. ptr[i+1]) — and to the element SIDES whose boundary id
    // is (a boundary is a domain of the lower-dimensional elements, mesh/MeshPartition.hpp) — and activates the dofs in mask[i]
    std::vector< int >      ptr{0};
    std::vector< int >      ids;
    std::vector< uint32_t > mask;
    int  size() const { return static_cast< int >(mask.size()); }
    bool has(int def, int id) const
    {
        for (int k = ptr[def]; k < ptr[def + 1]; ++k)
            if (ids[k] == id)
                return true;
        return false;
    }
};

struct DofMap
{
    int                      dpn = 0;
    long long                n_nodes = 0;
    std::vector< uint8_t >   active;  // [n_nodes][dpn]
    std::vector< long long > dof;     // [n_nodes][dpn]: compact dof id (base_dof + position in node-major order of the active pairs) or -1
    long long                base_dof = 0, n_dofs = 0;
    std::vector< long long > row_ptr; // compact CRS graph over the n_dofs rows, columns = compact local ids (dof - base_dof), ascending
    std::vector< int32_t >   col_ind;
};

// side -> local node indices (mesh/ElementTraits.hpp:72-98, 118-137)
inline std::vector< int > sideNodes(int dim, int order, int side)
{
    const int          n = order + 1;
    std::vector< int > out;
    if (dim == 2)
    {
        for (int k = 0; k < n; ++k)
            out.push_back(side == 0 ? k : side == 1 ? k + n * (n - 1) : side == 2 ? k * n : k * n + n - 1);
        return out;
    }
    const int nps = n * n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            switch (side)
            {
            case 0: out.push_back(i * n + j); break;
            case 1: out.push_back(i * n + j + nps * (n - 1)); break;
            case 2: out.push_back(i * nps + j); break;
            case 3: out.push_back(i * nps + j + n * (n - 1)); break;
            case 4: out.push_back(i * nps + j * n); break;
            default: out.push_back(i * nps + j * n + n - 1);
            }
    return out;
}

inline DofMap makeDofMap(int dim, int order, long long n_nodes, long long n_elems, const uint32_t* nodes, const int32_t* elem_domains,
                         const uint16_t* side_bnd, int dpn, const ProblemDefinition& def, long long base_dof)
{
    if (dpn < 1 or dpn > 32)
        throw std::invalid_argument{"dofs per node must be in [1, 32]"};
    int nn = 1;
    for (int d = 0; d < dim; ++d)
        nn *= order + 1;
    const int n_sides = 2 * dim;
    DofMap    m;
    m.dpn     = dpn;
    m.n_nodes = n_nodes;
    m.active.assign(static_cast< size_t >(n_nodes) * dpn, 0);
    std::vector< std::vector< int > > side_nodes(n_sides);
    for (int s = 0; s < n_sides; ++s)
        side_nodes[s] = sideNodes(dim, order, s);
    // every (entity, definition) pair that applies, as a node list + dof mask: visited twice (bitmap, graph)
    const auto forEachClique = [&](auto&& fn) {
        std::vector< uint32_t > cl;
        for (int i = 0; i < def.size(); ++i)
        {
            for (long long e = 0; e < n_elems; ++e)
            {
                if (def.has(i, elem_domains ? elem_domains[e] : 0))
                    fn(nodes + e * nn, nn, def.mask[i]);
                if (side_bnd)
                    for (int s = 0; s < n_sides; ++s)
                    {
                        const auto id = side_bnd[e * n_sides + s];
                        if (id == no_boundary or not def.has(i, id))
                            continue;
                        cl.clear();
                        for (int a : side_nodes[s])
                            cl.push_back(nodes[e * nn + a]);
                        fn(cl.data(), static_cast< int >(cl.size()), def.mask[i]);
                    }
            }
        }
    };
    // makeLocalDofBmp (dofs/NodeToDofMap.hpp:188-214)
    forEachClique([&](const uint32_t* cn, int cnt, uint32_t mask) {
        for (int a = 0; a < cnt; ++a)
            for (int d = 0; d < dpn; ++d)
                if (mask >> d & 1u)
                    m.active[static_cast< size_t >(cn[a]) * dpn + d] = 1;
    });
    // computeOwnedDofs (:248-264): node-major, dof-minor, active pairs only
    m.base_dof = base_dof;
    m.dof.assign(m.active.size(), -1);
    long long next = base_dof;
    for (size_t i = 0; i < m.active.size(); ++i)
        if (m.active[i])
            m.dof[i] = next++;
    m.n_dofs = next - base_dof;
    // computeLocalGraph (algsys/SparsityGraph.hpp:25-81): count with duplicates, fill, sort + unique per row
    std::vector< long long > cnt(m.n_dofs + 1, 0);
    forEachClique([&](const uint32_t* cn, int c, uint32_t mask) {
        const long long per = static_cast< long long >(c) * __builtin_popcount(mask & ((dpn == 32 ? 0u : (1u << dpn)) - 1u));
        for (int a = 0; a < c; ++a)
            for (int d = 0; d < dpn; ++d)
                if (mask >> d & 1u)
                    cnt[m.dof[static_cast< size_t >(cn[a]) * dpn + d] - base_dof + 1] += per;
    });
    for (long long r = 0; r < m.n_dofs; ++r)
        cnt[r + 1] += cnt[r];
    std::vector< int32_t >   raw(cnt.back());
    std::vector< long long > fill(cnt.begin(), cnt.end() - 1);
    std::vector< int32_t >   el;
    forEachClique([&](const uint32_t* cn, int c, uint32_t mask) {
        el.clear();
        for (int a = 0; a < c; ++a)
            for (int d = 0; d < dpn; ++d)
                if (mask >> d & 1u)
                    el.push_back(static_cast< int32_t >(m.dof[static_cast< size_t >(cn[a]) * dpn + d] - base_dof));
        for (int32_t r : el)
        {
            std::copy(el.begin(), el.end(), raw.begin() + fill[r]);
            fill[r] += static_cast< long long >(el.size());
        }
    });
    m.row_ptr.assign(m.n_dofs + 1, 0);
    for (long long r = 0; r < m.n_dofs; ++r)
    {
        auto* b = raw.data() + cnt[r];
        auto* e = raw.data() + cnt[r + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        m.col_ind.insert(m.col_ind.end(), b, e);
        m.row_ptr[r + 1] = static_cast< long long >(m.col_ind.size());
    }
    return m;
}
} // namespace l3b::host
#endif
