// Partition import: everything the multi-rank hot path needs from a mesh partition that was computed elsewhere (the reference calls
// METIS_PartMeshNodal, mesh/PartitionMesh.hpp:142-183 — a third-party heuristic that is not rebuilt here), i.e. the deterministic
// part of the reference's partitioner and of the set-up of its distributed objects, restated for a given (epart, npart):
//   * node assignment + repair of nodes METIS put where no element holds them: assignNodes / reassignDisjointNodes
//     (mesh/PartitionMesh.hpp:322-395);
//   * global renumbering so that every rank owns one contiguous id range, ghosts sorted by global id: renumberNodes (:411-440), and
//     the local numbering [owned | ghost sorted by GID] of dofs/NodeToDofMap.hpp:144-163;
//   * who shares what: comm::ImportExportContext (comm/ImportExport.hpp:29-72, 250-261) at node granularity;
//   * the row-complete sparsity graph of the owned rows, with the neighbours' contributions to shared rows and the column map they
//     induce: algsys/SparsityGraph.hpp:83-233, 254-278 (owned columns first, then every other column in ascending global id);
//   * the receive side of the shared-row export (Tpetra FECrsMatrix::endAssembly as called at algsys/AssembledSystem.hpp:384-389):
//     where each value a neighbour sends for one of my rows lands in my row.
// The reference exchanges rows and bitmaps between ranks with MPI at set-up; here every rank holds the global connectivity and the
// partition vectors (they generate or read the same mesh), so all of it is computed locally and identically on every rank — no
// set-up communication at all. Boundary elements follow their parent elements (assignBoundaryElements, :259-319) by construction:
// sides carry their boundary id in the element record.
//
// Two local numberings per rank ("views"):
//   plain     [owned | ghost]            ghosts = nodes of my elements owned elsewhere. What the matrix-free system uses.
//   extended  [owned | ghost + extra]    extra = columns that appear only in the neighbours' contributions to my rows (no element of
//                                        mine touches them): the column map of the reference's assembled matrix.
#ifndef L3B_PARTITION_HOST_HPP
#define L3B_PARTITION_HOST_HPP

#include "mesh_host.hpp"

#include <numeric>
#include <string>

namespace l3b::host
{
struct Partition
{
    int                      dim = 0, order = 1, nn = 0, n_sides = 0, n_parts = 1;
    long long                n_nodes = 0, n_elems = 0;
    std::vector< uint32_t >  nodes;  // global connectivity, input numbering
    std::vector< int32_t >   epart, npart;
    std::vector< long long > new_id; // input node id -> global id after renumberNodes
    std::vector< long long > dist;   // rank r owns the global ids [dist[r], dist[r + 1])
    // global node -> elements adjacency (input numbering)
    std::vector< long long > n2e_ptr;
    std::vector< uint32_t >  n2e;
    struct Rank
    {
        std::vector< long long > elems;                 // global element ids, border elements (touching a ghost node) first
        long long                n_border = 0;
        std::vector< long long > ghosts, ghosts_ext;    // global ids (new numbering), ascending
    };
    std::vector< Rank > ranks;

    int       ownerOf(long long gid) const { return static_cast< int >(std::upper_bound(dist.begin(), dist.end(), gid) - dist.begin()) - 1; }
    long long nOwned(int r) const { return dist[r + 1] - dist[r]; }
    const std::vector< long long >& ghostsOf(int r, bool extended) const { return extended ? ranks[r].ghosts_ext : ranks[r].ghosts; }
    long long nLocal(int r, bool extended) const { return nOwned(r) + static_cast< long long >(ghostsOf(r, extended).size()); }
    // local id of a global id in rank r's view, or -1
    long long localId(int r, bool extended, long long gid) const
    {
        if (gid >= dist[r] and gid < dist[r + 1])
            return gid - dist[r];
        const auto& g  = ghostsOf(r, extended);
        const auto  it = std::lower_bound(g.begin(), g.end(), gid);
        return it != g.end() and *it == gid ? nOwned(r) + (it - g.begin()) : -1;
    }
};

inline Partition makePartition(int dim, int order, long long n_nodes, long long n_elems, const uint32_t* nodes, int n_parts, const int32_t* epart,
                               const int32_t* npart_in)
{
    Partition p;
    p.dim     = dim;
    p.order   = order;
    p.nn      = 1;
    for (int d = 0; d < dim; ++d)
        p.nn *= order + 1;
    p.n_sides = 2 * dim;
    p.n_parts = n_parts;
    p.n_nodes = n_nodes;
    p.n_elems = n_elems;
    const int nn = p.nn;
    p.nodes.assign(nodes, nodes + n_elems * nn);
    p.epart.assign(epart, epart + n_elems);
    for (long long e = 0; e < n_elems; ++e)
        if (epart[e] < 0 or epart[e] >= n_parts)
            throw std::invalid_argument{"epart entry out of range"};
    // adjacency
    p.n2e_ptr.assign(n_nodes + 1, 0);
    for (long long i = 0; i < n_elems * nn; ++i)
    {
        if (nodes[i] >= n_nodes)
            throw std::invalid_argument{"element node id out of range"};
        ++p.n2e_ptr[nodes[i] + 1];
    }
    for (long long n = 0; n < n_nodes; ++n)
        p.n2e_ptr[n + 1] += p.n2e_ptr[n];
    p.n2e.resize(p.n2e_ptr.back());
    {
        std::vector< long long > fill(p.n2e_ptr.begin(), p.n2e_ptr.end() - 1);
        for (long long e = 0; e < n_elems; ++e)
            for (int a = 0; a < nn; ++a)
                p.n2e[fill[nodes[e * nn + a]]++] = static_cast< uint32_t >(e);
    }
    // npart: given (METIS' node partition, un-condensed as in uncondenseNodes :189-207), or — when absent — the lowest part among the
    // elements holding the node (a valid nodal partition; no node is then "disjoint")
    p.npart.resize(n_nodes);
    for (long long n = 0; n < n_nodes; ++n)
    {
        if (p.n2e_ptr[n] == p.n2e_ptr[n + 1])
            throw std::invalid_argument{"At least one node in the mesh does not belong to any element"};
        if (npart_in)
        {
            if (npart_in[n] < 0 or npart_in[n] >= n_parts)
                throw std::invalid_argument{"npart entry out of range"};
            p.npart[n] = npart_in[n];
        }
        else
        {
            int32_t lo = n_parts;
            for (long long k = p.n2e_ptr[n]; k < p.n2e_ptr[n + 1]; ++k)
                lo = std::min(lo, epart[p.n2e[k]]);
            p.npart[n] = lo;
        }
    }
    // assignNodes (:354-395): per part the nodes of its elements, split into owned (npart == part) and ghost
    std::vector< std::vector< long long > > owned(n_parts), ghost(n_parts);
    {
        std::vector< int32_t > seen(n_nodes, -1); // last part that recorded the node
        std::vector< std::vector< long long > > elems_of(n_parts);
        for (long long e = 0; e < n_elems; ++e)
            elems_of[epart[e]].push_back(e);
        for (int part = 0; part < n_parts; ++part)
        {
            for (long long e : elems_of[part])
                for (int a = 0; a < nn; ++a)
                {
                    const long long n = nodes[e * nn + a];
                    if (seen[n] == part)
                        continue;
                    seen[n] = part;
                    (p.npart[n] == part ? owned[part] : ghost[part]).push_back(n);
                }
            std::sort(owned[part].begin(), owned[part].end());
            std::sort(ghost[part].begin(), ghost[part].end());
        }
        // disjoint nodes: assigned to a part none of whose elements holds them; collected part by part, node ids ascending (:372-377)
        std::vector< long long > disjoint;
        {
            std::vector< char > in_own_part(n_nodes, 0);
            for (int part = 0; part < n_parts; ++part)
                for (long long n : owned[part])
                    in_own_part[n] = 1;
            for (int part = 0; part < n_parts; ++part)
                for (long long n = 0; n < n_nodes; ++n)
                    if (p.npart[n] == part and not in_own_part[n])
                        disjoint.push_back(n);
        }
        // reassignDisjointNodes (:322-351): the first part (ascending) that holds the node as a ghost claims it
        for (int part = 0; part < n_parts and not disjoint.empty(); ++part)
        {
            std::vector< long long > claimed, rest;
            for (long long n : disjoint)
            {
                const auto it = std::lower_bound(ghost[part].begin(), ghost[part].end(), n);
                if (it != ghost[part].end() and *it == n)
                {
                    claimed.push_back(n);
                    ghost[part].erase(it);
                }
                else
                    rest.push_back(n);
            }
            disjoint.swap(rest);
            for (long long n : claimed)
                p.npart[n] = part;
            owned[part].insert(owned[part].end(), claimed.begin(), claimed.end());
            std::sort(owned[part].begin(), owned[part].end());
        }
        if (not disjoint.empty())
            throw std::invalid_argument{"At least one node in the mesh does not belong to any element"};
    }
    // renumberNodes (:411-440): part by part, owned nodes in ascending old id
    p.new_id.assign(n_nodes, -1);
    p.dist.assign(n_parts + 1, 0);
    {
        long long next = 0;
        for (int part = 0; part < n_parts; ++part)
        {
            p.dist[part] = next;
            for (long long n : owned[part])
                p.new_id[n] = next++;
        }
        p.dist[n_parts] = next;
        if (next != n_nodes)
            throw std::logic_error{"node renumbering does not cover the mesh"};
    }
    p.ranks.resize(n_parts);
    for (int part = 0; part < n_parts; ++part)
    {
        auto& r = p.ranks[part];
        r.ghosts.reserve(ghost[part].size());
        for (long long n : ghost[part])
            r.ghosts.push_back(p.new_id[n]);
        std::sort(r.ghosts.begin(), r.ghosts.end());
    }
    // elements of every rank: input order, then the ones touching a ghost node first (stable) — mesh/SplitMesh.hpp's border / interior
    for (long long e = 0; e < n_elems; ++e)
        p.ranks[epart[e]].elems.push_back(e);
    for (int part = 0; part < n_parts; ++part)
    {
        auto&      r       = p.ranks[part];
        const auto touches = [&](long long e) {
            for (int a = 0; a < nn; ++a)
                if (p.npart[nodes[e * nn + a]] != part)
                    return true;
            return false;
        };
        const auto mid = std::stable_partition(r.elems.begin(), r.elems.end(), touches);
        r.n_border     = mid - r.elems.begin();
    }
    // extended ghosts: plus the columns of the neighbours' contributions to my rows (makeColumnOwnership, SparsityGraph.hpp:221-233)
    for (int part = 0; part < n_parts; ++part)
    {
        auto&                    r = p.ranks[part];
        std::vector< long long > ext(r.ghosts);
        for (long long n : owned[part])
            for (long long k = p.n2e_ptr[n]; k < p.n2e_ptr[n + 1]; ++k)
            {
                const long long e = p.n2e[k];
                if (epart[e] == part)
                    continue;
                for (int a = 0; a < nn; ++a)
                {
                    const long long m = nodes[e * nn + a];
                    if (p.npart[m] != part)
                        ext.push_back(p.new_id[m]);
                }
            }
        std::sort(ext.begin(), ext.end());
        ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
        r.ghosts_ext.swap(ext);
    }
    return p;
}

// comm::ImportExportContext of rank r at node granularity: shared neighbours = owners of my ghosts with their contiguous ranges of
// the ghost block; owned neighbours = ranks whose ghost block holds nodes of mine, with those nodes (my local ids) in the order of
// their ghost block (ascending global id)
struct HaloLists
{
    std::vector< int >       owned_nbrs, shared_nbrs;
    std::vector< long long > owned_ptr{0}, shared_off{0};
    std::vector< int32_t >   owned_nodes;
};
inline HaloLists makeHaloLists(const Partition& p, int r, bool extended)
{
    HaloLists h;
    const auto& mine = p.ghostsOf(r, extended);
    for (size_t i = 0; i < mine.size();)
    {
        const int owner = p.ownerOf(mine[i]);
        size_t    j     = i;
        while (j < mine.size() and mine[j] < p.dist[owner + 1])
            ++j;
        h.shared_nbrs.push_back(owner);
        h.shared_off.push_back(static_cast< long long >(j));
        i = j;
    }
    for (int q = 0; q < p.n_parts; ++q)
    {
        if (q == r)
            continue;
        const auto& theirs = p.ghostsOf(q, extended);
        const auto  lo     = std::lower_bound(theirs.begin(), theirs.end(), p.dist[r]);
        const auto  hi     = std::lower_bound(theirs.begin(), theirs.end(), p.dist[r + 1]);
        if (lo == hi)
            continue;
        h.owned_nbrs.push_back(q);
        for (auto it = lo; it != hi; ++it)
            h.owned_nodes.push_back(static_cast< int32_t >(*it - p.dist[r]));
        h.owned_ptr.push_back(static_cast< long long >(h.owned_nodes.size()));
    }
    return h;
}

// nodes (global new ids) of the row of node `n_old` (input numbering) restricted to the elements of part `part` (or of all parts: -1)
inline void rowNodes(const Partition& p, long long n_old, int part, std::vector< long long >& out)
{
    out.clear();
    for (long long k = p.n2e_ptr[n_old]; k < p.n2e_ptr[n_old + 1]; ++k)
    {
        const long long e = p.n2e[k];
        if (part >= 0 and p.epart[e] != part)
            continue;
        for (int a = 0; a < p.nn; ++a)
            out.push_back(p.new_id[p.nodes[e * p.nn + a]]);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

// node graph of rank r in the extended view: owned rows complete (the contributions of every rank's elements), rows of ghost nodes as
// my own elements fill them, rows of extra nodes empty; columns = extended local ids, ascending (SparsityGraph.hpp:254-278)
inline NodeGraph makeRankGraph(const Partition& p, int r)
{
    // inverse of new_id, to walk my local nodes in local order
    std::vector< long long > old_of(p.n_nodes);
    for (long long n = 0; n < p.n_nodes; ++n)
        old_of[p.new_id[n]] = n;
    const long long          n_owned = p.nOwned(r), n_local = p.nLocal(r, true);
    NodeGraph                g;
    g.ptr.assign(n_local + 1, 0);
    std::vector< long long > row;
    std::vector< uint32_t >  cols;
    for (int pass = 0; pass < 2; ++pass)
    {
        for (long long l = 0; l < n_local; ++l)
        {
            const long long gid = l < n_owned ? p.dist[r] + l : p.ranks[r].ghosts_ext[l - n_owned];
            rowNodes(p, old_of[gid], l < n_owned ? -1 : r, row);
            if (pass == 0)
            {
                g.ptr[l + 1] = g.ptr[l] + static_cast< long long >(row.size());
                continue;
            }
            cols.clear();
            for (long long c : row)
                cols.push_back(static_cast< uint32_t >(p.localId(r, true, c)));
            std::sort(cols.begin(), cols.end());
            std::copy(cols.begin(), cols.end(), g.nbr.begin() + g.ptr[l]);
        }
        if (pass == 0)
            g.nbr.resize(g.ptr.back());
    }
    return g;
}

// receive side of the shared-row export of rank r (extended view): for every owned neighbour q (in the order of makeHaloLists) and every
// node of mine in q's ghost block, q's row of that node (the columns its elements touch, in q's local column order: q's owned nodes by
// global id, then the others by global id) and, per entry, the position of that column in MY row of the node.
struct RowExportPlan
{
    std::vector< long long > entry_ptr{0}; // per (neighbour, shared node) in halo order
    std::vector< uint32_t >  pos;
};
inline RowExportPlan makeRowExportPlan(const Partition& p, int r, const HaloLists& halo, const NodeGraph& my_graph)
{
    std::vector< long long > old_of(p.n_nodes);
    for (long long n = 0; n < p.n_nodes; ++n)
        old_of[p.new_id[n]] = n;
    RowExportPlan            plan;
    std::vector< long long > row;
    for (size_t k = 0; k < halo.owned_nbrs.size(); ++k)
    {
        const int q = halo.owned_nbrs[k];
        for (long long i = halo.owned_ptr[k]; i < halo.owned_ptr[k + 1]; ++i)
        {
            const long long lid = halo.owned_nodes[i], gid = p.dist[r] + lid;
            rowNodes(p, old_of[gid], q, row);
            // q's column order: its owned nodes first
            std::stable_partition(row.begin(), row.end(), [&](long long c) { return c >= p.dist[q] and c < p.dist[q + 1]; });
            const uint32_t* beg = my_graph.nbr.data() + my_graph.ptr[lid];
            const uint32_t* end = my_graph.nbr.data() + my_graph.ptr[lid + 1];
            for (long long c : row)
            {
                const auto want = static_cast< uint32_t >(p.localId(r, true, c));
                const auto it   = std::lower_bound(beg, end, want);
                if (it == end or *it != want)
                    throw std::logic_error{"shared-row export: a neighbour's column is missing from the owner's row"};
                plan.pos.push_back(static_cast< uint32_t >(it - beg));
            }
            plan.entry_ptr.push_back(static_cast< long long >(plan.pos.size()));
        }
    }
    return plan;
}
} // namespace l3b::host
#endif
