// Host-side mesh front end of the B200 path: structured quad/hex generators with the reference's order-p node numbering,
// and the node-level sparsity graph.
//
// What it reproduces (bit-exact integer contracts, checked against the oracle's restatement of the reference algorithms):
//   * mesh/primitives/CubeMesh.hpp:16-138, SquareMesh.hpp:14-76 — element order (x fastest), vertex order, boundary ids;
//   * mesh/ConvertMeshToOrder.hpp:52-104 — ids of the order-p nodes. The reference discovers shared nodes by matching
//     physical node locations of dual-graph neighbours; on a structured grid the outcome has a closed form: a node is
//     created by the lowest-id element containing it, and each element numbers its new nodes boundary-first, then
//     interior, in ascending local index. That closed form is what is implemented here (no hashing, no matching);
//   * algsys/SparsityGraph.hpp:25-81, 254-278 — per row the sorted union of the dofs of all elements containing the row's
//     node. Stored at node granularity: all dofs of a node share one pattern and the dofs of a column node are contiguous
//     (dofs/NodeToDofMap.hpp:249-264), so the dof-level CRS is the U-fold expansion of the node graph.
#ifndef L3B_MESH_HOST_HPP
#define L3B_MESH_HOST_HPP

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace l3b::host
{
constexpr uint16_t no_boundary = 0xFFFF;

struct Mesh
{
    int                     dim = 0, order = 1;
    long long               n_nodes = 0, n_elems = 0;
    int                     nodes_per_elem = 0, n_sides = 0;
    std::vector< uint32_t > nodes;    // [n_elems][nodes_per_elem], lexicographic (x fastest) local order
    std::vector< double >   verts;    // [n_elems][2^dim][3]
    std::vector< uint16_t > side_bnd; // [n_elems][n_sides]: boundary domain id on that side or no_boundary
};

inline bool isBoundaryLocal(int dim, int nb, int a)
{
    const int i = a % nb, j = (a / nb) % nb, k = a / (nb * nb);
    const auto edge = [&](int v) { return v == 0 or v == nb - 1; };
    return dim == 2 ? (edge(i) or edge(j)) : (edge(i) or edge(j) or edge(k));
}

// ex, ey, ez: number of elements per direction (ez = 1 and dim = 2 for the square)
inline Mesh makeStructured(int dim, const std::vector< double >& xs, const std::vector< double >& ys, const std::vector< double >& zs, int order)
{
    Mesh m;
    m.dim   = dim;
    m.order = order;
    const long long ndx = static_cast< long long >(xs.size()), ndy = static_cast< long long >(ys.size()),
                    ndz = dim == 3 ? static_cast< long long >(zs.size()) : 1;
    const long long ex = ndx - 1, ey = ndy - 1, ez = dim == 3 ? ndz - 1 : 1;
    if (ex < 1 or ey < 1 or ez < 1)
        throw std::invalid_argument{"structured mesh needs at least 2 grid points per direction"};
    const int nb = order + 1, p = order;
    m.nodes_per_elem = dim == 3 ? nb * nb * nb : nb * nb;
    m.n_sides        = 2 * dim;
    m.n_elems        = ex * ey * ez;
    const int nv     = 1 << dim;
    m.nodes.assign(static_cast< size_t >(m.n_elems) * m.nodes_per_elem, 0);
    m.verts.resize(static_cast< size_t >(m.n_elems) * nv * 3);
    m.side_bnd.assign(static_cast< size_t >(m.n_elems) * m.n_sides, no_boundary);
    // traversal order of the local nodes when creating ids: boundary nodes ascending, then internal ascending
    std::vector< int > creation_order;
    for (int a = 0; a < m.nodes_per_elem; ++a)
        if (isBoundaryLocal(dim, nb, a))
            creation_order.push_back(a);
    for (int a = 0; a < m.nodes_per_elem; ++a)
        if (not isBoundaryLocal(dim, nb, a))
            creation_order.push_back(a);
    long long next_node = ndx * ndy * ndz; // order-1 nodes keep their ids
    if (next_node + m.n_elems * static_cast< long long >(m.nodes_per_elem) > 0xFFFFFFFFll and order > 1)
        if (static_cast< long long >(ex * p + 1) * (ey * p + 1) * (dim == 3 ? ez * p + 1 : 1) > 0xFFFFFFFFll)
            throw std::overflow_error{"local node ids exceed 32 bits"};
    for (long long kz = 0; kz < ez; ++kz)
        for (long long ky = 0; ky < ey; ++ky)
            for (long long kx = 0; kx < ex; ++kx)
            {
                const long long e  = kx + ex * (ky + ey * kz);
                uint32_t*       en = &m.nodes[static_cast< size_t >(e) * m.nodes_per_elem];
                for (int v = 0; v < nv; ++v)
                {
                    double* vp = &m.verts[(static_cast< size_t >(e) * nv + v) * 3];
                    vp[0]      = xs[kx + (v & 1)];
                    vp[1]      = ys[ky + ((v >> 1) & 1)];
                    vp[2]      = dim == 3 ? zs[kz + ((v >> 2) & 1)] : 0.;
                }
                for (int a : creation_order)
                {
                    const int  i = a % nb, j = (a / nb) % nb, k = dim == 3 ? a / (nb * nb) : 0;
                    const bool vx = i == 0 or i == p, vy = j == 0 or j == p, vz = dim == 2 or k == 0 or k == p;
                    if (vx and vy and vz) // vertex of the order-1 grid
                    {
                        const long long gx = kx + (i == p), gy = ky + (j == p), gz = dim == 3 ? kz + (k == p) : 0;
                        en[a]              = static_cast< uint32_t >(gx + ndx * (gy + ndy * gz));
                        continue;
                    }
                    // owner = lowest-id element containing the node
                    const bool sx = i == 0 and kx > 0, sy = j == 0 and ky > 0, sz = dim == 3 and k == 0 and kz > 0;
                    if (not(sx or sy or sz))
                        en[a] = static_cast< uint32_t >(next_node++);
                    else
                    {
                        const long long oe = (kx - sx) + ex * ((ky - sy) + ey * (kz - sz));
                        const int       oa = (sx ? p : i) + nb * ((sy ? p : j) + nb * (sz ? p : k));
                        en[a]              = m.nodes[static_cast< size_t >(oe) * m.nodes_per_elem + oa];
                    }
                }
                // sides on the domain boundary (mesh/ElementTraits.hpp:88-93, CubeMesh.hpp / SquareMesh.hpp ids)
                uint16_t* sb = &m.side_bnd[static_cast< size_t >(e) * m.n_sides];
                if (dim == 3)
                {
                    if (kz == 0)
                        sb[0] = 1; // z-min face: side 0, boundary id "back" = 1
                    if (kz == ez - 1)
                        sb[1] = 2;
                    if (ky == 0)
                        sb[2] = 3;
                    if (ky == ey - 1)
                        sb[3] = 4;
                    if (kx == 0)
                        sb[4] = 5;
                    if (kx == ex - 1)
                        sb[5] = 6;
                }
                else
                {
                    if (ky == 0)
                        sb[0] = 1; // bottom
                    if (ky == ey - 1)
                        sb[1] = 2; // top
                    if (kx == 0)
                        sb[2] = 3; // left
                    if (kx == ex - 1)
                        sb[3] = 4; // right
                }
            }
    m.n_nodes = next_node;
    return m;
}

// node-level CSR graph: for each node the sorted list of nodes sharing an element with it (itself included)
struct NodeGraph
{
    std::vector< long long > ptr;
    std::vector< uint32_t >  nbr;
};

inline NodeGraph makeNodeGraph(long long n_nodes, long long n_elems, int nn, const uint32_t* nodes)
{
    // node → elements
    std::vector< long long > n2e_ptr(n_nodes + 1, 0);
    for (long long i = 0; i < n_elems * nn; ++i)
        ++n2e_ptr[nodes[i] + 1];
    for (long long n = 0; n < n_nodes; ++n)
        n2e_ptr[n + 1] += n2e_ptr[n];
    std::vector< uint32_t >  n2e(n2e_ptr.back());
    std::vector< long long > fill(n2e_ptr.begin(), n2e_ptr.end() - 1);
    for (long long e = 0; e < n_elems; ++e)
        for (int a = 0; a < nn; ++a)
            n2e[fill[nodes[e * nn + a]]++] = static_cast< uint32_t >(e);
    NodeGraph g;
    g.ptr.assign(n_nodes + 1, 0);
    std::vector< uint32_t > scratch;
    // two passes: count, fill
    for (int pass = 0; pass < 2; ++pass)
    {
        for (long long n = 0; n < n_nodes; ++n)
        {
            scratch.clear();
            for (long long k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k)
            {
                const uint32_t* en = nodes + static_cast< long long >(n2e[k]) * nn;
                scratch.insert(scratch.end(), en, en + nn);
            }
            std::sort(scratch.begin(), scratch.end());
            scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
            if (pass == 0)
                g.ptr[n + 1] = g.ptr[n] + static_cast< long long >(scratch.size());
            else
                std::copy(scratch.begin(), scratch.end(), g.nbr.begin() + g.ptr[n]);
        }
        if (pass == 0)
            g.nbr.resize(g.ptr.back());
    }
    return g;
}
} // namespace l3b::host
#endif
