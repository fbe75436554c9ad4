// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp). mesh/ restatement: structured primitives, element
// traits tables, order conversion (the node numbering the DOF maps and sparsity graph are built on), boundary matching.
#include "l3ster_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <unordered_map>

namespace orc
{
// mesh/ElementTraits.hpp:72-98 (hex), :118-137 (quad), :153-157 (line)
std::vector< int > sideNodeInds(ElementType et, int order, int side)
{
    const int          n = order + 1;
    std::vector< int > r;
    if (et == Hex)
    {
        const int nps = n * n, back_shift = nps * (n - 1), top_shift = n * (n - 1), right_shift = n - 1;
        int       index = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
            {
                int v = 0;
                switch (side)
                {
                case 0:
                    v = index;
                    break;
                case 1:
                    v = index + back_shift;
                    break;
                case 2:
                    v = i * nps + j;
                    break;
                case 3:
                    v = i * nps + j + top_shift;
                    break;
                case 4:
                    v = i * nps + j * n;
                    break;
                case 5:
                    v = i * nps + j * n + right_shift;
                    break;
                default:
                    throw std::out_of_range{"hex side"};
                }
                r.push_back(v);
                ++index;
            }
    }
    else if (et == Quad)
    {
        const int top_shift = n * (n - 1), right_shift = n - 1;
        for (int i = 0; i < n; ++i)
            switch (side)
            {
            case 0:
                r.push_back(i);
                break;
            case 1:
                r.push_back(i + top_shift);
                break;
            case 2:
                r.push_back(i * n);
                break;
            case 3:
                r.push_back(i * n + right_shift);
                break;
            default:
                throw std::out_of_range{"quad side"};
            }
    }
    else
        r.push_back(side == 0 ? 0 : order);
    return r;
}

namespace
{
struct Traits
{
    std::vector< int > boundary, internal;
};
const Traits& traits(ElementType et, int order)
{
    static std::map< std::pair< int, int >, Traits > cache;
    static std::mutex                                mtx;
    std::lock_guard                                  lock{mtx};
    const auto                                       key = std::make_pair(static_cast< int >(et), order);
    if (auto it = cache.find(key); it != cache.end())
        return it->second;
    Traits t;
    for (int s = 0; s < numSides(et); ++s)
        for (int i : sideNodeInds(et, order, s))
            t.boundary.push_back(i);
    std::sort(t.boundary.begin(), t.boundary.end());
    t.boundary.erase(std::unique(t.boundary.begin(), t.boundary.end()), t.boundary.end());
    for (int i = 0; i < numNodes(et, order); ++i)
        if (not std::binary_search(t.boundary.begin(), t.boundary.end(), i))
            t.internal.push_back(i);
    return cache.emplace(key, std::move(t)).first->second;
}
} // namespace

// mesh/ElementTraits.hpp:29-57
const std::vector< int >& boundaryNodeInds(ElementType et, int order)
{
    return traits(et, order).boundary;
}
const std::vector< int >& internalNodeInds(ElementType et, int order)
{
    return traits(et, order).internal;
}

// mesh/primitives/CubeMesh.hpp:16-138 (ids: domain 0, back 1, front 2, bottom 3, top 4, left 5, right 6)
Mesh makeCubeMesh(const std::vector< val_t >& distx, const std::vector< val_t >& disty, const std::vector< val_t >& distz)
{
    const std::size_t n_dx = distx.size(), n_dy = disty.size(), n_dz = distz.size();
    const std::size_t e_dx = n_dx - 1, e_dy = n_dy - 1, e_dz = n_dz - 1;
    Mesh              m;
    m.et      = Hex;
    m.order   = 1;
    m.n_nodes = n_dx * n_dy * n_dz;
    m.n_elems = e_dx * e_dy * e_dz;
    n_id_t el_ind = 0;
    for (std::size_t iz = 0; iz < e_dz; ++iz)
        for (std::size_t iy = 0; iy < e_dy; ++iy)
            for (std::size_t ix = 0; ix < e_dx; ++ix)
            {
                const n_id_t base = n_dx * n_dy * iz + n_dx * iy + ix;
                const n_id_t nodes[8] = {base,
                                         base + 1,
                                         base + n_dx,
                                         base + n_dx + 1,
                                         base + n_dx * n_dy,
                                         base + n_dx * n_dy + 1,
                                         base + n_dx * n_dy + n_dx,
                                         base + n_dx * n_dy + n_dx + 1};
                m.elem_nodes.insert(m.elem_nodes.end(), nodes, nodes + 8);
                for (int v = 0; v < 8; ++v)
                {
                    m.elem_verts.push_back(distx[ix + (v & 1)]);
                    m.elem_verts.push_back(disty[iy + ((v >> 1) & 1)]);
                    m.elem_verts.push_back(distz[iz + ((v >> 2) & 1)]);
                }
                m.elem_ids.push_back(el_ind++);
            }
    const auto push_face = [&](int domain, std::array< n_id_t, 4 > nodes, std::array< std::array< val_t, 3 >, 4 > verts) {
        Mesh::BoundaryElem b;
        b.domain_id = domain;
        b.id        = el_ind++;
        b.nodes.assign(nodes.begin(), nodes.end());
        for (const auto& v : verts)
            b.verts.insert(b.verts.end(), v.begin(), v.end());
        m.boundary.push_back(std::move(b));
    };
    // z = const faces
    for (std::size_t iy = 0; iy < e_dy; ++iy)
        for (std::size_t ix = 0; ix < e_dx; ++ix)
        {
            std::array< n_id_t, 4 > n1 = {n_dx * iy + ix, n_dx * iy + ix + 1, n_dx * (iy + 1) + ix, n_dx * (iy + 1) + ix + 1};
            std::array< std::array< val_t, 3 >, 4 > v1 = {{{distx[ix], disty[iy], distz[0]},
                                                          {distx[ix + 1], disty[iy], distz[0]},
                                                          {distx[ix], disty[iy + 1], distz[0]},
                                                          {distx[ix + 1], disty[iy + 1], distz[0]}}};
            auto n2 = n1;
            for (auto& n : n2)
                n += n_dx * n_dy * e_dz;
            auto v2 = v1;
            for (auto& v : v2)
                v[2] = distz[e_dz];
            push_face(1, n1, v1);
            push_face(2, n2, v2);
        }
    // y = const faces
    for (std::size_t iz = 0; iz < e_dz; ++iz)
        for (std::size_t ix = 0; ix < e_dx; ++ix)
        {
            std::array< n_id_t, 4 > n1 = {
                n_dx * n_dy * iz + ix, n_dx * n_dy * iz + ix + 1, n_dx * n_dy * (iz + 1) + ix, n_dx * n_dy * (iz + 1) + ix + 1};
            std::array< std::array< val_t, 3 >, 4 > v1 = {{{distx[ix], disty[0], distz[iz]},
                                                          {distx[ix + 1], disty[0], distz[iz]},
                                                          {distx[ix], disty[0], distz[iz + 1]},
                                                          {distx[ix + 1], disty[0], distz[iz + 1]}}};
            auto n2 = n1;
            for (auto& n : n2)
                n += n_dx * e_dy;
            auto v2 = v1;
            for (auto& v : v2)
                v[1] = disty[e_dy];
            push_face(3, n1, v1);
            push_face(4, n2, v2);
        }
    // x = const faces
    for (std::size_t iz = 0; iz < e_dz; ++iz)
        for (std::size_t iy = 0; iy < e_dy; ++iy)
        {
            std::array< n_id_t, 4 > n1 = {n_dx * n_dy * iz + n_dx * iy,
                                          n_dx * n_dy * iz + n_dx * (iy + 1),
                                          n_dx * n_dy * (iz + 1) + n_dx * iy,
                                          n_dx * n_dy * (iz + 1) + n_dx * (iy + 1)};
            std::array< std::array< val_t, 3 >, 4 > v1 = {{{distx[0], disty[iy], distz[iz]},
                                                          {distx[0], disty[iy + 1], distz[iz]},
                                                          {distx[0], disty[iy], distz[iz + 1]},
                                                          {distx[0], disty[iy + 1], distz[iz + 1]}}};
            auto n2 = n1;
            for (auto& n : n2)
                n += e_dx;
            auto v2 = v1;
            for (auto& v : v2)
                v[0] = distx[e_dx];
            push_face(5, n1, v1);
            push_face(6, n2, v2);
        }
    return m;
}

// mesh/primitives/SquareMesh.hpp:14-76 (ids: domain 0, bottom 1, top 2, left 3, right 4)
Mesh makeSquareMesh(const std::vector< val_t >& distx, const std::vector< val_t >& disty)
{
    const std::size_t n_dx = distx.size(), n_dy = disty.size(), e_dx = n_dx - 1, e_dy = n_dy - 1;
    Mesh              m;
    m.et      = Quad;
    m.order   = 1;
    m.n_nodes = n_dx * n_dy;
    m.n_elems = e_dx * e_dy;
    n_id_t el_ind = 0;
    for (std::size_t iy = 0; iy < e_dy; ++iy)
        for (std::size_t ix = 0; ix < e_dx; ++ix)
        {
            const n_id_t nodes[4] = {n_dx * iy + ix, n_dx * iy + ix + 1, n_dx * (iy + 1) + ix, n_dx * (iy + 1) + ix + 1};
            m.elem_nodes.insert(m.elem_nodes.end(), nodes, nodes + 4);
            for (int v = 0; v < 4; ++v)
            {
                m.elem_verts.push_back(distx[ix + (v & 1)]);
                m.elem_verts.push_back(disty[iy + ((v >> 1) & 1)]);
                m.elem_verts.push_back(0.);
            }
            m.elem_ids.push_back(el_ind++);
        }
    const auto push_edge = [&](int domain, std::array< n_id_t, 2 > nodes, std::array< std::array< val_t, 3 >, 2 > verts) {
        Mesh::BoundaryElem b;
        b.domain_id = domain;
        b.id        = el_ind++;
        b.nodes.assign(nodes.begin(), nodes.end());
        for (const auto& v : verts)
            b.verts.insert(b.verts.end(), v.begin(), v.end());
        m.boundary.push_back(std::move(b));
    };
    for (std::size_t ix = 0; ix < e_dx; ++ix)
    {
        std::array< n_id_t, 2 >                 n1 = {ix, ix + 1};
        std::array< std::array< val_t, 3 >, 2 > v1 = {{{distx[ix], disty[0], 0.}, {distx[ix + 1], disty[0], 0.}}};
        auto                                    n2 = n1;
        for (auto& n : n2)
            n += n_dx * e_dy;
        auto v2 = v1;
        for (auto& v : v2)
            v[1] = disty[e_dy];
        push_edge(1, n1, v1);
        push_edge(2, n2, v2);
    }
    for (std::size_t iy = 0; iy < e_dy; ++iy)
    {
        std::array< n_id_t, 2 >                 n1 = {iy * n_dx, (iy + 1) * n_dx};
        std::array< std::array< val_t, 3 >, 2 > v1 = {{{distx[0], disty[iy], 0.}, {distx[0], disty[iy + 1], 0.}}};
        auto                                    n2 = n1;
        for (auto& n : n2)
            n += e_dx;
        auto v2 = v1;
        for (auto& v : v2)
            v[0] = distx[e_dx];
        push_edge(3, n1, v1);
        push_edge(4, n2, v2);
    }
    return m;
}

namespace
{
ElementType lowerType(ElementType et)
{
    return et == Hex ? Quad : Line;
}

// mesh/ElementIntersecting.hpp:83-101 element_outer_features<T, O, DIM>
std::vector< std::vector< int > > outerFeatures(ElementType et, int order, int dim)
{
    const int                         nd = nativeDim(et);
    std::vector< std::vector< int > > r;
    if (dim == 0 or nd < dim)
        return r;
    if (nd == dim)
    {
        std::vector< int > all(numNodes(et, order));
        for (int i = 0; i < static_cast< int >(all.size()); ++i)
            all[i] = i;
        r.push_back(std::move(all));
    }
    else if (dim == nd - 1)
        for (int s = 0; s < numSides(et); ++s)
            r.push_back(sideNodeInds(et, order, s));
    else // hex edges: pairwise side intersections holding >= 2 nodes (:17-76)
        for (int s1 = 0; s1 < numSides(et) - 1; ++s1)
            for (int s2 = s1 + 1; s2 < numSides(et); ++s2)
            {
                auto f1 = sideNodeInds(et, order, s1), f2 = sideNodeInds(et, order, s2);
                std::sort(f1.begin(), f1.end());
                std::sort(f2.begin(), f2.end());
                std::vector< int > inter;
                std::set_intersection(f1.begin(), f1.end(), f2.begin(), f2.end(), std::back_inserter(inter));
                if (inter.size() >= 2)
                    r.push_back(std::move(inter));
            }
    return r;
}

// mesh/NodeReferenceLocation.hpp:13-57 + NodePhysicalLocation.hpp:21-25
std::array< val_t, 3 > nodePhysLoc(ElementType et, int order, const val_t* verts, int i)
{
    const auto& absc = lobattoAbsc(order + 1);
    const int   n    = order + 1;
    val_t       xi[3] = {absc[i % n], 0., 0.};
    if (et != Line)
        xi[1] = absc[(i / n) % n];
    if (et == Hex)
        xi[2] = absc[i / (n * n)];
    std::array< val_t, 3 > out{};
    mapToPhysicalSpace(et, verts, xi, out.data());
    return out;
}

bool pointsMatch(const std::array< val_t, 3 >& a, const std::array< val_t, 3 >& b) // ElementIntersecting.hpp:171-177
{
    const val_t dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return std::sqrt(dx * dx + dy * dy + dz * dz) < 1e-12;
}

struct AnyElem // uniform view over volume and boundary elements of the order-1 mesh
{
    ElementType   et;
    const n_id_t* nodes1;
    const val_t*  verts;
};
} // namespace

// mesh/ConvertMeshToOrder.hpp:52-104 (+ ElementIntersecting.hpp:103-228, MeshUtils.hpp:45-95 with ncommon = 2)
Mesh convertMeshToOrder(const Mesh& mesh, int OC)
{
    if (mesh.order != 1)
        throw std::invalid_argument{"convertMeshToOrder expects an order-1 mesh"};
    const ElementType vt = mesh.et, bt = lowerType(mesh.et);
    const int         nv1 = numNodes(vt, 1);
    const std::size_t n_all = mesh.n_elems + mesh.boundary.size();
    // elements by id: volume elements first (ids 0..), then boundary elements in generation order — this is also the
    // conversion order (domains ascending, elements by ascending id within a domain) as far as new node ids are
    // concerned, because boundary elements never create nodes
    std::vector< AnyElem > elems(n_all);
    for (std::size_t e = 0; e < mesh.n_elems; ++e)
        elems[e] = {vt, &mesh.elem_nodes[e * nv1], &mesh.elem_verts[e * nv1 * 3]};
    for (std::size_t b = 0; b < mesh.boundary.size(); ++b)
        elems[mesh.n_elems + b] = {bt, mesh.boundary[b].nodes.data(), mesh.boundary[b].verts.data()};

    // dual graph: elements sharing >= 2 nodes (METIS_MeshToDual, ncommon = 2)
    std::vector< std::vector< std::size_t > > node2el(mesh.n_nodes);
    for (std::size_t e = 0; e < n_all; ++e)
        for (int i = 0; i < numNodes(elems[e].et, 1); ++i)
            node2el[elems[e].nodes1[i]].push_back(e);
    const auto neighbours = [&](std::size_t e) {
        std::unordered_map< std::size_t, int > cnt;
        for (int i = 0; i < numNodes(elems[e].et, 1); ++i)
            for (auto o : node2el[elems[e].nodes1[i]])
                if (o != e)
                    ++cnt[o];
        std::vector< std::size_t > r;
        for (auto [o, c] : cnt)
            if (c >= 2)
                r.push_back(o);
        std::sort(r.begin(), r.end());
        return r;
    };

    Mesh out;
    out.et         = vt;
    out.order      = OC;
    out.n_elems    = mesh.n_elems;
    out.elem_verts = mesh.elem_verts;
    out.elem_ids   = mesh.elem_ids;
    const int nvN = numNodes(vt, OC), nbN = numNodes(bt, OC);
    out.elem_nodes.assign(mesh.n_elems * nvN, 0);
    out.boundary = mesh.boundary;
    std::vector< std::vector< n_id_t > > new_nodes_of(n_all);
    std::vector< char >                  converted(n_all, 0);
    n_id_t                               max_node = mesh.n_nodes;
    std::vector< std::vector< std::array< val_t, 3 > > > locs(n_all); // order-OC node locations, filled on conversion

    const auto convert = [&](std::size_t e) {
        const auto&           el  = elems[e];
        const int             nN  = numNodes(el.et, OC), n1 = numNodes(el.et, 1);
        std::vector< char >   mask(nN, 0);
        std::vector< n_id_t > nn(nN, 0);
        auto&                 my_locs = locs[e];
        my_locs.resize(nN);
        for (int iN = 0; iN < nN; ++iN)
            my_locs[iN] = nodePhysLoc(el.et, OC, el.verts, iN);
        // updateMatchMask<OC>(el, mask, nodes): the order-1 (vertex) nodes (ElementIntersecting.hpp:180-195)
        for (int i1 = 0; i1 < n1; ++i1)
        {
            const auto p1 = nodePhysLoc(el.et, 1, el.verts, i1);
            for (int iN = 0; iN < nN; ++iN)
                if (pointsMatch(my_locs[iN], p1))
                {
                    mask[iN] = 1;
                    nn[iN]   = el.nodes1[i1];
                    break;
                }
        }
        for (auto nbr : neighbours(e))
        {
            if (not converted[nbr])
                continue;
            const auto& pat = elems[nbr];
            // elementIntersection (ElementIntersecting.hpp:147-169)
            const int d1 = nativeDim(pat.et), d2 = nativeDim(el.et);
            const int highest = d1 == d2 ? d1 - 1 : std::min(d1, d2);
            if (highest == 0)
                continue;
            std::vector< int > pat_inds, match_inds;
            for (int dim = highest; dim >= 1 and pat_inds.empty(); --dim)
            {
                const auto f1_o1 = outerFeatures(pat.et, 1, dim), f2_o1 = outerFeatures(el.et, 1, dim);
                const auto f1_oN = outerFeatures(pat.et, OC, dim), f2_oN = outerFeatures(el.et, OC, dim);
                for (std::size_t a = 0; a < f1_o1.size() and pat_inds.empty(); ++a)
                {
                    std::vector< n_id_t > s1;
                    for (int i : f1_o1[a])
                        s1.push_back(pat.nodes1[i]);
                    std::sort(s1.begin(), s1.end());
                    for (std::size_t b = 0; b < f2_o1.size(); ++b)
                    {
                        std::vector< n_id_t > s2;
                        for (int i : f2_o1[b])
                            s2.push_back(el.nodes1[i]);
                        std::sort(s2.begin(), s2.end());
                        if (s1 == s2)
                        {
                            pat_inds   = f1_oN[a];
                            match_inds = f2_oN[b];
                            break;
                        }
                    }
                }
                if (highest == 1)
                    break;
            }
            // updateMatchMask(pattern_o1, pattern_oN, match_o1, mask, nodes)  (ElementIntersecting.hpp:197-228)
            for (int m : match_inds)
            {
                if (mask[m])
                    continue;
                const auto& pm = my_locs[m];
                for (int p : pat_inds)
                    if (pointsMatch(locs[nbr][p], pm))
                    {
                        mask[m] = 1;
                        nn[m]   = new_nodes_of[nbr][p];
                        break;
                    }
            }
        }
        for (int i : boundaryNodeInds(el.et, OC))
            if (not mask[i])
                nn[i] = max_node++;
        for (int i : internalNodeInds(el.et, OC))
            if (not mask[i])
                nn[i] = max_node++;
        new_nodes_of[e] = std::move(nn);
        converted[e]    = 1;
    };
    // domain 0 (volume) first, then the boundary domains by ascending id, elements by ascending id within a domain
    for (std::size_t e = 0; e < mesh.n_elems; ++e)
        convert(e);
    std::vector< std::size_t > border(mesh.boundary.size());
    for (std::size_t b = 0; b < border.size(); ++b)
        border[b] = b;
    std::stable_sort(border.begin(), border.end(), [&](std::size_t a, std::size_t b) {
        return std::make_pair(mesh.boundary[a].domain_id, mesh.boundary[a].id) <
               std::make_pair(mesh.boundary[b].domain_id, mesh.boundary[b].id);
    });
    for (auto b : border)
        convert(mesh.n_elems + b);
    for (std::size_t e = 0; e < mesh.n_elems; ++e)
        std::copy(new_nodes_of[e].begin(), new_nodes_of[e].end(), out.elem_nodes.begin() + e * nvN);
    for (std::size_t b = 0; b < mesh.boundary.size(); ++b)
    {
        out.boundary[b].nodes = new_nodes_of[mesh.n_elems + b];
        if (static_cast< int >(out.boundary[b].nodes.size()) != nbN)
            throw std::logic_error{"boundary element conversion failed"};
    }
    out.n_nodes = max_node;
    return out;
}

// mesh/MeshPartition.hpp:505-596 — match each boundary element to (parent volume element, side) by sorted node sets
void matchBoundaries(Mesh& mesh)
{
    const int                                                               npe = mesh.nodesPerElem();
    std::map< std::vector< n_id_t >, std::pair< std::size_t, int > >         side_map;
    std::vector< std::vector< int > >                                        side_inds;
    for (int s = 0; s < numSides(mesh.et); ++s)
        side_inds.push_back(sideNodeInds(mesh.et, mesh.order, s));
    // only sides that can be boundaries need to be registered; register all, first match wins (ascending element id)
    for (std::size_t e = 0; e < mesh.n_elems; ++e)
        for (int s = 0; s < numSides(mesh.et); ++s)
        {
            std::vector< n_id_t > key;
            for (int i : side_inds[s])
                key.push_back(mesh.elem_nodes[e * npe + i]);
            std::sort(key.begin(), key.end());
            side_map.emplace(std::move(key), std::make_pair(e, s));
        }
    for (auto& b : mesh.boundary)
    {
        auto key = b.nodes;
        std::sort(key.begin(), key.end());
        const auto it = side_map.find(key);
        if (it == side_map.end())
            throw std::runtime_error{"BoundaryView could not be constructed: boundary element is not a side of any domain element"};
        b.parent = it->second.first;
        b.side   = it->second.second;
    }
}
} // namespace orc
