"""Mesh front end for unstructured meshes: the reference's Gmsh reader and its conversion of an order-1 mesh to order p
(mesh/ReadMesh.hpp:107-362, mesh/ConvertMeshToOrder.hpp:52-104, mesh/MeshPartition.hpp:505-596), producing what `Context.upload_mesh`
takes. Host-side index work only — no arithmetic of the hot path happens here.

Node numbering of the order-p mesh. The reference converts the elements domain by domain (ascending id), by ascending element id,
and lets each element number the nodes it is the first to hold — element-boundary nodes first, then interior ones, in ascending local
index; vertices keep their order-1 ids. It recognises an already numbered node by matching physical locations against converted
dual-graph neighbours (ElementIntersecting.hpp:103-228). On a conforming mesh that is the same as identifying a node by the mesh entity
it sits on — the edge (two vertex ids) or face (four vertex ids) plus its position counted from the lowest-numbered vertex — which is
what is done here, with a dictionary instead of geometry."""
from dataclasses import dataclass, field

import numpy as np

from . import node_graph, side_node_inds

NO_BOUNDARY = 0xFFFF
_GMSH_TYPES = {1: 1, 3: 2, 5: 3}  # gmsh element type -> dimension (line, quad, hex: ReadMesh.hpp:38-39)


@dataclass
class Order1Mesh:
    """what mesh::readMesh returns, flattened: nodes numbered in file order, elements with ids in file order, vertex lists in the
    reference's lexicographic order (x fastest)"""
    dim: int
    coords: np.ndarray                      # (n_nodes, 3)
    elems: np.ndarray                       # (n_elems, 2^dim) node ids of the volume elements
    elem_ids: np.ndarray
    elem_domains: np.ndarray
    bnd_elems: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), dtype=np.int64))   # (n_bnd, 2^(dim-1))
    bnd_ids: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    bnd_domains: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))


def _to_lexicographic(dim, nodes, coords):
    """detail::makeElementFromNodeData (ReadMesh.hpp:53-103): gmsh's counter-clockwise order to the reference's, upside-down planar
    quads flipped so that the Jacobian is positive"""
    nodes = list(nodes)
    if dim == 2:
        nodes[2], nodes[3] = nodes[3], nodes[2]
        c = coords[nodes]
        if np.all(np.abs(c[1:, 2] - c[0, 2]) < 1e-9):
            u, v = c[1] - c[0], c[2] - c[0]
            if u[0] * v[1] - u[1] * v[0] < 0.0:
                nodes[1], nodes[2] = nodes[2], nodes[1]
    elif dim == 3:
        nodes[2], nodes[3] = nodes[3], nodes[2]
        nodes[6], nodes[7] = nodes[7], nodes[6]
    return nodes


def read_gmsh(path, boundary_ids):
    """mesh::readMesh(path, boundary_ids, gmsh_tag): ASCII .msh 4.x with lines, quadrangles and hexahedra; the physical tag of an
    element's entity is its domain id; `boundary_ids` says which domains are boundaries (all others hold the volume elements)"""
    tok = open(path).read().split()
    pos = 0

    def section(name):
        nonlocal pos
        try:
            pos = tok.index(name, 0) + 1
        except ValueError:
            raise ValueError("Error while reading .msh file: required section not present") from None

    section("$MeshFormat")
    version, binary = float(tok[pos]), int(tok[pos + 1])
    if not (4.0 <= version < 5.0) or binary:
        raise ValueError("Unsupported .msh format. Only the ASCII v4 gmsh format is currently supported")
    section("$Entities")
    counts = [int(t) for t in tok[pos:pos + 4]]
    pos += 4
    physical = [dict() for _ in range(4)]
    for dim in range(4):
        for _ in range(counts[dim]):
            tag = int(tok[pos])
            pos += 1 + (6 if dim > 0 else 3)
            n_phys = int(tok[pos])
            pos += 1
            if n_phys > 1:
                raise ValueError("Error while reading .msh file: entity has more than 1 physical tag")
            if n_phys == 1:
                physical[dim][tag] = int(tok[pos])
                pos += 1
            if dim > 0:
                pos += 1 + int(tok[pos])
    section("$Nodes")
    n_blocks, n_nodes = int(tok[pos]), int(tok[pos + 1])
    pos += 4
    node_of, coords = {}, np.zeros((n_nodes, 3))
    for _ in range(n_blocks):
        parametric, size = int(tok[pos + 2]), int(tok[pos + 3])
        if parametric:
            raise ValueError("Encountered parametric node while reading .msh file")
        pos += 4
        tags = [int(t) for t in tok[pos:pos + size]]
        pos += size
        for t in tags:
            node_of[t] = len(node_of)
            coords[node_of[t]] = [float(x) for x in tok[pos:pos + 3]]
            pos += 3
    section("$Elements")
    n_blocks = int(tok[pos])
    pos += 4
    per_dim = {1: [], 2: [], 3: []}
    domain_dim, el_id = {}, 0
    for _ in range(n_blocks):
        ent_dim, ent_tag, etype, size = (int(t) for t in tok[pos:pos + 4])
        pos += 4
        if etype not in _GMSH_TYPES:
            raise ValueError("Error while reading .msh file: Encountered unsupported element type")
        if ent_tag not in physical[ent_dim]:
            raise KeyError(f"entity ({ent_dim}, {ent_tag}) carries elements but no physical tag")
        dom, d = physical[ent_dim][ent_tag], _GMSH_TYPES[etype]
        if domain_dim.setdefault(dom, d) != d:
            raise ValueError("Domain contains elements of different dimensions")
        for _ in range(size):
            nodes = [node_of[int(t)] for t in tok[pos + 1:pos + 1 + 2**d]]
            pos += 1 + 2**d
            per_dim[d].append((el_id, dom, _to_lexicographic(d, nodes, coords)))
            el_id += 1
    bset = set(int(b) for b in boundary_ids)
    vol_dims = {d for dom, d in domain_dim.items() if dom not in bset}
    if len(vol_dims) != 1:
        raise ValueError("the domains that are not boundaries must hold elements of one dimension")
    dim = vol_dims.pop()
    vol = [e for e in per_dim[dim] if e[1] not in bset]
    bnd = [e for e in per_dim.get(dim - 1, []) if e[1] in bset]
    as_arr = lambda rows, w: np.array([r[2] for r in rows], dtype=np.int64).reshape(len(rows), w)
    return Order1Mesh(dim, coords, as_arr(vol, 2**dim), np.array([e[0] for e in vol], dtype=np.int64), np.array([e[1] for e in vol], dtype=np.int64),
                      as_arr(bnd, 2**(dim - 1)), np.array([e[0] for e in bnd], dtype=np.int64), np.array([e[1] for e in bnd], dtype=np.int64))


def write_gmsh(path, mesh: Order1Mesh):
    """ASCII .msh 4.1 of an order-1 mesh, one entity per domain (the inverse of read_gmsh, for tests and for exporting generated meshes)"""
    def gmsh_order(d, nodes):
        n = list(nodes)
        if d == 2:
            n[2], n[3] = n[3], n[2]
        elif d == 3:
            n[2], n[3] = n[3], n[2]
            n[6], n[7] = n[7], n[6]
        return n

    blocks = []  # (element id, dim, domain, nodes)
    for i in range(len(mesh.elems)):
        blocks.append((int(mesh.elem_ids[i]), mesh.dim, int(mesh.elem_domains[i]), gmsh_order(mesh.dim, mesh.elems[i])))
    for i in range(len(mesh.bnd_elems)):
        blocks.append((int(mesh.bnd_ids[i]), mesh.dim - 1, int(mesh.bnd_domains[i]), gmsh_order(mesh.dim - 1, mesh.bnd_elems[i])))
    blocks.sort()
    domains = sorted({(b[1], b[2]) for b in blocks})
    ent_tag = {dd: i + 1 for i, dd in enumerate(domains)}
    lo, hi = mesh.coords.min(axis=0), mesh.coords.max(axis=0)
    out = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$Entities"]
    out.append(" ".join(str(sum(1 for d, _ in domains if d == k)) for k in range(4)))
    for k in range(1, 4):
        for d, dom in domains:
            if d == k:
                out.append(f"{ent_tag[(d, dom)]} {lo[0]} {lo[1]} {lo[2]} {hi[0]} {hi[1]} {hi[2]} 1 {dom} 0")
    out += ["$EndEntities", "$Nodes", f"1 {len(mesh.coords)} 1 {len(mesh.coords)}", f"{mesh.dim} 1 0 {len(mesh.coords)}"]
    out += [str(i + 1) for i in range(len(mesh.coords))]
    out += [" ".join(repr(float(x)) for x in c) for c in mesh.coords]
    out += ["$EndNodes", "$Elements"]
    runs = []  # consecutive elements of one domain form a block: element ids are file order (ReadMesh.hpp:290-315)
    for b in blocks:
        if runs and runs[-1][0] == (b[1], b[2]):
            runs[-1][1].append(b)
        else:
            runs.append(((b[1], b[2]), [b]))
    out.append(f"{len(runs)} {len(blocks)} 1 {len(blocks)}")
    gtype = {1: 1, 2: 3, 3: 5}
    for (d, dom), els in runs:
        out.append(f"{d} {ent_tag[(d, dom)]} {gtype[d]} {len(els)}")
        out += [" ".join([str(b[0] + 1)] + [str(int(n) + 1) for n in b[3]]) for b in els]
    out.append("$EndElements")
    open(path, "w").write("\n".join(out) + "\n")


def _local_multi_index(dim, order):
    nb = order + 1
    a = np.arange(nb**dim)
    return np.stack([(a // nb**d) % nb for d in range(dim)], axis=1)


def _node_keys(dim, order, vertex_ids):
    """identity of each order-p node of an element by the mesh entity it sits on (see the module docstring); interior nodes: None"""
    keys = []
    for idx in _local_multi_index(dim, order):
        free = [d for d in range(dim) if 0 < idx[d] < order]
        corner = lambda bits: int(vertex_ids[sum(((idx[d] == order) if d not in free else bits[free.index(d)]) << d for d in range(dim))])
        if not free:
            keys.append(("v", corner(())))
        elif len(free) == 1:
            a, b, i = corner((0,)), corner((1,)), int(idx[free[0]])
            keys.append(("e", a, b, i) if a < b else ("e", b, a, order - i))
        elif len(free) == 2:
            ids = {(s, t): corner((s, t)) for s in (0, 1) for t in (0, 1)}
            (s0, t0) = min(ids, key=ids.get)
            i = int(idx[free[0]]) if s0 == 0 else order - int(idx[free[0]])
            j = int(idx[free[1]]) if t0 == 0 else order - int(idx[free[1]])
            first_axis_is_d1 = ids[(1 - s0, t0)] < ids[(s0, 1 - t0)]
            keys.append(("f", *sorted(ids.values()), i, j) if first_axis_is_d1 else ("f", *sorted(ids.values()), j, i))
        else:
            keys.append(None)
    return keys


class UnstructuredHostMesh:
    """Order-p mesh in the layout of HostMesh (what Context.upload_mesh, node_graph and the systems take)"""

    def __init__(self, dim, order, verts, nodes, side_boundaries, n_nodes, elem_ids, elem_domains, bnd_nodes, bnd_domains):
        self.dim, self.order, self.n_nodes, self.n_elems = dim, order, int(n_nodes), len(nodes)
        self.nodes_per_elem, self.n_sides = (order + 1) ** dim, 2 * dim
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        self.side_boundaries = np.ascontiguousarray(side_boundaries, dtype=np.uint16)
        self.elem_ids, self.elem_domains, self.bnd_nodes, self.bnd_domains = elem_ids, elem_domains, bnd_nodes, bnd_domains

    def node_graph(self):
        return node_graph(self.n_nodes, self.nodes)

    def boundary_nodes(self, boundary_ids):
        sel = np.zeros(self.n_nodes, dtype=bool)
        for side in range(self.n_sides):
            on = np.isin(self.side_boundaries[:, side], list(boundary_ids))
            if on.any():
                sel[self.nodes[on][:, side_node_inds(self.dim, self.order, side)].ravel()] = True
        return np.nonzero(sel)[0]


def convert_to_order(mesh: Order1Mesh, order) -> UnstructuredHostMesh:
    """mesh::convertMeshToOrder< order > + the boundary matching of MeshPartition (:505-596)"""
    dim, nbd = mesh.dim, mesh.dim - 1
    work = [(int(mesh.elem_domains[i]), int(mesh.elem_ids[i]), dim, i) for i in range(len(mesh.elems))]
    work += [(int(mesh.bnd_domains[i]), int(mesh.bnd_ids[i]), nbd, i) for i in range(len(mesh.bnd_elems))]
    work.sort()
    number, next_id = {}, len(mesh.coords)
    vol_nodes = np.zeros((len(mesh.elems), (order + 1) ** dim), dtype=np.int64)
    bnd_nodes = np.zeros((len(mesh.bnd_elems), (order + 1) ** max(nbd, 0)), dtype=np.int64)
    splits = {}
    for d in {dim, nbd}:
        mi = _local_multi_index(d, order)
        on_bnd = ((mi == 0) | (mi == order)).any(axis=1)
        splits[d] = list(np.flatnonzero(on_bnd)) + list(np.flatnonzero(~on_bnd))  # boundary nodes first, then interior, ascending
    for _, _, d, i in work:
        verts = mesh.elems[i] if d == dim else mesh.bnd_elems[i]
        keys = _node_keys(d, order, verts)
        out = vol_nodes[i] if d == dim else bnd_nodes[i]
        for a in splits[d]:
            k = keys[a]
            if k is not None and k[0] == "v":
                out[a] = k[1]
            elif k is not None and k in number:
                out[a] = number[k]
            else:
                out[a] = next_id
                if k is not None:
                    number[k] = next_id
                next_id += 1
    # sides: the boundary element whose vertices are those of the side (MeshPartition.hpp:505-596)
    sides = np.full((len(mesh.elems), 2 * dim), NO_BOUNDARY, dtype=np.uint16)
    by_verts = {tuple(sorted(int(n) for n in mesh.bnd_elems[i])): int(mesh.bnd_domains[i]) for i in range(len(mesh.bnd_elems))}
    corner_of_side = [[int(a) for a in side_node_inds(dim, 1, s)] for s in range(2 * dim)]
    for e in range(len(mesh.elems)):
        for s in range(2 * dim):
            dom = by_verts.get(tuple(sorted(int(mesh.elems[e][c]) for c in corner_of_side[s])))
            if dom is not None:
                sides[e, s] = dom
    verts = np.zeros((len(mesh.elems), 2**dim, 3))
    verts[:] = mesh.coords[mesh.elems]
    return UnstructuredHostMesh(dim, order, verts, vol_nodes, sides, next_id, mesh.elem_ids, mesh.elem_domains, bnd_nodes, mesh.bnd_domains)


def order1_from_host(host, domain_id=0) -> Order1Mesh:
    """the order-1 skeleton of a structured HostMesh (vertices = its first 2^D ... nodes by id), for round trips through write_gmsh"""
    if host.order != 1:
        raise ValueError("pass the order-1 mesh")
    coords = np.zeros((host.n_nodes, 3))
    coords[host.nodes.ravel()] = host.verts.reshape(-1, 3)
    bnd, bdom = [], []
    for e in range(host.n_elems):
        for s in range(host.n_sides):
            if host.side_boundaries[e, s] != NO_BOUNDARY:
                bnd.append([int(host.nodes[e][a]) for a in side_node_inds(host.dim, 1, s)])
                bdom.append(int(host.side_boundaries[e, s]))
    n_e = host.n_elems
    return Order1Mesh(host.dim, coords, host.nodes.astype(np.int64), np.arange(n_e), np.full(n_e, domain_id),
                      np.array(bnd, dtype=np.int64).reshape(len(bnd), 2 ** (host.dim - 1)), n_e + np.arange(len(bnd)), np.array(bdom, dtype=np.int64))
