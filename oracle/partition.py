"""TEST INFRASTRUCTURE (oracle): numpy restatement of the deterministic half of the reference's mesh partitioner and of the
distributed set-up that follows it, for a GIVEN METIS output (epart, npart). Only tests may import this.

Follows, line by line in spirit:
  assignNodes              mesh/PartitionMesh.hpp:354-395
  reassignDisjointNodes    mesh/PartitionMesh.hpp:322-351
  renumberNodes            mesh/PartitionMesh.hpp:411-440
  local numbering          dofs/NodeToDofMap.hpp:144-163  ([owned | shared sorted by global id])
  ImportExportContext      comm/ImportExport.hpp:29-72, 250-261
  column ownership         algsys/SparsityGraph.hpp:221-233
Parity is pinned by construction checks in tests/test_partition.py (every invariant the reference asserts: contiguous ownership, sorted
ghosts, each node owned exactly once) — the reference's own tests for this path (tests/MeshUtilsTests.cpp partitioning cases) need METIS
and are therefore *unpinned* here beyond those invariants."""
import numpy as np


def assign_and_renumber(elem_nodes, epart, npart, n_parts):
    """returns (new_id, npart_repaired, owned[part] (old ids, ascending), ghost[part] (NEW ids, ascending))"""
    elem_nodes = np.asarray(elem_nodes, dtype=np.int64)
    npart = np.array(npart, dtype=np.int64)
    n_nodes = int(elem_nodes.max()) + 1
    owned, ghost = [], []
    for part in range(n_parts):  # assignNodes
        o, g = set(), set()
        for e in np.nonzero(epart == part)[0]:
            for n in elem_nodes[e]:
                (o if npart[n] == part else g).add(int(n))
        owned.append(sorted(o))
        ghost.append(sorted(g))
    disjoint = []
    for part in range(n_parts):
        oset = set(owned[part])
        for n in range(n_nodes):
            if npart[n] == part and n not in oset:
                disjoint.append(n)
    for part in range(n_parts):  # reassignDisjointNodes
        claimed = [n for n in disjoint if n in set(ghost[part])]
        if claimed:
            disjoint = [n for n in disjoint if n not in set(claimed)]
            ghost[part] = sorted(set(ghost[part]) - set(claimed))
            owned[part] = sorted(owned[part] + claimed)
            npart[claimed] = part
    assert not disjoint, "At least one node in the mesh does not belong to any element"
    new_id = np.full(n_nodes, -1, dtype=np.int64)  # renumberNodes
    nxt = 0
    for part in range(n_parts):
        for n in owned[part]:
            new_id[n] = nxt
            nxt += 1
    ghost_new = [sorted(int(new_id[n]) for n in g) for g in ghost]
    return new_id, npart, owned, ghost_new


def default_npart(elem_nodes, epart, n_nodes):
    """the rule the product applies when no npart is supplied: lowest part among the elements holding the node"""
    out = np.full(n_nodes, np.iinfo(np.int64).max, dtype=np.int64)
    for e, nodes in enumerate(np.asarray(elem_nodes, dtype=np.int64)):
        out[nodes] = np.minimum(out[nodes], epart[e])
    return out


def global_rows(elem_nodes_new):
    """node-level rows of the global sparsity graph (algsys/SparsityGraph.hpp:25-81): row n = sorted union of the nodes of all elements
    containing n"""
    rows = {}
    for nodes in np.asarray(elem_nodes_new, dtype=np.int64):
        s = set(int(n) for n in nodes)
        for n in s:
            rows.setdefault(n, set()).update(s)
    return {n: sorted(v) for n, v in rows.items()}
