"""Partition import on the host (no GPU): the product's l3b_partition_* against the oracle's restatement of the reference's
assignNodes / reassignDisjointNodes / renumberNodes (mesh/PartitionMesh.hpp:322-440), the halo lists against each other
(comm/ImportExport.hpp:29-72), the row-complete owner graph against the global sparsity graph (algsys/SparsityGraph.hpp:25-278), and
the receive plan of the shared-row export by carrying symbolic values through it (AssembledSystem.hpp:384-389)."""
import numpy as np
import pytest

import l3ster_b200 as l3b
from l3ster_b200.partition import Partition, bisection_epart
from oracle import partition as opart


def _mesh(dim, n, order):
    x = np.linspace(0.0, 1.0, n + 1)
    return l3b.make_cube_mesh(x, order=order) if dim == 3 else l3b.make_square_mesh(x, order=order)


def _centroids(host):
    return host.verts.mean(axis=1)


def _checker_epart(host, n_parts, seed):
    """a deliberately ragged (non-slab, non-contiguous) partition"""
    rng = np.random.default_rng(seed)
    ep = bisection_epart(_centroids(host), n_parts)
    flip = rng.random(host.n_elems) < 0.25
    ep[flip] = rng.integers(0, n_parts, size=int(flip.sum()))
    for p in range(n_parts):  # every part keeps at least one element
        if not (ep == p).any():
            ep[p] = p
    return ep.astype(np.int32)


CASES = [(2, 4, 2, 3, "bisect"), (2, 5, 3, 4, "ragged"), (3, 3, 2, 4, "bisect"), (3, 3, 2, 5, "ragged"), (3, 2, 4, 2, "bisect")]


def _setup(dim, n, order, n_parts, kind, with_npart=False, seed=0):
    host = _mesh(dim, n, order)
    ep = bisection_epart(_centroids(host), n_parts) if kind == "bisect" else _checker_epart(host, n_parts, seed)
    npart = None
    if with_npart:  # an external nodal partition with nodes placed where no element of the part holds them ("disjoint", METIS does that)
        npart = opart.default_npart(host.nodes, ep, host.n_nodes)
        rng = np.random.default_rng(seed + 1)
        cand = rng.choice(host.n_nodes, size=max(3, host.n_nodes // 10), replace=False)
        npart[cand] = rng.integers(0, n_parts, size=len(cand))
    return host, ep, npart


@pytest.mark.parametrize("dim,n,order,n_parts,kind", CASES)
@pytest.mark.parametrize("with_npart", [False, True])
def test_node_assignment_and_renumbering_are_bit_exact(dim, n, order, n_parts, kind, with_npart):
    host, ep, npart = _setup(dim, n, order, n_parts, kind, with_npart)
    part = Partition.from_host_mesh(host, n_parts, ep, npart)
    np_in = opart.default_npart(host.nodes, ep, host.n_nodes) if npart is None else npart
    new_id, np_fixed, owned, ghost_new = opart.assign_and_renumber(host.nodes, ep, np_in, n_parts)
    assert np.array_equal(part.new_id, new_id)
    assert np.array_equal(part.npart, np_fixed)
    assert part.dist.tolist() == [0] + list(np.cumsum([len(o) for o in owned]))
    # the reference's invariants: a permutation; every rank owns a contiguous range; ghosts ascending by global id
    assert sorted(new_id.tolist()) == list(range(host.n_nodes))
    for r in range(n_parts):
        v = part.rank_view(r)
        assert v.n_owned_nodes == len(owned[r]) and v.first_gid == part.dist[r]
        assert v.gids[:v.n_owned_nodes].tolist() == list(range(int(part.dist[r]), int(part.dist[r + 1])))
        assert v.gids[v.n_owned_nodes:].tolist() == ghost_new[r]
        # local connectivity maps back to the renumbered global one; border elements first
        assert np.array_equal(v.gids[v.nodes.astype(np.int64)], new_id[host.nodes[v.elem_ids].astype(np.int64)])
        assert sorted(v.elem_ids.tolist()) == np.nonzero(ep == r)[0].tolist()
        touches = (v.nodes >= v.n_owned_nodes).any(axis=1)
        assert touches[:v.n_border_elems].all() and not touches[v.n_border_elems:].any()


@pytest.mark.parametrize("dim,n,order,n_parts,kind", CASES)
@pytest.mark.parametrize("extended", [False, True])
def test_halo_lists_pair_up(dim, n, order, n_parts, kind, extended):
    """what rank r packs for q (its owned nodes, in order) is exactly q's ghost range owned by r — comm::ImportExportContext both ways"""
    host, ep, _ = _setup(dim, n, order, n_parts, kind)
    part = Partition.from_host_mesh(host, n_parts, ep)
    views = [part.rank_view(r, extended) for r in range(n_parts)]
    for r, v in enumerate(views):
        off = 0
        for q, o, size in v.shared_halo:  # ghost ranges tile the ghost block, owners ascending
            assert o == off and size > 0
            off += size
            gids = v.gids[v.n_owned_nodes + o:v.n_owned_nodes + o + size]
            assert (gids >= part.dist[q]).all() and (gids < part.dist[q + 1]).all()
            mine_there = [nodes for rr, nodes in views[q].owned_halo if rr == r]
            assert len(mine_there) == 1
            assert np.array_equal(views[q].gids[mine_there[0]], gids)
        assert off == v.n_local_nodes - v.n_owned_nodes
        for q, nodes in v.owned_halo:
            assert (np.asarray(nodes) < v.n_owned_nodes).all()
            assert any(rr == r for rr, _, _ in views[q].shared_halo)
    if extended:  # the extended view only adds ghosts
        for r in range(n_parts):
            plain = part.rank_view(r, False)
            assert set(plain.gids.tolist()) <= set(views[r].gids.tolist())


@pytest.mark.parametrize("dim,n,order,n_parts,kind", CASES)
def test_owner_rows_are_the_global_rows(dim, n, order, n_parts, kind):
    """algsys/SparsityGraph.hpp:83-278: after the neighbour-row exchange every owned row holds the columns of ALL elements around the
    node; columns are the extended local ids in ascending order (owned first, then the rest by global id)"""
    host, ep, _ = _setup(dim, n, order, n_parts, kind)
    part = Partition.from_host_mesh(host, n_parts, ep)
    rows = opart.global_rows(part.new_id[host.nodes.astype(np.int64)])
    for r in range(n_parts):
        v = part.rank_view(r, True)
        (ptr, nbr), _ = part.rank_graph(r)
        assert len(ptr) == v.n_local_nodes + 1
        local_rows = opart.global_rows(v.gids[v.nodes.astype(np.int64)]) if v.n_elems else {}
        for l in range(v.n_local_nodes):
            cols = nbr[ptr[l]:ptr[l + 1]].astype(np.int64)
            assert (np.diff(cols) > 0).all()
            gcols = sorted(v.gids[cols].tolist())
            gid = int(v.gids[l])
            if l < v.n_owned_nodes:
                assert gcols == rows[gid]
            else:
                assert gcols == local_rows.get(gid, [])  # ghost rows: my elements only; extra columns: empty rows


@pytest.mark.parametrize("dim,n,order,n_parts,kind", CASES[:4])
def test_shared_row_export_plan_carries_values_to_the_right_entries(dim, n, order, n_parts, kind):
    """every rank 'assembles' symbolic element contributions into its rows (device layout: node blocks of dpn rows, column-dof-major),
    ghost-row slices travel to the owners and are added through the receive plan with the index arithmetic of rowExportAddKernel;
    the owned rows must then equal the globally assembled ones"""
    dpn = 2
    host, ep, _ = _setup(dim, n, order, n_parts, kind)
    part = Partition.from_host_mesh(host, n_parts, ep)
    nn = host.nodes.shape[1]

    def contrib(e, a, b, d, v):  # symbolic K_e entry
        return 1.0 + 0.001 * e + 0.37 * a + 0.11 * b + 0.05 * d + 0.013 * v

    views = [part.rank_view(r, True) for r in range(n_parts)]
    graphs, plans, vals = [], [], []
    for r, v in enumerate(views):
        (ptr, nbr), (eptr, pos) = part.rank_graph(r)
        graphs.append((ptr, nbr))
        plans.append((eptr, pos))
        val = np.zeros(int(ptr[-1]) * dpn * dpn)
        for le, e in enumerate(v.elem_ids):
            for a in range(nn):
                na = int(v.nodes[le, a])
                beg, deg = int(ptr[na]), int(ptr[na + 1] - ptr[na])
                row_cols = nbr[beg:beg + deg]
                for b in range(nn):
                    k = int(np.searchsorted(row_cols, v.nodes[le, b]))
                    assert row_cols[k] == v.nodes[le, b]
                    for d in range(dpn):
                        for u in range(dpn):
                            val[dpn * dpn * beg + (d * dpn + u) * deg + k] += contrib(int(e), a, b, d, u)
        vals.append(val)
    # export
    sent = {}
    for s, v in enumerate(views):
        ptr = graphs[s][0]
        for q, off, size in v.shared_halo:
            n0, n1 = v.n_owned_nodes + off, v.n_owned_nodes + off + size
            sent[(s, q)] = vals[s][dpn * dpn * ptr[n0]:dpn * dpn * ptr[n1]].copy()
    for r, v in enumerate(views):
        ptr, nbr = graphs[r]
        eptr, pos = plans[r]
        j = 0
        for q, nodes in v.owned_halo:
            buf = sent[(q, r)]
            base = int(eptr[j])
            assert len(buf) == dpn * dpn * int(eptr[j + len(nodes)] - eptr[j])
            for node in nodes:
                deg_s, deg_r = int(eptr[j + 1] - eptr[j]), int(ptr[node + 1] - ptr[node])
                src0, dst0 = dpn * dpn * (int(eptr[j]) - base), dpn * dpn * int(ptr[node])
                for k in range(deg_s):
                    for dv in range(dpn * dpn):
                        vals[r][dst0 + dv * deg_r + int(pos[eptr[j] + k])] += buf[src0 + dv * deg_s + k]
                j += 1
        assert j + 1 == len(eptr)
    # expected: global assembly in the renumbered ids
    glob = {}
    gnodes = part.new_id[host.nodes.astype(np.int64)]
    for e in range(host.n_elems):
        for a in range(nn):
            for b in range(nn):
                for d in range(dpn):
                    for u in range(dpn):
                        key = (int(gnodes[e, a]), int(gnodes[e, b]), d, u)
                        glob[key] = glob.get(key, 0.0) + contrib(e, a, b, d, u)
    for r, v in enumerate(views):
        ptr, nbr = graphs[r]
        for l in range(v.n_owned_nodes):
            beg, deg = int(ptr[l]), int(ptr[l + 1] - ptr[l])
            for k in range(deg):
                gc = int(v.gids[nbr[beg + k]])
                for d in range(dpn):
                    for u in range(dpn):
                        got = vals[r][dpn * dpn * beg + (d * dpn + u) * deg + k]
                        assert abs(got - glob[(int(v.gids[l]), gc, d, u)]) < 1e-12


def test_partition_rejects_bad_input():
    host = _mesh(2, 2, 1)
    with pytest.raises(l3b.L3BError):
        Partition.from_host_mesh(host, 2, np.array([0, 1, 2, 0], dtype=np.int32))  # epart entry out of range
    with pytest.raises(l3b.L3BError):
        Partition.from_host_mesh(host, 2, np.array([0, 1, 1, 0], dtype=np.int32), np.full(host.n_nodes, 5, dtype=np.int32))
