"""Host logic of the multi-GPU path: z-slab partition, [owned | ghost] numbering, halo lists, and the
Import / Export exchange itself over torch.distributed with the gloo backend, world_size 2 and 3 (incl. an empty rank,
tests/EmptyPartitionTest.cpp:10-40)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from l3ster_b200.slab import Halo, make_slab, split_layers


def _dists(n=3, nz=5):
    return np.linspace(0.0, 1.0, n + 1), np.linspace(0.0, 1.3, n + 2), np.linspace(-0.5, 0.7, nz + 1)


def test_split_layers_covers_everything_and_allows_empty_ranks():
    assert split_layers(5, 2) == [(0, 3), (3, 5)]
    assert split_layers(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for n, w in ((64, 8), (7, 3), (1, 1)):
        ls = split_layers(n, w)
        assert ls[0][0] == 0 and ls[-1][1] == n and all(a[1] == b[0] for a, b in zip(ls, ls[1:]))


@pytest.mark.parametrize("order,world", [(1, 2), (2, 3), (4, 2), (3, 4)])
def test_slabs_partition_the_global_lattice(order, world):
    x, y, z = _dists()
    slabs = [make_slab(x, y, z, order, r, world) for r in range(world)]
    n_global = (len(x) - 1) * order + 1, (len(y) - 1) * order + 1, (len(z) - 1) * order + 1
    # every lattice point is owned exactly once
    owned = np.concatenate([s.lattice[: s.n_owned_nodes] for s in slabs])
    keys = owned[:, 0] + n_global[0] * (owned[:, 1] + n_global[1] * owned[:, 2])
    assert len(keys) == np.prod(n_global) and len(np.unique(keys)) == len(keys)
    assert sum(s.n_elems for s in slabs) == (len(x) - 1) * (len(y) - 1) * (len(z) - 1)
    for s in slabs:
        if s.n_elems == 0:
            continue
        # ghosts are the tail of the local numbering, in the order their owner packs them
        if s.lower >= 0:
            lo = slabs[s.lower]
            assert np.array_equal(s.lattice[s.n_owned_nodes:], lo.lattice[lo.send_up_nodes])
            assert s.n_ghost_nodes == len(lo.send_up_nodes) == n_global[0] * n_global[1]
        else:
            assert s.n_ghost_nodes == 0
        # border elements (touching ghosts) come first and are exactly one layer
        touches = (s.nodes >= s.n_owned_nodes).any(axis=1)
        assert touches[: s.n_border_elems].all() and not touches[s.n_border_elems:].any()
        assert s.n_border_elems == (0 if s.lower < 0 else (len(x) - 1) * (len(y) - 1))
        # inter-slab faces are not physical boundaries; the cube's own faces still are
        bottom = s.lattice[s.nodes[:, 0].astype(np.int64), 2] == s.lattice[:, 2].min()
        assert ((s.side_boundaries[bottom, 0] == 0xFFFF) == (s.lower >= 0)).all()


def test_empty_rank_gets_an_empty_slab():
    x, y, z = _dists(nz=2)
    s = make_slab(x, y, z, 2, 3, 4)
    assert s.n_elems == 0 and s.n_local_nodes == 0 and s.lower == -1 and s.upper == -1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank, world, port, order, dpn, nz, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, y, z = _dists(nz=nz)
        s = make_slab(x, y, z, order, rank, world)
        halo = Halo(s, dpn)
        f = lambda lat, d: 1.0 + lat[:, 0] + 100.0 * lat[:, 1] + 1e4 * lat[:, 2] + 0.25 * d  # a global field, by lattice point and dof
        vals = np.stack([f(s.lattice, d) for d in range(dpn)], axis=1).ravel() if s.n_local_nodes else np.zeros(0)
        # Import: owned values travel to the ghost copies
        xv = torch.from_numpy(vals.copy())
        xv[halo.n_owned_dofs:] = -1.0
        halo.pack(xv)
        halo.import_x(xv)
        ok_import = bool(np.array_equal(xv.numpy(), vals))
        # Export: ghost contributions are added into their owners; every top-plane dof receives exactly its own value
        yv = torch.from_numpy(vals.copy())
        halo.export_y(yv)
        halo.unpack_add(yv)
        expect = vals.copy()
        if halo.n_up:
            expect[halo.up_idx.numpy()] *= 2.0
        ok_export = bool(np.array_equal(yv.numpy()[: halo.n_owned_dofs], expect[: halo.n_owned_dofs]))
        # a reduction over the owned dofs (the Krylov dots of §8(e)): allreduce of the local sums
        tot = torch.tensor([float(vals[: halo.n_owned_dofs].sum()), float(s.n_owned_nodes)], dtype=torch.float64)
        dist.all_reduce(tot)
        out[rank] = (ok_import, ok_export, tot.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz", [(2, 5), (3, 2)])
def test_halo_exchange_over_gloo(world, nz):
    order, dpn = 2, 3
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_halo_worker, args=(world, _free_port(), order, dpn, nz, out), nprocs=world, join=True)
        res = dict(out)
    x, y, z = _dists(nz=nz)
    n_nodes = ((len(x) - 1) * order + 1) * ((len(y) - 1) * order + 1) * (nz * order + 1)
    for r in range(world):
        ok_import, ok_export, tot = res[r]
        assert ok_import and ok_export, (r, res[r])
        assert tot[1] == n_nodes
    assert len({tuple(v[2]) for v in res.values()}) == 1


def test_two_dimensional_strips_cover_the_square():
    """make_slab(x, y, None, ...): y-strips of the square mesh — every global node owned once, ghosts = the line below, border
    elements first, interface sides not physical boundaries"""
    import numpy as np

    from l3ster_b200.slab import make_slab

    x, y, p = np.linspace(0, 2, 5), np.linspace(0, 1, 6), 3
    stride = (len(x) - 1) * p + 1
    for world in (1, 2, 3, 7):
        owned = []
        for r in range(world):
            s = make_slab(x, y, None, p, r, world)
            assert s.dim == 2 and s.nodes.shape[1] == (p + 1) ** 2
            key = s.lattice[:, 0] + stride * s.lattice[:, 1] if s.n_local_nodes else np.zeros(0, dtype=np.int64)
            owned.append(key[: s.n_owned_nodes])
            if s.n_elems:
                assert (s.nodes[: s.n_border_elems] >= s.n_owned_nodes).any(axis=1).all()
                assert not (s.nodes[s.n_border_elems:] >= s.n_owned_nodes).any()
                if s.lower >= 0:
                    assert (s.side_boundaries[:, 0] == 0xFFFF).all()  # no strip below the first one touches the bottom boundary
                    assert s.n_ghost_nodes == stride
                if s.upper >= 0:
                    assert len(s.send_up_nodes) == stride
        allk = np.concatenate(owned)
        assert len(allk) == len(np.unique(allk)) == stride * ((len(y) - 1) * p + 1)
