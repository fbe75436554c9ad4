"""BASELINE configs[3] on the REAL mesh of examples/07-karman-2D (tests/golden/karman_order1.npz, see make_karman_fixture.py): the
steady Navier-Stokes problem as source.cpp:81-155 sets it up — quad p = 4, unknowns (u, v, vorticity, p), AssemblyOptions{1, 1},
steady kernel linearised about a previous velocity held as 2 nodal fields, outlet boundary kernel on (u, v, p), Dirichlet u, v on
wall (no slip) and inlet (parabolic profile, source.cpp:167-171)."""
import os

import numpy as np

import l3ster_b200 as l3b
from l3ster_b200 import meshio

DOMAIN, INLET, WALL, OUTLET = 44, 45, 46, 47
U, P = 4, 4
IU, IV, IO, IP = 0, 1, 2, 3
OPTS = l3b.AssemblyOptions(value_order=1, derivative_order=1)
KERNELS = [dict(name="karman_steady", asm_opts=OPTS, field_inds=[0, 1]),
           dict(name="karman_outlet", boundary_ids=[OUTLET], dof_inds=[IU, IV, IP], asm_opts=OPTS)]


def load_order1():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "karman_order1.npz"))
    return meshio.Order1Mesh(int(d["dim"]), d["coords"], d["elems"], d["elem_ids"], d["elem_domains"], d["bnd_elems"], d["bnd_ids"], d["bnd_domains"])


def node_coords(host):
    """physical coordinates of the order-p nodes: the bilinear map of the element's vertices at the Gauss-Lobatto lattice"""
    gll = l3b.tables_gll(host.order + 1)
    nb = host.order + 1
    xi = (1.0 + gll[np.arange(nb * nb) % nb]) / 2.0
    eta = (1.0 + gll[np.arange(nb * nb) // nb]) / 2.0
    w = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=1)  # vertex order: lexicographic
    xy = np.zeros((host.n_nodes, 2))
    xy[host.nodes.ravel()] = np.einsum("av,evd->ead", w, host.verts[:, :, :2]).reshape(-1, 2)
    return xy


def previous_velocity(xy):
    """a smooth, divergence-free-ish stand-in for the previous Newton iterate (2 fields: u, v)"""
    x, y = xy[:, 0], xy[:, 1]
    u = 1.5 * (1.0 - y * y) * (1.0 + 0.1 * np.sin(0.8 * x))
    v = 0.05 * np.sin(np.pi * y) * np.cos(0.7 * x)
    return np.stack([u, v])


def dirichlet(nodes_wall, nodes_inlet, xy):
    """(dofs, values): u = v = 0 on the wall, u = 1.5 (1 - y^2), v = 0 at the inlet; the wall wins where both meet (set last there,
    source.cpp:165-166 sets the wall first and the inlet second: the INLET value wins)"""
    nodes = np.union1d(nodes_wall, nodes_inlet)
    u = np.zeros(len(nodes))
    at_inlet = np.isin(nodes, nodes_inlet)
    u[at_inlet] = 1.5 * (1.0 - xy[nodes[at_inlet], 1] ** 2)
    dofs = np.concatenate([nodes * U + IU, nodes * U + IV])
    vals = np.concatenate([u, np.zeros(len(nodes))])
    order = np.argsort(dofs)
    return dofs[order], vals[order]
