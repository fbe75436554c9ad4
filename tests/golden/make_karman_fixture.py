"""Regenerates tests/golden/karman_order1.npz: the order-1 mesh of the reference's examples/07-karman-2D/karman.msh (996 nodes, 914
quadrangles in domain 44, 164 boundary lines in the domains 45 inlet / 46 wall / 47 outlet) as parsed by l3ster_b200.meshio.read_gmsh
— mesh::readMesh's conventions (mesh/ReadMesh.hpp:107-362). The .msh itself lives only in /root/reference (absent on the GPU box);
tests/test_karman_mesh.py re-reads it whenever it is present and checks the reader still yields exactly these arrays.

    python tests/golden/make_karman_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from l3ster_b200 import meshio  # noqa: E402

SRC = "/root/reference/examples/07-karman-2D/karman.msh"
INLET, WALL, OUTLET = 45, 46, 47  # examples/07-karman-2D/source.cpp:11


def main():
    m = meshio.read_gmsh(SRC, [INLET, WALL, OUTLET])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "karman_order1.npz")
    np.savez_compressed(out, dim=m.dim, coords=m.coords, elems=m.elems, elem_ids=m.elem_ids, elem_domains=m.elem_domains, bnd_elems=m.bnd_elems,
                        bnd_ids=m.bnd_ids, bnd_domains=m.bnd_domains)
    print(out, os.path.getsize(out), "bytes:", m.coords.shape, m.elems.shape, m.bnd_elems.shape)


if __name__ == "__main__":
    main()
