"""Short workload for ncu: static condensation of n^3 hex p=4 elements (bench_diffusion3d). Usage: profile_condense.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import l3ster_b200 as l3b  # noqa: E402
from l3ster_b200.condensation import CondensedAssembledSystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = np.linspace(0.0, 1.0, n + 1)
ctx = l3b.Context(0)
host = l3b.make_cube_mesh(x, order=4)
cs = CondensedAssembledSystem(ctx, host, 4)
for _ in range(2):
    cs.beginAssembly()
    cs.assembleProblem("bench_diffusion3d")
    cs.endAssembly()
ctx.synchronize()
print("condensed", host.n_elems, "elements,", cs.n_primary_dofs, "primary dofs")
