"""Golden vectors (tests/golden/*.npz, frozen oracle outputs on the reference's single-element fixtures, generator:
tests/golden/make_golden.py): the oracle must keep reproducing them on the CPU, the CUDA path must match them on the GPU."""
import glob
import os

import numpy as np
import pytest

import l3ster_b200 as l3b
from common import oracle, rel_err

KERNELS = {"quad_p4_diffusion2d": "diffusion_kernel_2D", "hex_p3_diffusion3d": "diffusion_kernel_3D", "hex_p4_benchmark": "bench_diffusion3d"}
# the element-level vectors (tests/golden/karman_order1.npz is a mesh fixture with its own tests: test_karman_mesh.py)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")) if os.path.basename(p)[:-4] in KERNELS)
TOL = 1e-12


def _load(path):
    g = np.load(path)
    dim, p, U, n_rhs, vo = (int(v) for v in g["meta"])
    return g, dim, p, U, n_rhs, vo, KERNELS[os.path.basename(path)[:-4]]


def test_golden_files_are_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_golden(path):
    g, dim, p, U, n_rhs, vo, kernel = _load(path)
    orc = oracle()
    K, F = orc.assemble_local(kernel, dim, p, g["verts"], n_rhs=n_rhs, value_order=vo)
    assert rel_err(K, g["K"]) < 1e-14 and np.abs(F - g["F"]).max() <= 1e-14 * max(1.0, np.abs(g["F"]).max())
    assert rel_err(orc.eval_sumfact(kernel, dim, p, g["verts"], g["x"], value_order=vo, eval_strategy=2), g["y_sumfact"]) < 1e-14
    assert rel_err(orc.eval_local_operator(kernel, dim, p, g["verts"], g["x"], value_order=vo), g["y_local"]) < 1e-14
    # self-consistency of the frozen vectors: both evaluations are K_e x
    assert rel_err(g["y_sumfact"], g["K"] @ g["x"]) < TOL and rel_err(g["y_local"], g["K"] @ g["x"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_path_matches_golden(path):
    g, dim, p, U, n_rhs, vo, kernel = _load(path)
    ctx = l3b.Context(0)
    nn = (p + 1) ** dim
    mesh = l3b.Mesh(ctx, dim, p, g["verts"][None], np.arange(nn, dtype=np.uint32)[None, :], None, nn, nn)
    opts = l3b.AssemblyOptions(value_order=vo)
    a = l3b.AssembledSystem(ctx, mesh, U, n_rhs)
    a.beginAssembly()
    a.assembleProblem(kernel, asm_opts=opts)
    vals, rhs = a.download()
    assert rel_err(vals.reshape(nn * U, nn * U), g["K"]) < TOL
    assert np.abs(rhs - g["F"]).max() <= TOL * max(1.0, np.abs(g["F"]).max())
    for strategy, key in ((2, "y_sumfact"), (1, "y_local")):
        s = l3b.MatrixFreeSystem(ctx, mesh, U, n_rhs)
        s.assembleProblem(kernel, asm_opts=l3b.AssemblyOptions(vo, 0, strategy))
        s.endAssembly()
        assert rel_err(s.apply(g["x"]), g[key]) < TOL
        diag, rhs_mf = s.download()
        assert rel_err(diag, g["diag"]) < TOL
