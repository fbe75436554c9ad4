"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/, pinned by the reference's known-answer tests — the reference
itself cannot be built or imported in this image, so these are oracle outputs frozen as regression vectors, not reference dumps).

    python tests/golden/make_golden.py

Cases: the reference's own single-element fixtures (tests/LocalOperatorCommon.hpp:17-61: distorted quad p=4, distorted hex p=3,
asm_opts{.value_order = 2}) and the benchmark element (benchmarks/Common.hpp:20-29, hex p=4, default options) with seeded operands
(numpy default_rng(5489), U[-1, 1], as tests/LocalOperatorCommon.hpp:226-227 draws them). Stored per case: K_e, F_e, x, y = K_e x by
the sum-factorised and by the local-element evaluation, diag(K_e)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402

CASES = {
    "quad_p4_diffusion2d": dict(dim=2, p=4, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0]], kernel="diffusion_kernel_2D", U=3, n_rhs=2,
                                value_order=2),
    "hex_p3_diffusion3d": dict(dim=3, p=3, verts=[[1, 1, 0], [2, 1, 0], [1, 3, 0], [3, 4, 0], [1, 1, 1], [2, 1, 1.5], [1, 3, 2], [3, 4, 3.5]],
                               kernel="diffusion_kernel_3D", U=4, n_rhs=3, value_order=2),
    "hex_p4_benchmark": dict(dim=3, p=4, verts=[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [2, 2, 2]],
                             kernel="bench_diffusion3d", U=4, n_rhs=1, value_order=1),
}


def main():
    orc = Oracle()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, c in CASES.items():
        nn = (c["p"] + 1) ** c["dim"]
        K, F = orc.assemble_local(c["kernel"], c["dim"], c["p"], c["verts"], n_rhs=c["n_rhs"], value_order=c["value_order"])
        x = np.random.default_rng(5489).uniform(-1, 1, size=(nn * c["U"], c["n_rhs"]))
        y_sf = orc.eval_sumfact(c["kernel"], c["dim"], c["p"], c["verts"], x, value_order=c["value_order"], eval_strategy=2)
        y_le = orc.eval_local_operator(c["kernel"], c["dim"], c["p"], c["verts"], x, value_order=c["value_order"])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), K=K, F=F, x=x, y_sumfact=y_sf, y_local=y_le, diag=np.diag(K).copy(),
                            verts=np.array(c["verts"], dtype=float), meta=np.array([c["dim"], c["p"], c["U"], c["n_rhs"], c["value_order"]]))
        print(name, K.shape, float(np.linalg.norm(K)), float(np.abs(K @ x - y_sf).max()))


if __name__ == "__main__":
    main()
