"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/ssc_oracle.c).

The reference ships no golden vectors for this path (its two tests are Firedrake end-to-end solves) and
cannot be built or imported here, so these fixtures are outputs of the oracle restatement, NOT of the reference:
they pin the oracle against regressions and give the GPU tests fixed vectors to hit.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "rect_p2_scalar": dict(mesh=("RectangleMesh", (6, 5, 2.0, 3.0)), p=2, kind="scalar"),
    "box_p2_vector": dict(mesh=("BoxMesh", (3, 2, 3, 1.0, 2.0, 3.0)), p=2, kind="vector"),
    "hex_q3_scalar": dict(mesh=("TensorPlex", ((3, 3, 3),)), p=3, kind="scalar"),
    "quad_q4_mixed_pou": dict(mesh=("TensorPlex", ((3, 2),)), p=4, kind="mixed", opts=dict(partition_of_unity=1)),
}


def build_case(spec):
    from ssc_b200 import fem, plex
    ctor, args = spec["mesh"]
    mesh = getattr(plex, ctor)(*args)
    form, bcs = fem.poisson_problem(mesh, spec["p"], spec["kind"])
    disc = fem.discretisation_from_form(form, bcs)
    return form, bcs, disc, form.assemble(bcs)


def main():
    from helpers import make_oracle
    out = os.path.dirname(os.path.abspath(__file__))
    for name, spec in CASES.items():
        form, bcs, disc, A = build_case(spec)
        o = make_oracle(disc, A, **spec.get("opts", {}))
        off, gtol = o.patches()
        coff, cells = o.cells()
        x = np.random.default_rng(20260101).uniform(-1, 1, disc.total_dofs)
        y = o.apply(x)
        np.savez_compressed(os.path.join(out, name + ".npz"), gtol_off=off, gtol=gtol, cell_off=coff, cells=cells, x=x, y=y,
                            ghost_bc=disc.ghost_bc_nodes, indptr=A.indptr, indices=A.indices, data=A.data)
        print(name, disc.total_dofs, "dofs", off.size - 1, "patches", os.path.getsize(os.path.join(out, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
