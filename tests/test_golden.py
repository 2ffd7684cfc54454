"""Committed fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle).

CPU: the oracle and the product's host builder still reproduce them.  GPU: PCApply through the C ABI hits the stored
vectors to 1e-12 and the stored dof lists bit-exactly -- both from the full pipeline and when the stored dof lists and
CSR are handed straight to SSCPatchSetPatches / SSCPatchSetOperatorCSR (the "caller already has gtol" entry)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

from helpers import make_gpu, make_oracle, relerr  # noqa: E402

CASES = sorted(make_golden.CASES)


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_and_builder_reproduce_golden(name):
    from ssc_b200 import PC
    from ssc_b200.patch import setup_patch_pc
    g = load(name)
    spec = make_golden.CASES[name]
    form, bcs, disc, A = make_golden.build_case(spec)
    assert np.array_equal(A.indptr, g["indptr"]) and np.array_equal(A.indices, g["indices"])
    assert np.allclose(A.data, g["data"], rtol=1e-13, atol=1e-15)
    o = make_oracle(disc, A, **spec.get("opts", {}))
    off, gtol = o.patches()
    assert np.array_equal(off, g["gtol_off"]) and np.array_equal(gtol, g["gtol"])
    coff, cells = o.cells()
    assert np.array_equal(coff, g["cell_off"]) and np.array_equal(cells, g["cells"])
    assert relerr(o.apply(g["x"]), g["y"]) < 1e-13
    pc = PC(device=-1)
    setup_patch_pc(pc, form, bcs)
    off2, gtol2 = pc.getPatches()
    assert np.array_equal(off2, g["gtol_off"]) and np.array_equal(gtol2, g["gtol"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_hits_golden(name):
    from ssc_b200 import PC
    g = load(name)
    spec = make_golden.CASES[name]
    form, bcs, disc, A = make_golden.build_case(spec)
    pc = make_gpu(form, bcs, **spec.get("opts", {}))
    pc.setUp()
    off, gtol = pc.getPatches()
    assert np.array_equal(off, g["gtol_off"]) and np.array_equal(gtol, g["gtol"])
    y = np.zeros_like(g["x"])
    pc.apply(g["x"], y)
    assert relerr(y, g["y"]) < 1e-12
    # inner drop-in level: hand the stored dof lists and CSR to the library directly
    pc2 = PC(device=0)
    pc2.setPatches(g["gtol_off"], g["gtol"], g["x"].size, global_bc_nodes=g["ghost_bc"])
    for k, v in spec.get("opts", {}).items():
        pc2.setOption("pc_patch_" + k, v)
    pc2.setOperatorCSR(g["indptr"], g["indices"], g["data"])
    pc2.setUp()
    y2 = np.zeros_like(y)
    pc2.apply(g["x"], y2)
    assert relerr(y2, g["y"]) < 1e-12
