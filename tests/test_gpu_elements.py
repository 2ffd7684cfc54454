"""Patch operators from ELEMENT matrices on the GPU (SURVEY.md 8(a) row A6 by the reference's own mechanism:
PCPatchComputeOperator ssc/libssc.c:1120-1156 + the element-kernel callback ssc/patch.py:28-62, 210-213).

Cells are added in the order of the patch's cell list, so the blocks must be bit-identical to the oracle's
element-wise assembly; the preconditioner built from them must agree with the one cut out of the assembled operator
to the fp64 tolerance of BASELINE.json (1e-12 relative).
"""
import numpy as np
import pytest

from conftest import rhs_vector
from helpers import make_gpu, make_oracle, relerr
from ssc_b200 import fem, plex
from ssc_b200._lib import SSCError

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _problem(mesh, p, kind):
    form, bcs = fem.poisson_problem(mesh, p, kind)
    disc = fem.discretisation_from_form(form, bcs)
    return form, bcs, disc, form.assemble(bcs)


def _per_cell(Ke, ncells):
    return Ke if Ke.ndim == 3 else np.broadcast_to(Ke, (ncells,) + Ke.shape).copy()


def _element_pc(form, bcs, Ke, **opts):
    pc = make_gpu(form, bcs, **opts)
    pc._compute_op = None
    pc.setElementMatrices(Ke)
    return pc


@pytest.mark.parametrize("mesh,p,kind", [
    (plex.RectangleMesh(6, 5, 1, 1), 3, "scalar"), (plex.RectangleMesh(4, 5, 1, 1), 2, "vector"),
    (plex.RectangleMesh(4, 3, 1, 1), 2, "mixed"), (plex.BoxMesh(3, 3, 2), 2, "scalar"),
    (plex.TensorPlex((5, 4)), 3, "scalar"), (plex.TensorPlex((3, 3, 3)), 3, "scalar"),
    (plex.TensorPlex((3, 2, 3)), 2, "vector"), (plex.TensorPlex((2, 2, 2)), 5, "scalar"),
], ids=lambda v: v if isinstance(v, (int, str)) else type(v).__name__)
def test_blocks_bit_identical_and_apply_matches(mesh, p, kind):
    form, bcs, disc, A = _problem(mesh, p, kind)
    Ke = fem.poisson_element_matrices(form.spaces)
    o = make_oracle(disc, A)
    pc = _element_pc(form, bcs, Ke, ssc_keep_patch_operators=True)
    pc.setUp()
    off, _ = pc.getPatches()
    Kc = _per_cell(Ke, mesh.num_cells)
    for i in range(0, off.size - 1, max(1, (off.size - 1) // 25)):
        if off[i + 1] == off[i]:
            continue
        B = pc.getPatchMatrix(i)
        assert np.array_equal(B, o.patchMatrixFromElements(i, Kc))      # same additions in the same order
        assert np.allclose(B, o.patchMatrix(i), rtol=1e-12, atol=1e-13)  # == A[gtol, gtol] up to summation order
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < 1e-11          # two different summation orders of the same blocks, inverted
    pc_csr = make_gpu(form, bcs)
    pc_csr.setUp()
    y2 = np.zeros_like(x)
    pc_csr.apply(x, y2)
    assert relerr(y, y2) < 1e-11


def test_elasticity_tets_from_elements():
    mesh = plex.BoxMesh(3, 3, 3)
    V = fem.FunctionSpace(mesh, 2, 3)
    A0, Ke = fem.assemble_elasticity_simplex(V)
    bcs = [fem.DirichletBC(0, V.boundary_nodes())]
    form = fem.Form([V], lambda b: fem.apply_bcs(A0, fem.discretisation_from_form(form, b).ghost_bc_nodes))
    disc = fem.discretisation_from_form(form, bcs)
    o = make_oracle(disc, form.assemble(bcs))
    pc = _element_pc(form, bcs, Ke, ssc_keep_patch_operators=True)
    pc.setOption("sub_pc_type", "cholesky")
    pc.setUp()
    off, _ = pc.getPatches()
    big = int(np.argmax(np.diff(off)))
    assert off[big + 1] - off[big] == 45
    assert np.array_equal(pc.getPatchMatrix(big), o.patchMatrixFromElements(big, Ke))
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < 1e-11


def test_elements_without_saved_operators_and_update():
    """save_operators false re-assembles inside every apply; a second setUp re-reads the element matrices."""
    mesh = plex.TensorPlex((4, 4, 3))
    form, bcs, disc, A = _problem(mesh, 2, "scalar")
    Ke = fem.poisson_element_matrices(form.spaces)
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    ref = o.apply(x)
    pc = _element_pc(form, bcs, Ke, pc_patch_save_operators=False)
    pc.setUp()
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, ref) < 1e-11
    pc2 = _element_pc(form, bcs, Ke)
    pc2.setUp()
    pc2.setElementMatrices(3.0 * Ke)
    pc2.setUp()
    pc2.apply(x, y)
    free = np.setdiff1d(np.arange(x.size), disc.global_bc_nodes)          # Dirichlet dofs pass x through unscaled
    assert relerr(3.0 * y[free], ref[free]) < 1e-11
    assert np.array_equal(y[disc.global_bc_nodes], x[disc.global_bc_nodes])


def test_element_route_errors():
    mesh = plex.RectangleMesh(4, 4, 1, 1)
    form, bcs, disc, A = _problem(mesh, 2, "scalar")
    Ke = fem.poisson_element_matrices(form.spaces)
    pc = _element_pc(form, bcs, Ke[:, :5, :5])
    with pytest.raises(SSCError, match="dofs per cell"):
        pc.setUp()
    pc = _element_pc(form, bcs, Ke[:3])
    with pytest.raises(SSCError, match="element matrices were given"):
        pc.setUp()
    pc = _element_pc(form, bcs, Ke, pc_patch_multiplicative=True)
    with pytest.raises(SSCError, match="assembled operator"):
        pc.setUp()
    from ssc_b200 import PC
    o = make_oracle(disc)
    off, gtol = o.patches()
    pc = PC(0)
    pc.setPatches(off, gtol, disc.total_dofs, global_bc_nodes=disc.global_bc_nodes)
    pc.setElementMatrices(Ke)
    with pytest.raises(SSCError, match="built from the mesh"):
        pc.setUp()
