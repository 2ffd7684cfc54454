"""The reference's own end-to-end test, through the same options surface, on the GPU.

tests/test_jacobi_equivalent.py:23-85 of the reference: for CG1 a vertex star owns exactly the vertex's dofs,
so CG preconditioned with ssc.PatchPC must produce the same residual history as CG + Jacobi
(numpy.allclose, the reference's tolerance).  Also: identical Krylov iteration counts GPU vs oracle for p >= 2.
"""
import numpy as np
import pytest

from conftest import rhs_vector
from helpers import OraclePythonPC, solve
from ssc_b200 import fem
from ssc_b200 import petsc_lite as pl

pytestmark = pytest.mark.gpu


def _load_vector(form, disc):
    """inner(Constant(1), v)*dx lumped: any fixed rhs works for the equivalence; BC rows hold zero."""
    b = np.ones(disc.total_dofs)
    b[disc.ghost_bc_nodes] = 0.0
    return b


def test_jacobi_equivalence(mesh, problem_type):
    form, bcs = fem.poisson_problem(mesh, 1, problem_type)
    disc = fem.discretisation_from_form(form, bcs)
    A = form.assemble(bcs)
    b = _load_vector(form, disc)
    pl.Options.clear()
    _, _, jacobi_history = solve(pl.Mat(A), b, "cg", lambda pc: pc.setType("jacobi"))

    # mat_type matfree: the operator is a python shell whose context carries the form and the BCs
    ctx = fem.ImplicitMatrixContext(form, bcs)
    Amf = pl.Mat(context=ctx, size=A.shape)
    pl.Options.insert_parameters({"ksp_type": "cg", "pc_type": "python", "pc_python_type": "ssc.PatchPC",
                                  "patch_pc_patch_multiplicative": False, "patch_pc_patch_save_operators": True,
                                  "patch_pc_patch_sub_mat_type": "aij", "patch_sub_ksp_type": "preonly",
                                  "patch_sub_pc_type": "lu"})
    _, _, patch_history = solve(Amf, b, "cg", lambda pc: pc.setFromOptions())
    pl.Options.clear()
    assert len(jacobi_history) == len(patch_history)
    assert np.allclose(jacobi_history, patch_history)


@pytest.mark.parametrize("ksp_type", ["cg", "gmres"])
@pytest.mark.parametrize("p", [2, 3])
def test_iteration_counts_match_oracle(mesh, ksp_type, p):
    form, bcs = fem.poisson_problem(mesh, p, "scalar")
    disc = fem.discretisation_from_form(form, bcs)
    A = form.assemble(bcs)
    b = rhs_vector(disc.total_dofs)
    b[disc.ghost_bc_nodes] = 0.0
    pl.Options.clear()
    oracle_pc = OraclePythonPC(disc, A)

    def use_oracle(pc):
        pc.setType("python")
        pc.setPythonContext(oracle_pc)
    xo, its_o, hist_o = solve(pl.Mat(A), b, ksp_type, use_oracle)

    dm = pl.DM(fem.SNESContext(form, bcs))           # assembled operator: form found through the DM's appctx
    pl.Options.insert_parameters({"pc_type": "python", "pc_python_type": "ssc.PatchPC"})
    xg, its_g, hist_g = solve(pl.Mat(A), b, ksp_type, lambda pc: pc.setFromOptions(), dm=dm)
    pl.Options.clear()
    assert its_g == its_o
    assert np.allclose(hist_g, hist_o, rtol=1e-8)
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)
