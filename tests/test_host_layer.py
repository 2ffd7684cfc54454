"""Host-side pieces that need no GPU: the stand-in Krylov drivers against scipy, the options database, and the places
where ssc.PatchPC looks for the form (ssc/patch.py:232-253)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from ssc_b200 import fem, plex
from ssc_b200 import petsc_lite as pl


def _spd(n=60, seed=0):
    rng = np.random.default_rng(seed)
    R = sp.random(n, n, density=0.1, random_state=rng)
    return (R @ R.T + sp.identity(n) * 2.0).tocsr()


@pytest.mark.parametrize("ksp_type", ["cg", "gmres"])
def test_krylov_drivers_solve(ksp_type):
    A = _spd()
    b = np.random.default_rng(1).uniform(-1, 1, A.shape[0])
    ksp = pl.KSP()
    ksp.setOperators(pl.Mat(A))
    ksp.type = ksp_type
    ksp.rtol = 1e-10
    ksp.getPC().setType("jacobi")
    ksp.setConvergenceHistory()
    x = pl.Vec(np.zeros_like(b))
    ksp.solve(pl.Vec(b), x)
    assert np.linalg.norm(A @ x.array - b) < 1e-8 * np.linalg.norm(b)
    hist = ksp.getConvergenceHistory()
    assert len(hist) == ksp.getIterationNumber() + 1 and hist[-1] <= 1e-10 * hist[0]
    xs = spla.spsolve(A.tocsc(), b)
    assert np.allclose(x.array, xs, rtol=1e-6, atol=1e-9)


def test_options_database_prefixes_and_nesting():
    pl.Options.clear()
    pl.Options.insert_parameters({"ksp_type": "cg", "ssc_sub_0": {"pc_patch_save_operators": True, "sub_pc_type": "lu"}}, prefix="a_")
    assert pl.Options("a_").getString("ksp_type") == "cg"
    assert pl.Options("a_ssc_sub_0_").getBool("pc_patch_save_operators") is True
    assert pl.Options("a_ssc_sub_0_").getString("sub_pc_type") == "lu"
    assert not pl.Options("b_").hasName("ksp_type")
    assert pl.Options("a_").getInt("missing", 7) == 7
    pl.Options.clear()


def test_patchpc_finds_form_like_the_reference():
    from ssc_b200.patch import _form_and_bcs
    form, bcs = fem.poisson_problem(plex.IntervalMesh(4), 1, "scalar")
    A = form.assemble(bcs)
    # (1) python (matfree) operator: form and bcs from the matrix context
    pc = pl.PC()
    pc.setOperators(pl.Mat(context=fem.ImplicitMatrixContext(form, bcs), size=A.shape))
    assert _form_and_bcs(pc)[2] is form
    # (2) assembled operator: from the DM's application context
    pc = pl.PC()
    pc.setOperators(pl.Mat(A))
    pc.setDM(pl.DM(fem.SNESContext(form, bcs)))
    assert _form_and_bcs(pc)[2] is form
    # errors of ssc/patch.py:236-250
    pc = pl.PC()
    pc.setOperators(pl.Mat(context=None, size=A.shape))
    with pytest.raises(ValueError, match="No context found on matrix"):
        _form_and_bcs(pc)
    pc = pl.PC()
    pc.setOperators(pl.Mat(A))
    pc.setDM(pl.DM(None))
    with pytest.raises(ValueError, match="No context found on form"):
        _form_and_bcs(pc)
    ctx = fem.ImplicitMatrixContext(form, bcs, col_bcs=[])
    pc = pl.PC()
    pc.setOperators(pl.Mat(context=ctx, size=A.shape))
    with pytest.raises(NotImplementedError, match="Row and column bcs must match"):
        _form_and_bcs(pc)


def test_ssc_names_resolve():
    """-pc_python_type ssc.PatchPC / ssc.SSC and the relaxation classes (ssc/__init__.py:1-3)."""
    import ssc
    from ssc_b200.patch import PatchPC
    assert ssc.PatchPC is PatchPC and isinstance(ssc.PatchPC, type)
    for name in ("SSC", "PlaneSmoother", "OrderedVanka", "OrderedStar"):
        assert hasattr(ssc, name)
    ssc.PatchPC()          # instantiable with no arguments


def test_injection_is_the_left_inverse_of_the_embedding():
    """ssc/lo.py:146-159 interpolates the near-nullspace vectors of P to P1; for nodal spaces that is the injection of the
    vertex values: injection(E u) == u for every P1 function u, scalar / vector / mixed, simplices and tensor cells."""
    from ssc_b200 import fem, plex
    from ssc_b200.lo import embedding_matrix, injection_dofs
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    for mesh, p, bs in ((plex.TensorPlex((3, 2, 2)), 3, 1), (plex.TensorPlex((4, 3)), 2, 2), (plex.UnitSquareMesh(3, 3), 3, 1), (plex.BoxMesh(2, 2, 2), 2, 3)):
        Vk, V1 = fem.FunctionSpace(mesh, p, bs), fem.FunctionSpace(mesh, 1, bs)
        E = sp.kron(embedding_matrix(Vk, V1), sp.identity(bs), format="csr")
        u = rng.uniform(-1, 1, V1.dof_count)
        inj = injection_dofs([Vk], [V1])
        assert inj.shape == (V1.dof_count,)
        assert np.allclose((E @ u)[inj], u, atol=1e-14)
