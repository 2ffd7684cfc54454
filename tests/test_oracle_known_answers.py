"""Pins the CPU oracle against every known answer the reference's own tests hold for this path,
plus the size census and algebraic identities of the method."""
import numpy as np
import pytest

from conftest import rhs_vector
from helpers import OraclePythonPC, make_oracle, solve
from ssc_b200 import fem, plex
from ssc_b200 import petsc_lite as pl


def _problem(mesh, p, kind):
    form, bcs = fem.poisson_problem(mesh, p, kind)
    disc = fem.discretisation_from_form(form, bcs)
    return form, bcs, disc, form.assemble(bcs)


def test_p1_patch_pc_is_jacobi(mesh, problem_type):
    """tests/test_jacobi_equivalent.py:23-85: at degree 1 a vertex star owns exactly the vertex's dofs."""
    form, bcs, disc, A = _problem(mesh, 1, problem_type)
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    y = o.apply(x)
    expect = x / A.diagonal()
    expect[disc.global_bc_nodes] = x[disc.global_bc_nodes]          # BC pass-through, ssc/libssc.c:1408-1413
    assert np.allclose(y, expect, rtol=1e-14, atol=0)
    off, _ = o.patches()
    sizes = np.diff(off)
    bs_total = sum(W.bs for W in form.spaces)
    assert set(np.unique(sizes)) <= {0, bs_total}


def test_p1_residual_history_equals_jacobi(mesh, problem_type):
    """The reference's assertion itself (numpy.allclose of the two CG residual histories), with the oracle as the PC."""
    form, bcs, disc, A = _problem(mesh, 1, problem_type)
    b = np.ones(disc.total_dofs)
    b[disc.ghost_bc_nodes] = 0
    _, _, hj = solve(pl.Mat(A), b, "cg", lambda pc: pc.setType("jacobi"))
    opc = OraclePythonPC(disc, A)

    def use(pc):
        pc.setType("python")
        pc.setPythonContext(opc)
    _, _, hp = solve(pl.Mat(A), b, "cg", use)
    assert len(hj) == len(hp) and np.allclose(hj, hp)


@pytest.mark.parametrize("p,expected", [(1, 1), (2, 7), (3, 19), (4, 37), (5, 61), (6, 91)])
def test_triangle_patch_sizes(p, expected):
    """6-valent interior vertex of a P_p triangulation: 1 + 6(p-1) + 3(p-1)(p-2) dofs."""
    form, bcs, disc, _ = _problem(plex.RectangleMesh(6, 6, 1, 1), p, "scalar") if p <= 3 else (None, None, None, None)
    if disc is None:
        V = fem.FunctionSpace(plex.RectangleMesh(6, 6, 1, 1), p)
        disc = fem.Discretisation([V])
    o = make_oracle(disc)
    assert np.diff(o.patches()[0]).max() == expected == 1 + 6 * (p - 1) + 3 * (p - 1) * (p - 2)


@pytest.mark.parametrize("d,p", [(1, 3), (2, 2), (2, 3), (3, 2), (3, 3), (3, 4)])
def test_tensor_patch_size_census(d, p):
    """Q_p on a box: a vertex next to k boundary faces has (2p-1)^(d-k) (p-1)^k free dofs."""
    m = 4
    V = fem.FunctionSpace(plex.TensorPlex((m,) * d), p)
    o = make_oracle(fem.Discretisation([V]))
    sizes = np.diff(o.patches()[0])
    from math import comb
    want = {}
    for k in range(d + 1):
        cnt = comb(d, k) * 2 ** k * (m - 1) ** (d - k)
        n = (2 * p - 1) ** (d - k) * (p - 1) ** k
        want[n] = want.get(n, 0) + cnt
    got = dict(zip(*np.unique(sizes, return_counts=True)))
    assert {int(k): int(v) for k, v in got.items()} == want
    assert sizes.sum() == ((2 * p - 1) * (m - 1) + 2 * (p - 1)) ** d


def test_freudenthal_p2_vector_patch_is_45():
    V = fem.FunctionSpace(plex.BoxMesh(4, 4, 4), 2, 3)
    o = make_oracle(fem.Discretisation([V]))
    assert np.diff(o.patches()[0]).max() == 3 * (1 + 14)


def test_patch_matrix_is_submatrix_and_equals_elementwise_assembly():
    """Row A6 of SURVEY.md 8(a): for star patches the reference's cell-by-cell MatSetValues assembly
    (ssc/libssc.c:1150, negative indices dropped) equals A[gtol, gtol]."""
    mesh = plex.RectangleMesh(5, 4, 1, 1)
    V = fem.FunctionSpace(mesh, 3)
    A0, Ke = fem.assemble_poisson_simplex(V)
    disc = fem.Discretisation([V])
    A = fem.apply_bcs(A0, disc.ghost_bc_nodes)
    o = make_oracle(disc, A)
    off, gtol = o.patches()
    for p in range(0, o.npatch, 3):
        if off[p + 1] == off[p]:
            continue
        g = gtol[off[p]:off[p + 1]]
        B = o.patchMatrix(p)
        assert np.array_equal(B, A[g][:, g].toarray())
        assert np.allclose(o.patchMatrixFromElements(p, Ke), B, rtol=1e-12, atol=1e-13)
        assert np.all(np.linalg.eigvalsh((B + B.T) / 2) > 0)           # SPD (8c(iv))


@pytest.mark.parametrize("mesh,p,kind", [(plex.RectangleMesh(4, 3, 1, 1), 2, "mixed"), (plex.TensorPlex((3, 3, 2)), 3, "scalar"),
                                         (plex.TensorPlex((3, 4)), 2, "vector"), (plex.BoxMesh(2, 2, 2), 2, "tensor")],
                         ids=["tri-mixed", "hex-q3", "quad-vector", "tet-tensor"])
def test_elementwise_assembly_in_dof_table_order(mesh, p, kind):
    """Cell dofs of the element matrices run subspace -> node -> component (ssc/libssc.c:846-889); with that order the
    cell-by-cell assembly equals the submatrix of the assembled operator on every patch, also for mixed spaces."""
    form, bcs, disc, A = _problem(mesh, p, kind)
    Ke = fem.poisson_element_matrices(form.spaces)
    if Ke.ndim == 2:
        Ke = np.broadcast_to(Ke, (mesh.num_cells,) + Ke.shape).copy()
    o = make_oracle(disc, A)
    off, _ = o.patches()
    for i in range(o.npatch):
        if off[i + 1] > off[i]:
            assert np.allclose(o.patchMatrixFromElements(i, Ke), o.patchMatrix(i), rtol=1e-12, atol=1e-13)


def test_multiplicity_and_partition_of_unity():
    form, bcs, disc, A = _problem(plex.BoxMesh(3, 3, 3), 2, "vector")
    o = make_oracle(disc, A, partition_of_unity=1)
    m = o.multiplicityLocal(disc.total_dofs)
    off, gtol = o.patches()
    assert np.array_equal(m, np.bincount(gtol, minlength=disc.total_dofs))
    x = rhs_vector(disc.total_dofs)
    y1 = make_oracle(disc, A).apply(x)
    w = np.where(m > 0, 1.0 / np.maximum(m, 1), 0.0)                    # VecReciprocal leaves zeros (ssc/libssc.c:1282)
    assert np.allclose(o.apply(x), y1 * w, rtol=1e-15, atol=0)
    assert np.allclose((w * m)[m > 0], 1.0)


def test_symmetrise_sweep_doubles_the_additive_result():
    """ssc/libssc.c:1336-1343 on the additive code path."""
    form, bcs, disc, A = _problem(plex.RectangleMesh(5, 5, 1, 1), 2, "scalar")
    x = rhs_vector(disc.total_dofs)
    y1 = make_oracle(disc, A).apply(x)
    y2 = make_oracle(disc, A, symmetrise_sweep=1).apply(x)
    free = np.setdiff1d(np.arange(x.size), disc.global_bc_nodes)
    assert np.allclose(y2[free], 2 * y1[free], rtol=1e-14)
    assert np.array_equal(y2[disc.global_bc_nodes], x[disc.global_bc_nodes])


def test_threaded_oracle_matches_serial():
    form, bcs, disc, A = _problem(plex.TensorPlex((4, 4, 4)), 2, "scalar")
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    assert np.allclose(o.apply(x, nthreads=4), o.apply(x, nthreads=1), rtol=1e-13, atol=1e-15)


def test_user_patches_and_iteration_set():
    """PCPatchConstruct_User (ssc/libssc.c:1522-1547): a user patch equal to a vertex star reproduces the star patch;
    the iteration set may repeat patches (each visit adds again)."""
    mesh = plex.RectangleMesh(4, 4, 1, 1)
    form, bcs, disc, A = _problem(mesh, 2, "scalar")
    star = make_oracle(disc, A)
    vs, ve = mesh.depth_stratum(0)
    patches = [np.array(mesh.star(v)) for v in range(vs, ve)]
    from oracle.oracle import OraclePatchPC
    o = OraclePatchPC()
    o.setFromDiscretisation(disc)
    o.setUserPatches(patches, np.arange(len(patches)))
    o.setUpPatches()
    assert all(np.array_equal(a, b) for a, b in zip(o.patches(), star.patches()))
    o.setOperatorCSR(A)
    o.setUpOperators()
    x = rhs_vector(disc.total_dofs)
    assert np.allclose(o.apply(x), star.apply(x), rtol=1e-14)
    o2 = OraclePatchPC()
    o2.setFromDiscretisation(disc)
    o2.setUserPatches(patches, np.concatenate([np.arange(len(patches)), np.arange(len(patches))]))
    o2.setUpPatches()
    o2.setOperatorCSR(A)
    o2.setUpOperators()
    free = np.setdiff1d(np.arange(x.size), disc.global_bc_nodes)
    assert np.allclose(o2.apply(x)[free], 2 * star.apply(x)[free], rtol=1e-14)
