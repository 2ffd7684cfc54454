"""The multigrid coarse solver of the low-order PC (-lo_pc_type gamg; SURVEY.md 8(f) N1, ssc/lo.py:141-198 hands the P1 operator
to the solver the options name).  Checked against a sparse direct solve of the same operator."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import rhs_vector
from ssc_b200 import fem, plex, synth

pytestmark = pytest.mark.gpu


def _coarse_problem(c):
    mesh = plex.TensorPlex((c, c, c))
    V1 = fem.FunctionSpace(mesh, 1)
    d1 = fem.Discretisation([V1])
    r1, c1, v1 = synth.tensor_poisson_csr(V1, d1.ghost_bc_nodes)
    n = V1.node_count
    A1 = sp.csr_matrix((v1, c1, r1), shape=(n, n))
    keep = np.ones(n)
    keep[d1.global_bc_nodes] = 0.0
    R = sp.diags(keep).tocsr()                 # fine space = coarse space: the restriction is the identity off the boundary
    R.eliminate_zeros()
    return A1, R, np.asarray(d1.global_bc_nodes), keep


def _make(A1, R, bc, ksp, pc, rtol=1e-12):
    from ssc_b200.lo import LowOrderCore
    lo = LowOrderCore(0)
    lo.setOption("lo_ksp_type", ksp)
    lo.setOption("lo_pc_type", pc)
    lo.setOption("lo_ksp_rtol", rtol)
    lo.setRestriction(R)
    lo.setBcs(bc)
    lo.setOperator(A1)
    lo.setUp()
    return lo


@pytest.mark.parametrize("c", [6, 14, 28])           # 343 dofs: one dense level; 3375: two levels; 24389: three
def test_multigrid_cg_solves_the_coarse_problem(c):
    A1, R, bc, keep = _coarse_problem(c)
    x = rhs_vector(A1.shape[0])
    ref = keep * spla.spsolve(A1.tocsc(), keep * x)
    ref[bc] = x[bc]                                    # Dirichlet values pass through (ssc/lo.py:196-198)
    lo = _make(A1, R, bc, "cg", "gamg")
    y = np.zeros_like(x)
    lo.apply(x, y)
    assert np.linalg.norm(y - ref) <= 1e-9 * np.linalg.norm(ref)
    last, total, solves = lo.iterations()
    assert solves == 1 and last <= 30                  # h-independent: 12-13 iterations to 1e-10 from 16^3 to 96^3 cells (tools/micro/amg_host_check.cpp)
    jac = _make(A1, R, bc, "cg", "jacobi")
    yj = np.zeros_like(x)
    jac.apply(x, yj)
    assert np.linalg.norm(y - yj) <= 1e-8 * np.linalg.norm(ref)
    if c >= 14:
        assert last < jac.iterations()[0]


def test_one_cycle_is_a_symmetric_positive_operator():
    """-lo_ksp_type preonly -lo_pc_type gamg applies ONE V(1,1) cycle: same smoother before and after, R = P^T, so the
    coarse correction stays a fixed symmetric positive operator (what CG outside needs)."""
    A1, R, bc, keep = _coarse_problem(14)
    lo = _make(A1, R, bc, "preonly", "gamg")
    n = A1.shape[0]
    rng = np.random.default_rng(4)
    u, v = keep * rng.uniform(-1, 1, n), keep * rng.uniform(-1, 1, n)
    Mu, Mv = np.zeros(n), np.zeros(n)
    lo.apply(u, Mu)
    lo.apply(v, Mv)
    assert abs(v @ Mu - u @ Mv) <= 1e-12 * abs(v @ Mu)
    assert u @ Mu > 0 and v @ Mv > 0
    # and it is a good approximation of the inverse: the error of one cycle contracts
    e = keep * spla.spsolve(A1.tocsc(), u) - Mu
    assert np.sqrt(e @ (A1 @ e)) < 0.9 * np.sqrt((Mu + e) @ (A1 @ (Mu + e)))


def test_options_are_checked():
    from ssc_b200 import SSCError
    from ssc_b200.lo import LowOrderCore
    lo = LowOrderCore(0)
    with pytest.raises(SSCError):
        lo.setOption("lo_pc_type", "hypre")
    A1, R, bc, keep = _coarse_problem(4)
    lo.setOption("lo_ksp_type", "cg")
    lo.setOption("lo_pc_type", "lu")
    lo.setRestriction(R)
    lo.setBcs(bc)
    lo.setOperator(A1)
    with pytest.raises(SSCError):
        lo.setUp()


def test_exact_solve_options_past_the_dense_limit():
    """-lo_ksp_type preonly -lo_pc_type lu (the reference's test options, an exact coarse solve) on a coarse space too large for
    the dense factorisation: the solve is multigrid-CG to 1e-12 -- the same operator to rounding, no error and no silent loss of accuracy."""
    A1, R, bc, keep = _coarse_problem(28)             # 24 389 dofs > 4096
    x = rhs_vector(A1.shape[0])
    ref = keep * spla.spsolve(A1.tocsc(), keep * x)
    ref[bc] = x[bc]
    lo = _make(A1, R, bc, "preonly", "lu", rtol=1e-5)   # the loose tolerance must not leak into the exact solve
    y = np.zeros_like(x)
    lo.apply(x, y)
    assert np.linalg.norm(y - ref) <= 1e-10 * np.linalg.norm(ref)
    assert lo.iterations()[0] > 0


def test_block_size_option_on_a_vector_valued_operator():
    """-lo_amg_block_size 3 on a vector-valued coarse operator (dof = 3 * node + component; here the scalar operator Kronecker a 3 x 3
    SPD coupling): with node aggregates or with aggregates of the scalar graph the solve is that of a sparse direct solver."""
    from ssc_b200.lo import LowOrderCore
    A1, R1, bc1, keep1 = _coarse_problem(12)
    Cc = np.array([[2.0, 0.6, 0.0], [0.6, 2.0, 0.6], [0.0, 0.6, 2.0]])
    A = sp.kron(A1, sp.csr_matrix(Cc), format="csr")
    A.sort_indices()
    n = A.shape[0]
    R = sp.identity(n, format="csr")
    x = rhs_vector(n)
    ref = spla.spsolve(A.tocsc(), x)
    its = {}
    for bs in (3, 1):
        lo = LowOrderCore(0)
        for k, v in (("lo_ksp_type", "cg"), ("lo_pc_type", "gamg"), ("lo_ksp_rtol", 1e-12), ("lo_amg_block_size", bs)):
            lo.setOption(k, v)
        lo.setRestriction(R)
        lo.setBcs(np.zeros(0, dtype=np.int64))
        lo.setOperator(A)
        lo.setUp()
        y = np.zeros(n)
        lo.apply(x, y)
        assert np.linalg.norm(y - ref) <= 1e-9 * np.linalg.norm(ref)
        its[bs] = lo.iterations()[0]
    assert 0 < its[3] <= 40 and its[1] > 0


def test_near_nullspace_vectors():
    """SSCLowOrderSetNearNullSpace (ssc/lo.py:146-159): handing over the three componentwise constants explicitly builds the same
    hierarchy as -lo_amg_block_size 3 alone (same iteration count, same solution); a vector of the wrong length is refused."""
    from ssc_b200 import SSCError
    from ssc_b200.lo import LowOrderCore
    A1, R1, bc1, keep1 = _coarse_problem(10)
    Cc = np.array([[2.0, 0.6, 0.0], [0.6, 2.0, 0.6], [0.0, 0.6, 2.0]])
    A = sp.kron(A1, sp.csr_matrix(Cc), format="csr")
    A.sort_indices()
    n = A.shape[0]
    x = rhs_vector(n)
    ref = spla.spsolve(A.tocsc(), x)
    consts = [np.tile(np.eye(3)[c], n // 3) for c in range(3)]
    its, sols = [], []
    for vecs in (None, consts):
        lo = LowOrderCore(0)
        for k, v in (("lo_ksp_type", "cg"), ("lo_pc_type", "gamg"), ("lo_ksp_rtol", 1e-12), ("lo_amg_block_size", 3)):
            lo.setOption(k, v)
        if vecs is not None:
            lo.setNearNullSpace(vecs)
        lo.setRestriction(sp.identity(n, format="csr"))
        lo.setBcs(np.zeros(0, dtype=np.int64))
        lo.setOperator(A)
        lo.setUp()
        y = np.zeros(n)
        lo.apply(x, y)
        assert np.linalg.norm(y - ref) <= 1e-9 * np.linalg.norm(ref)
        its.append(lo.iterations()[0])
        sols.append(y)
    assert its[0] == its[1]
    lo = LowOrderCore(0)
    lo.setOption("lo_ksp_type", "cg")
    lo.setOption("lo_pc_type", "gamg")
    lo.setNearNullSpace([np.ones(n + 3)])
    lo.setRestriction(sp.identity(n, format="csr"))
    lo.setBcs(np.zeros(0, dtype=np.int64))
    lo.setOperator(A)
    with pytest.raises(SSCError):
        lo.setUp()
