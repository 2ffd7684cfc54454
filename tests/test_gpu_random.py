"""Randomised parity: arbitrary (ragged, overlapping, partly empty) patch dof lists over a random SPD-ish operator,
handed over with SSCPatchSetPatches, against the oracle.  Covers sizes the mesh-based tests do not hit (n = 1 .. 260,
every remainder of the 8-wide unrolling, sub-warp groups, the n > 128 blocked factor kernel, the n > 256 apply path)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rhs_vector
from helpers import relerr
from oracle.oracle import OraclePatchPC

pytestmark = pytest.mark.gpu


def random_problem(seed, N=1500, npatch=60, nmax=260):
    rng = np.random.default_rng(seed)
    R = sp.random(N, N, density=8.0 / N, random_state=rng, format="csr")
    A = (R + R.T + sp.diags(np.full(N, 30.0) + rng.uniform(0, 5, N))).tocsr()      # diagonally dominant, symmetric
    A.sort_indices()
    sizes = rng.integers(0, nmax + 1, npatch)
    forced = np.minimum([0, 1, 2, 7, 8, 9, 129, 257], nmax)[:npatch]
    sizes[:forced.size] = forced
    lists = [np.sort(rng.choice(N, size=int(n), replace=False)) for n in sizes]
    off = np.concatenate(([0], np.cumsum([l.size for l in lists])))
    gtol = np.concatenate(lists) if off[-1] else np.zeros(0, dtype=np.int64)
    bc = np.sort(rng.choice(N, size=40, replace=False))
    return A, off, gtol, bc, N


@pytest.mark.parametrize("opts", [{}, {"ssc_apply_variant": 0}, {"ssc_deterministic": True},
                                  {"sub_pc_type": "cholesky"}, {"sub_pc_type": "cholesky", "ssc_symmetric_storage": True, "nmax": 200},
                                  {"pc_patch_save_operators": False}, {"pc_patch_partition_of_unity": True, "pc_patch_symmetrise_sweep": True}])
@pytest.mark.parametrize("seed", [1, 2])
def test_random_patch_lists(seed, opts):
    from ssc_b200 import PC
    opts = dict(opts)
    A, off, gtol, bc, N = random_problem(seed, nmax=opts.pop("nmax", 260))
    o = OraclePatchPC()
    o.setPatches(off, gtol, N)
    o.setOperatorCSR(A)
    o.setUpOperators()
    pc = PC(device=0)
    for k, v in opts.items():
        pc.setOption(k, v)
    pc.setPatches(off, gtol, N, global_bc_nodes=bc)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    x = rhs_vector(N, seed)
    ref = o.applyLocal(x)
    if opts.get("pc_patch_symmetrise_sweep"):
        ref = 2 * ref
    ref[bc] = x[bc]
    if opts.get("pc_patch_partition_of_unity"):
        m = np.bincount(gtol, minlength=N)
        ref = ref * np.where(m > 0, 1.0 / np.maximum(m, 1), 0.0)
    y = np.full(N, np.nan)
    pc.apply(x, y)
    assert relerr(y, ref) < 1e-12
    assert pc.size("npatch") == off.size - 1 and pc.size("sum_n") == gtol.size


def test_lifecycle_reset_and_reuse():
    """PCReset_PATCH (ssc/libssc.c:902-985): after reset the same handle takes a new problem."""
    from ssc_b200 import PC, SSCError
    pc = PC(device=0)
    for seed in (3, 4):
        A, off, gtol, bc, N = random_problem(seed, N=400, npatch=12, nmax=40)
        pc.setPatches(off, gtol, N, global_bc_nodes=bc)
        pc.setOperatorCSR(A.indptr, A.indices, A.data)
        pc.setUp()
        o = OraclePatchPC()
        o.setPatches(off, gtol, N)
        o.setOperatorCSR(A)
        o.setUpOperators()
        x = rhs_vector(N)
        ref = o.applyLocal(x)
        ref[bc] = x[bc]
        y = np.zeros(N)
        pc.apply(x, y)
        assert relerr(y, ref) < 1e-12
        pc.reset()
        with pytest.raises(SSCError):
            pc.apply(x, y)
    pc.destroy()
    pc.destroy()          # idempotent


def test_unsupported_combinations_are_errors():
    from ssc_b200 import PC, SSCError
    A, off, gtol, bc, N = random_problem(5, N=300, npatch=6, nmax=20)
    # ghosts named but no transport configured
    pc = PC(device=0)
    pc.setPatches(off, gtol, N)
    pc.setStarForest(N - 10, np.arange(N - 10), np.arange(N - 10), [1], [0, 0], [], [0, 10], np.arange(N - 10, N))
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert "SSCPatchCommInit" in e.value.message
    # multiplicative needs cells, deterministic is single-rank, operator must be set
    pc = PC(device=0)
    pc.setPatches(off, gtol, N)
    pc.setOption("pc_patch_multiplicative", True)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    with pytest.raises(SSCError):
        pc.setUp()
    pc = PC(device=0)
    pc.setPatches(off, gtol, N)
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert "SSCPatchSetOperatorCSR" in e.value.message
    # wrong vector length
    pc = PC(device=0)
    pc.setPatches(off, gtol, N)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    with pytest.raises(ValueError):
        pc.apply(np.zeros(N + 1), np.zeros(N + 1))
    # out-of-range dof in a patch list
    with pytest.raises(SSCError):
        PC(device=0).setPatches([0, 1], [N + 5], N)


def test_degenerate_inputs():
    """No patches at all / only empty patches: PCApply is zero on free dofs and the identity on Dirichlet dofs."""
    from ssc_b200 import PC
    import scipy.sparse as sp
    N = 10
    A = sp.identity(N, format="csr")
    x = rhs_vector(N)
    for off in ([0], [0, 0, 0]):
        pc = PC(device=0)
        pc.setPatches(off, [], N, global_bc_nodes=[2, 7])
        pc.setOperatorCSR(A.indptr, A.indices, A.data)
        pc.setUp()
        y = np.full(N, np.nan)
        pc.apply(x, y)
        ref = np.zeros(N)
        ref[[2, 7]] = x[[2, 7]]
        assert np.array_equal(y, ref)
        assert pc.size("sum_n") == 0 and pc.size("num_classes") == 0
    # a single 1x1 patch
    pc = PC(device=0)
    pc.setPatches([0, 1], [4], N)
    pc.setOperatorCSR((A * 4.0).indptr, A.indices, (A * 4.0).data)
    pc.setUp()
    y = np.zeros(N)
    pc.apply(x, y)
    assert y[4] == x[4] / 4.0 and np.count_nonzero(y) == 1


@pytest.mark.parametrize("opts", [{}, {"ssc_deterministic": True}, {"pc_patch_partition_of_unity": True}], ids=["default", "deterministic", "pou"])
@pytest.mark.parametrize("ctas", [1, 3])
def test_persistent_apply_kernel_many_rounds(ctas, opts):
    """The persistent small-patch apply kernel (n <= 96) prefetches the gathers of the next two patches while one streams.
    It is only chosen for classes with many patches per CTA, which no small test reaches: here it is forced
    (ssc_apply_stream_min_waves 0) onto one or three CTAs, so every group walks through many rounds, including
    partial last rounds, for every group width from a thread per patch (n = 1) to 96."""
    from ssc_b200 import PC
    rng = np.random.default_rng(99)
    N = 3000
    R = sp.random(N, N, density=6.0 / N, random_state=rng, format="csr")
    A = (R + R.T + sp.diags(np.full(N, 25.0) + rng.uniform(0, 5, N))).tocsr()
    A.sort_indices()
    lists = []
    for n, cnt in ((1, 300), (2, 150), (3, 37), (5, 37), (9, 61), (16, 37), (27, 53), (31, 37), (32, 37), (45, 41), (64, 19), (81, 11), (96, 9)):
        lists += [np.sort(rng.choice(N, size=n, replace=False)) for _ in range(cnt)]
    order = rng.permutation(len(lists))                   # sizes interleaved in patch order: the classes are gathered by the sort
    lists = [lists[i] for i in order]
    off = np.concatenate(([0], np.cumsum([l.size for l in lists])))
    gtol = np.concatenate(lists)
    bc = np.sort(rng.choice(N, size=30, replace=False))
    o = OraclePatchPC()
    o.setPatches(off, gtol, N)
    o.setOperatorCSR(A)
    o.setUpOperators()
    pc = PC(device=0)
    pc.setOption("ssc_apply_stream_min_waves", 0)
    pc.setOption("ssc_apply_stream_ctas", ctas)
    for k, v in opts.items():
        pc.setOption(k, v)
    pc.setPatches(off, gtol, N, global_bc_nodes=bc)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    for seed in (5, 6):
        x = rhs_vector(N, seed)
        ref = o.applyLocal(x)
        ref[bc] = x[bc]
        if opts.get("pc_patch_partition_of_unity"):
            m = np.bincount(gtol, minlength=N)
            ref = ref * np.where(m > 0, 1.0 / np.maximum(m, 1), 0.0)
        y = np.full(N, np.nan)
        pc.apply(x, y)
        assert relerr(y, ref) < 1e-12
