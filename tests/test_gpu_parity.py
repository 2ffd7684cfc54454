"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): PCApply within 1e-12 relative of the reference path in fp64, identical
Krylov iteration counts; patch dof lists bit-exact.
"""
import numpy as np
import pytest

from conftest import rhs_vector
from helpers import entry_err, make_gpu, make_oracle, relerr
from ssc_b200 import fem, plex

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _problem(mesh, p, kind):
    form, bcs = fem.poisson_problem(mesh, p, kind)
    disc = fem.discretisation_from_form(form, bcs)
    A = form.assemble(bcs)
    return form, bcs, disc, A


@pytest.mark.parametrize("p", [1, 2, 3])
def test_apply_matches_oracle(mesh, problem_type, p):
    if problem_type in ("tensor", "mixed") and p == 3 and mesh.dim == 3:
        pytest.skip("large on the CPU oracle side; covered by vector")
    form, bcs, disc, A = _problem(mesh, p, problem_type)
    o = make_oracle(disc, A)
    pc = make_gpu(form, bcs)
    pc.setUp()
    assert all(np.array_equal(a, b) for a, b in zip(o.patches(), pc.getPatches()))
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL


@pytest.mark.parametrize("shape,p", [((6, 5), 2), ((5, 4), 4), ((3, 4, 3), 2), ((4, 4, 4), 3), ((2, 2, 2), 4), ((2, 2, 2), 5)])
def test_tensor_apply_and_blocks(shape, p):
    form, bcs, disc, A = _problem(plex.TensorPlex(shape), p, "scalar")
    o = make_oracle(disc, A)
    pc = make_gpu(form, bcs, ssc_keep_patch_operators=True)
    pc.setUp()
    off, _ = pc.getPatches()
    sizes = np.diff(off)
    for n in np.unique(sizes[sizes > 0]):
        i = int(np.nonzero(sizes == n)[0][len(np.nonzero(sizes == n)[0]) // 2])
        B = pc.getPatchMatrix(i)
        assert np.array_equal(B, o.patchMatrix(i))                  # extraction copies values: bit-exact
        Binv = pc.getPatchMatrix(i, inverse=True)
        assert np.abs(Binv @ B - np.eye(n)).max() < 1e-10
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL
    # a second right-hand side through device-resident vectors
    from ssc_b200 import DeviceVec
    x2 = rhs_vector(disc.total_dofs, 20260102)
    dx, dy = DeviceVec(pc, x2.size), DeviceVec(pc, x2.size)
    dx.set(x2)
    pc.apply(dx, dy)
    pc.synchronize()
    assert relerr(dy.get(), o.apply(x2)) < TOL


@pytest.mark.parametrize("opts", [dict(partition_of_unity=1), dict(symmetrise_sweep=1), dict(construct_type=1),
                                  dict(dim=1), dict(codim=0), dict(partition_of_unity=1, symmetrise_sweep=1)])
def test_options_match_oracle(opts):
    form, bcs, disc, A = _problem(plex.RectangleMesh(6, 5, 1, 1), 3, "vector")
    o = make_oracle(disc, A, **opts)
    pc = make_gpu(form, bcs, **opts)
    pc.setUp()
    assert all(np.array_equal(a, b) for a, b in zip(o.patches(), pc.getPatches()))
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL


def test_cholesky_and_unsaved_operators():
    form, bcs, disc, A = _problem(plex.TensorPlex((4, 4, 4)), 2, "scalar")
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    ref = o.apply(x)
    for opts in (dict(sub_pc_type="cholesky"), dict(pc_patch_save_operators=False), dict(pc_patch_sub_mat_type="dense")):
        pc = make_gpu(form, bcs, **opts)
        pc.setUp()
        y = np.zeros_like(x)
        pc.apply(x, y)
        assert relerr(y, ref) < TOL, opts


def test_elasticity_tets():
    """3D vector P2 elasticity on tets (BASELINE.json config 4, reduced): non-diagonal blocks, n = 45."""
    mesh = plex.BoxMesh(4, 4, 4)
    V = fem.FunctionSpace(mesh, 2, 3)
    A, _ = fem.assemble_elasticity_simplex(V)
    bcs = [fem.DirichletBC(0, V.boundary_nodes())]
    form = fem.Form([V], lambda b: fem.apply_bcs(A, fem.discretisation_from_form(form, b).ghost_bc_nodes))
    disc = fem.discretisation_from_form(form, bcs)
    Abc = form.assemble(bcs)
    o = make_oracle(disc, Abc)
    pc = make_gpu(form, bcs)
    pc.setUp()
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert np.diff(pc.getPatches()[0]).max() == 45
    assert relerr(y, o.apply(x)) < TOL


def test_setup_is_repeatable_and_tracks_operator_values():
    """PatchPC.update re-runs PCSetUp (ssc/patch.py:269-270): new operator values must be picked up."""
    form, bcs, disc, A = _problem(plex.RectangleMesh(5, 5, 1, 1), 2, "scalar")
    pc = make_gpu(form, bcs)
    pc._compute_op = None            # feed CSR values by hand below
    x = rhs_vector(disc.total_dofs)
    for scale in (1.0, 3.0):
        As = fem.apply_bcs(A * scale, disc.ghost_bc_nodes)
        pc.setOperatorCSR(As.indptr, As.indices, As.data)
        pc.setUp()
        o = make_oracle(disc, As)
        y = np.zeros_like(x)
        pc.apply(x, y)
        assert relerr(y, o.apply(x)) < TOL


def test_errors():
    from ssc_b200 import SSCError
    form, bcs, disc, A = _problem(plex.IntervalMesh(4), 1, "scalar")
    pc = make_gpu(form, bcs)
    with pytest.raises(SSCError):
        pc.applyTranspose(None, None)                                  # ssc/libssc.c:1715
    with pytest.raises(SSCError):
        pc.setOption("sub_pc_type", "ilu")
    x = np.zeros(disc.total_dofs)
    with pytest.raises(SSCError):
        pc.apply(x, x.copy())                                          # before setUp
    pc2 = make_gpu(form, bcs)
    pc2.setOption("pc_patch_construction_dim", 0)
    with pytest.raises(SSCError):
        pc2.setOption("pc_patch_construction_codim", 0)                # ssc/libssc.c:1576-1578


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("shape,p,kind", [((4, 4, 4), 3, "scalar"), ((3, 3), 4, "vector"), ((2, 2, 2), 2, "mixed")])
def test_apply_kernel_variants(variant, shape, p, kind):
    """Both apply kernels (the simple loop, which is also the fallback for very large patches, and the pipelined default)
    give the same answer."""
    form, bcs, disc, A = _problem(plex.TensorPlex(shape), p, kind)
    o = make_oracle(disc, A)
    pc = make_gpu(form, bcs, ssc_apply_variant=variant, ssc_keep_patch_operators=True)
    pc.setUp()
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL
    i = int(np.argmax(np.diff(pc.getPatches()[0])))
    assert np.array_equal(pc.getPatchMatrix(i), o.patchMatrix(i))


@pytest.mark.parametrize("factor_variant", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("n", [5, 33, 64, 65, 97, 100, 128, 150, 200, 256, 257, 300, 343, 512, 520])
def test_general_blocks_need_pivoting(n, factor_variant):
    """Non-symmetric operator with zero diagonal entries: -sub_pc_type lu must pivot.  Patches are arbitrary dof sets
    handed over with SSCPatchSetPatches; the oracle solves with LAPACK-style partial-pivot LU."""
    import scipy.sparse as sp
    from ssc_b200 import PC
    from oracle.oracle import OraclePatchPC
    rng = np.random.default_rng(n)
    N = 4 * n
    A = sp.random(N, N, density=min(1.0, 40.0 / N), random_state=rng, format="lil")
    A.setdiag(0.0)
    perm = rng.permutation(N)
    for i in range(N):                      # a permutation matrix + noise: nonsingular, zero diagonal
        A[i, perm[i]] = 4.0 + rng.uniform()
    A = A.tocsr()
    A.sort_indices()
    npatch = 6
    lists = [np.sort(rng.choice(N, size=n, replace=False)) for _ in range(npatch)]
    # make every block nonsingular: include, for each row dof, the column where its big entry sits
    lists = [np.unique(np.concatenate([l[: n // 2], perm[l[: n // 2]]]))[:n] for l in lists]
    # kappa <= 3e4: two correct fp64 solvers (the oracle's LU, a stored inverse) agree to ~kappa * eps, so north_star's 1e-12 is a statement about
    # blocks in this range; worse-conditioned blocks are the subject of test_refined_sub_solve_on_ill_conditioned_blocks (against an exact-residual solve)
    lists = [l for l in lists if np.linalg.cond(A[l][:, l].toarray()) < 3e4]
    assert lists
    off = np.concatenate(([0], np.cumsum([l.size for l in lists])))
    gtol = np.concatenate(lists)
    o = OraclePatchPC()
    o.setPatches(off, gtol, N)
    o.setOperatorCSR(A)
    o.setUpOperators()
    pc = PC(device=0)
    pc.setOption("ssc_factor_variant", factor_variant)
    pc.setPatches(off, gtol, N)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    x = rhs_vector(N)
    y = np.zeros(N)
    pc.apply(x, y)
    ref = o.applyLocal(x)
    assert relerr(y, ref) < TOL                    # these blocks reach kappa = 2e4: measured 1e-13 with the stored inverse
    for i, l in enumerate(lists):
        B = A[l][:, l].toarray()
        assert np.abs(pc.getPatchMatrix(i, inverse=True) @ B - np.eye(l.size)).max() < 1e-10


@pytest.mark.parametrize("factor_variant", [0, 1, 2, 3])
def test_singular_block_is_reported(factor_variant):
    import scipy.sparse as sp
    from ssc_b200 import PC, SSCError
    A = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 4.0, 0.0], [0.0, 0.0, 1.0]]))
    pc = PC(device=0)
    pc.setOption("ssc_factor_variant", factor_variant)
    pc.setPatches([0, 2, 3], [0, 1, 2], 3)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert e.value.ierr == 71 and pc.size("failed_patches") == 1


@pytest.mark.parametrize("factor_variant", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("case", ["rank1_4", "nan_6", "zero_column_40", "zero_column_100", "zero_row_128", "zero_column_150", "zero_row_250", "zero_column_343", "zero_row_500"])
def test_block_that_fails_early_is_reported(case, factor_variant):
    """A block whose elimination breaks down at an EARLY step (rank 1, a NaN entry, a zero column): the kernels must leave
    through their failure path with the permutation only partly written (the unguarded qinv[perm[tid]]),
    report exactly that patch with the reference's error code, and keep the context usable."""
    import scipy.sparse as sp
    from ssc_b200 import PC, SSCError
    kind, n = case.rsplit("_", 1)
    n = int(n)
    rng = np.random.default_rng(n)
    M = rng.integers(-3, 4, (n, n)).astype(float) + 8.0 * np.eye(n)
    if kind == "rank1":
        M = np.outer(np.arange(1.0, n + 1), np.arange(2.0, n + 2))
    elif kind == "nan":
        M[0, 0] = np.nan
    elif kind == "zero_column":
        M[:, 1] = 0.0
    else:
        M[2, :] = 0.0
    good = 4.0 * np.eye(n) + np.diag(np.ones(n - 1), 1)
    A = sp.block_diag([sp.csr_matrix(good), sp.csr_matrix(np.nan_to_num(M, nan=1.0)), sp.csr_matrix(good)]).tocsr()
    A.sort_indices()
    if kind == "nan":
        A[n, n] = np.nan
    pc = PC(device=0)
    pc.setOption("ssc_factor_variant", factor_variant)
    pc.setPatches([0, n, 2 * n, 3 * n], np.arange(3 * n), 3 * n)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert e.value.ierr == 71 and pc.size("failed_patches") == 1
    assert list(pc.getFailedPatches()) == [1]
    # the device context survived: a good operator sets up and applies afterwards
    A2 = sp.block_diag([sp.csr_matrix(good)] * 3).tocsr()
    pc.setOperatorCSR(A2.indptr, A2.indices, A2.data)
    pc.setUp()
    x = rhs_vector(3 * n)
    y = np.zeros(3 * n)
    pc.apply(x, y)
    assert relerr(y[:n], np.linalg.solve(good, x[:n])) < 1e-13


@pytest.mark.parametrize("sub_pc", ["lu", "cholesky"])
@pytest.mark.parametrize("shape,p", [((4, 4, 4), 3), ((5, 6), 4), ((3, 3, 3), 2)])
def test_factor_variants_agree(shape, p, sub_pc):
    """The factor kernels (shared-memory Gauss-Jordan, register tiles, register-resident DMMA with look-ahead -- which picks
    its pivots with a 2^-13 threshold -- and the blocked DMMA kernel where it applies) produce the same inverses to rounding."""
    form, bcs, disc, A = _problem(plex.TensorPlex(shape), p, "scalar")
    inv = []
    for variant in (0, 1, 2, 3):
        pc = make_gpu(form, bcs, ssc_factor_variant=variant, sub_pc_type=sub_pc)
        pc.setUp()
        sizes = np.diff(pc.getPatches()[0])
        inv.append([pc.getPatchMatrix(int(i), inverse=True) for i in (np.argmax(sizes), np.argmin(np.where(sizes > 0, sizes, 1 << 30)), sizes.size // 2)])
    for other in inv[1:]:
        for X, Y in zip(inv[0], other):
            assert np.allclose(X, Y, rtol=1e-10, atol=1e-12 * np.abs(X).max())
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL


@pytest.mark.parametrize("pou", [0, 1])
@pytest.mark.parametrize("slabs", [-2, -5])
def test_pipelined_host_apply(slabs, pou):
    """Host-space applies overlap H2D(x) / patch kernels / D2H(y) slab by slab (forced on for this small case)."""
    from ssc_b200 import PinnedArray
    form, bcs, disc, A = _problem(plex.TensorPlex((5, 4, 4)), 3, "scalar")
    o = make_oracle(disc, A, partition_of_unity=pou)
    pc = make_gpu(form, bcs, ssc_host_pipeline=slabs, partition_of_unity=pou)
    pc.setUp()
    xh, yh = PinnedArray(pc, disc.total_dofs), PinnedArray(pc, disc.total_dofs)
    for seed in (20260101, 20260102, 20260103):
        xh.array[:] = rhs_vector(disc.total_dofs, seed)
        yh.array[:] = np.nan
        pc.apply(xh.array, yh.array)
        assert relerr(yh.array, o.apply(xh.array)) < TOL
    # pageable memory works too
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL


@pytest.mark.parametrize("sym", [0, 1])
@pytest.mark.parametrize("case", [("tensor3", 2), ("tri", 3), ("tet", 2)])
def test_multiplicative_coloured_sweep(case, sym):
    """N4: -pc_patch_multiplicative runs colour by colour.  The reference has no multiplicative code path to compare
    with (ssc/libssc.c:1568-1572), so the check is against the oracle's *sequential* multiplicative Schwarz visiting
    the patches in the colour-major order: equality shows the colours are conflict-free and the arithmetic is right."""
    kind, p = case
    mesh = {"tensor3": plex.TensorPlex((4, 4, 3)), "tri": plex.RectangleMesh(6, 5, 1, 1), "tet": plex.BoxMesh(3, 3, 3)}[kind]
    form, bcs, disc, A = _problem(mesh, p, "scalar")
    o = make_oracle(disc, A)
    pc = make_gpu(form, bcs, pc_patch_multiplicative=True, symmetrise_sweep=sym)
    pc.setUp()
    col = pc.getColours()
    ncol = pc.size("num_colours")
    assert ncol >= 2 and col.max() == ncol - 1
    # no two patches of a colour share a cell
    coff, cells = pc.getCells()
    for c in range(ncol):
        seen = np.concatenate([cells[coff[q]:coff[q + 1]] for q in np.nonzero(col == c)[0]])
        assert seen.size == np.unique(seen).size
    order = np.concatenate([np.nonzero(col == c)[0] for c in range(ncol)])
    if sym:
        order = np.concatenate([order, np.concatenate([np.nonzero(col == c)[0] for c in range(ncol - 1, -1, -1)])])
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.applyMultiplicative(x, order)) < TOL
    assert "Schwarz type: multiplicative" in pc.view(viewer=__import__("ssc_b200.petsc_lite", fromlist=["Viewer"]).Viewer())


def test_multiplicative_is_a_better_preconditioner():
    """GMRES with the multiplicative sweep needs fewer iterations than with the additive sum (sanity of the method)."""
    from helpers import solve
    from ssc_b200 import petsc_lite as pl
    form, bcs, disc, A = _problem(plex.RectangleMesh(8, 8, 1, 1), 3, "scalar")
    b = form.load_vector()
    b[disc.ghost_bc_nodes] = 0
    its = {}
    for mult in (False, True):
        pl.Options.clear()
        pl.Options.insert_parameters({"pc_type": "python", "pc_python_type": "ssc.PatchPC", "patch_pc_patch_multiplicative": mult})
        _, its[mult], _ = solve(pl.Mat(A), b, "gmres", lambda pc: pc.setFromOptions(), dm=pl.DM(fem.SNESContext(form, bcs)))
    pl.Options.clear()
    assert its[True] < its[False]


def test_deterministic_mode_is_bitwise_reproducible():
    """-ssc_deterministic: corrections go to per-patch slots and are summed per dof in a fixed order: no fp64 atomics,
    bit-identical results run to run, same values as the atomic path to rounding."""
    form, bcs, disc, A = _problem(plex.TensorPlex((5, 5, 4)), 3, "scalar")
    o = make_oracle(disc, A, symmetrise_sweep=1)
    pc = make_gpu(form, bcs, ssc_deterministic=True, symmetrise_sweep=1)
    pc.setUp()
    x = rhs_vector(disc.total_dofs)
    ys = []
    for _ in range(4):
        y = np.full_like(x, np.nan)
        pc.apply(x, y)
        ys.append(y)
    assert all(np.array_equal(ys[0], yy) for yy in ys[1:])
    assert relerr(ys[0], o.apply(x)) < TOL


@pytest.mark.parametrize("case", [("hex", (4, 4, 4), 3, "scalar"), ("tri", None, 4, "vector"), ("tet", None, 2, "scalar"), ("hex", (2, 2, 2), 1, "mixed")])
def test_packed_symmetric_storage(case):
    """-ssc_symmetric_storage with -sub_pc_type cholesky: half the block bytes, bulk-copy (TMA) apply kernel, same answer."""
    kind, shape, p, space = case
    mesh = plex.TensorPlex(shape) if kind == "hex" else (plex.RectangleMesh(5, 4, 1, 1) if kind == "tri" else plex.BoxMesh(3, 3, 3))
    form, bcs, disc, A = _problem(mesh, p, space)
    o = make_oracle(disc, A, partition_of_unity=1)
    pc = make_gpu(form, bcs, sub_pc_type="cholesky", ssc_symmetric_storage=True, partition_of_unity=1)
    pc.setUp()
    full = make_gpu(form, bcs)
    full.setUp()
    if pc.size("max_n") >= 100:                  # small blocks are dominated by the 128-byte alignment of each block
        assert pc.size("block_bytes") < 0.55 * full.size("block_bytes")
    for seed in (20260101, 20260102):
        x = rhs_vector(disc.total_dofs, seed)
        y = np.zeros_like(x)
        pc.apply(x, y)
        assert relerr(y, o.apply(x)) < TOL
    i = int(np.argmax(np.diff(pc.getPatches()[0])))
    Binv = pc.getPatchMatrix(i, inverse=True)
    assert np.array_equal(Binv, Binv.T)
    assert np.allclose(Binv, full.getPatchMatrix(i, inverse=True), rtol=1e-10, atol=1e-14)
    pc.setUp()                                   # set-up again: blocks are re-extracted, re-factored, re-packed
    y2 = np.zeros_like(x)
    pc.apply(x, y2)
    assert relerr(y2, o.apply(x)) < TOL


def test_packed_symmetric_storage_needs_cholesky():
    from ssc_b200 import SSCError
    form, bcs, disc, A = _problem(plex.IntervalMesh(6), 2, "scalar")
    pc = make_gpu(form, bcs, ssc_symmetric_storage=True)
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert e.value.ierr == 75


def test_packed_symmetric_many_patches():
    """Packed-symmetric storage on a mesh with thousands of patches of one size (many CTAs per SM, bulk copies in flight)."""
    from ssc_b200 import synth
    mesh = plex.TensorPlex((12, 12, 12))
    V = fem.FunctionSpace(mesh, 2)
    disc = fem.Discretisation([V])
    import scipy.sparse as sp
    r, c, v = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
    A = sp.csr_matrix((v, c, r), shape=(V.node_count,) * 2)
    o = make_oracle(disc, A)
    pc = make_gpu(disc, None, sub_pc_type="cholesky", ssc_symmetric_storage=True)
    pc._compute_op = None
    pc.setOperatorCSR(r, c, v)
    pc.setUp()
    for seed in (1, 2, 3):
        x = rhs_vector(disc.total_dofs, seed)
        y = np.zeros_like(x)
        pc.apply(x, y)
        assert relerr(y, o.apply(x)) < TOL


@pytest.mark.parametrize("case", [("hex", (17, 16, 16), 2, "scalar"), ("hex", (12, 12, 30), 1, "mixed"), ("tet", (16, 16, 16), 2, "vector"),
                                  ("quad", (70, 66), 3, "vector"), ("hexdim1", (12, 12, 12), 2, "scalar"), ("ghost", (16, 16, 20), 2, "scalar")])
def test_device_patch_construction_matches_host_and_oracle(case):
    """N2: star patches built by the CUDA kernel (more than 4096 patches) are bit-identical to the host builder's."""
    from ssc_b200 import PC
    from ssc_b200.patch import setup_patch_pc
    kind, shape, p, space = case
    opts = {}
    if kind == "tet":
        mesh = plex.BoxMesh(*shape)
    elif kind == "quad":
        mesh = plex.TensorPlex(shape)
    else:
        mesh = plex.TensorPlex(shape)
    if kind == "ghost":
        mesh.mark_ghost_vertices(2, 3, 15)
    if kind == "hexdim1":
        opts["pc_patch_construction_dim"] = 1
    spaces = {"scalar": [1], "vector": [mesh.dim], "mixed": [1, mesh.dim]}[space]
    Vs = [fem.FunctionSpace(mesh, p, bs) for bs in spaces]
    disc = fem.Discretisation(Vs, owned_nodes=[V.owned_node_count for V in Vs] if kind == "ghost" else None)
    out = {}
    for dev in (True, False):
        pc = PC(device=0)
        setup_patch_pc(pc, disc, None)
        pc.setOption("ssc_device_patch_construction", dev)
        for k, v in opts.items():
            pc.setOption(k, v)
        out[dev] = (pc.getPatches(), pc.getCells(), pc.size("lists_from_device"))
    assert out[True][2] == 1 and out[False][2] == 0          # the kernel really ran / the host builder really ran
    for a, b in zip(out[True][0] + out[True][1], out[False][0] + out[False][1]):
        assert np.array_equal(a, b)
    assert out[True][0][0][-1] > 0


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8])
def test_p_sweep_on_triangles(p):
    """BASELINE.json configs[1]: 2D Poisson P_p, p = 1..8, vertex-star patches (interior n = 1, 7, 19, 37, 61, 91, 127, 169:
    every factor kernel is hit -- register tiles up to 128, the blocked DMMA kernel above)."""
    mesh = plex.UnitSquareMesh(6, 6)
    form, bcs, disc, A = _problem(mesh, p, "scalar")
    o = make_oracle(disc, A)
    pc = make_gpu(form, bcs)
    pc.setUp()
    off, gtol = pc.getPatches()
    assert all(np.array_equal(a, b) for a, b in zip(o.patches(), (off, gtol)))
    assert np.diff(off).max() == 1 + 6 * (p - 1) + 3 * (p - 1) * (p - 2)
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    # north_star's tolerance for every degree (round 2 measured 2.7e-14 at p = 7 and 4.5e-14 at p = 8: the equispaced
    # Lagrange bases reach kappa = 4.5e3 there); relerr is the worse of the 2-norm and the max-norm error, and every
    # entry is also held to 1e-12 of the magnitude its own patches work at
    ref = o.apply(x)
    assert relerr(y, ref) < 1e-12
    assert entry_err(y, ref, o.applyScale(x)) < 1e-12


def test_released_operator_cannot_be_set_up_again():
    """The operator arrays are borrowed; after releaseOperator a further setUp must fail cleanly (wrong state), and
    work again once an operator is set."""
    from ssc_b200 import SSCError
    form, bcs, disc, A = _problem(plex.TensorPlex((3, 3, 3)), 2, "scalar")
    pc = make_gpu(form, bcs)
    pc._compute_op = None
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    pc.P = None
    pc.releaseOperator()
    y2 = np.zeros_like(x)
    pc.apply(x, y2)                           # the factors stay resident
    assert relerr(y2, y) < 1e-14               # (not bit-equal: the order of the fp64 atomics varies run to run)
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert e.value.ierr == 73
    pc.setOperatorCSR(A.indptr, A.indices, 2.0 * A.data)
    pc.setUp()
    pc.apply(x, y2)
    free = np.setdiff1d(np.arange(x.size), disc.global_bc_nodes)
    assert relerr(2.0 * y2[free], y[free]) < 1e-13


@pytest.mark.parametrize("row_buffers", [-1, 2, 0])
@pytest.mark.parametrize("mode", ["sorted", "unsorted", "duplicates"])
def test_extract_from_unsorted_and_duplicated_csr(mode, row_buffers):
    """The extract kernel stores a hit when every CSR row has strictly increasing columns (checked on the device when the
    operator is set) and accumulates otherwise: rows with shuffled columns, and rows whose entries are split into several
    duplicates (what MatSetValues(ADD_VALUES) leaves before assembly compresses them), give the same blocks.  Row buffers in
    shared memory (8 warps, fewer, or none: the path of very large patches) do not change the result."""
    import scipy.sparse as sp
    form, bcs, disc, A = _problem(plex.TensorPlex((3, 3, 3)), 3, "scalar")
    o = make_oracle(disc, A)
    indptr, indices, data = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.copy()
    rng = np.random.default_rng(5)
    if mode == "unsorted":
        for r in range(A.shape[0]):
            perm = rng.permutation(indptr[r + 1] - indptr[r]) + indptr[r]
            indices[indptr[r]:indptr[r + 1]] = indices[perm]
            data[indptr[r]:indptr[r + 1]] = data[perm]
    elif mode == "duplicates":
        # every entry split into two halves that are powers-of-two fractions (exactly summable), placed apart
        rows = np.repeat(np.arange(A.shape[0]), np.diff(indptr))
        rows2, cols2, vals2 = np.concatenate([rows, rows]), np.concatenate([indices, indices]), np.concatenate([0.25 * data, 0.75 * data])
        order = np.lexsort((rng.permutation(rows2.size), rows2))
        rows2, cols2, vals2 = rows2[order], cols2[order], vals2[order]
        indptr = np.concatenate(([0], np.cumsum(np.bincount(rows2, minlength=A.shape[0])))).astype(np.int64)
        indices, data = cols2.astype(np.int32), vals2
    pc = make_gpu(form, bcs, ssc_keep_patch_operators=True, ssc_extract_row_buffers=row_buffers)
    pc._compute_op = None
    pc.setOperatorCSR(indptr, indices, data)
    pc.setUp()
    sizes = np.diff(pc.getPatches()[0])
    for i in (int(np.argmax(sizes)), int(sizes.size // 2), 0):
        if sizes[i]:
            assert np.array_equal(pc.getPatchMatrix(i), o.patchMatrix(i))
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < TOL


def _solve_extended(B, b):
    """B^-1 b far beyond double precision: LU solve in float64, then refinement steps whose residual is accumulated EXACTLY
    (every product of two doubles and every sum is exact in rational arithmetic over Python ints scaled by a power of two);
    each step gains a factor 1/(kappa * eps), so five steps are exact to long double for kappa <= 1e10."""
    import scipy.linalg as sl
    from fractions import Fraction
    lu = sl.lu_factor(B)
    n = b.size
    Bf = [[Fraction(float(v)) for v in row] for row in B]
    bf = [Fraction(float(v)) for v in b]
    xf = [Fraction(float(v)) for v in sl.lu_solve(lu, b)]
    for _ in range(5):
        r = np.array([float(bf[i] - sum(Bf[i][j] * xf[j] for j in range(n))) for i in range(n)])
        d = sl.lu_solve(lu, r)
        xf = [xf[i] + Fraction(float(d[i])) for i in range(n)]
    return np.array([np.longdouble(float(v)) + np.longdouble(float(v - Fraction(float(v)))) for v in xf], dtype=np.longdouble)


@pytest.mark.parametrize("n", [24, 77, 125, 150])
def test_refined_sub_solve_on_ill_conditioned_blocks(n):
    """-ssc_sub_solve refine: stored-inverse product + one refinement step with a double-double residual.  On SPD blocks
    with kappa = 1e8 the plain stored inverse (and any LU solve in fp64, the oracle's included) is off by ~kappa * eps,
    the refined solve is accurate to a few eps of the extended-precision solution; -ssc_sub_solve auto picks the
    classes by their condition estimate."""
    import scipy.sparse as sp
    from ssc_b200 import PC
    rng = np.random.default_rng(n)
    blocks = []
    for k in range(2):
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        Bk = (Q * np.logspace(0, 8, n)) @ Q.T
        blocks.append(0.5 * (Bk + Bk.T))
    A = sp.block_diag([sp.csr_matrix(Bk) for Bk in blocks] + [sp.identity(5)]).tocsr()
    A.sort_indices()
    N = A.shape[0]
    off = np.concatenate((np.arange(3) * n, [2 * n + 5]))
    x = rhs_vector(N)
    truth = np.concatenate([_solve_extended(Bk, x[k * n:(k + 1) * n]) for k, Bk in enumerate(blocks)] + [x[2 * n:].astype(np.longdouble)])
    err = {}
    for mode in ("inverse", "refine", "auto", "refine2"):
        pc = PC(device=0)
        pc.setOption("ssc_sub_solve", mode.rstrip("2"))
        if mode == "refine2":
            pc.setOption("ssc_refine_steps", 2)
        pc.setPatches(off, np.arange(N), N)
        pc.setOperatorCSR(A.indptr, A.indices, A.data)
        pc.setUp()
        y = np.zeros(N)
        pc.apply(x, y)
        err[mode] = float(np.abs(y - truth).max() / np.abs(truth).max())
        if mode == "auto":
            assert pc.size("refined_classes") == 1 and pc.eventTime("SSCKappaMax") > 1e7      # the n-class, not the identity block
        if mode == "refine":
            assert pc.size("refined_classes") == 2
            assert pc.size("bytes_apply") == 8 * 3 * (2 * n * n + 25) + 20 * N + 16 * N
    # a step contracts the error by ||I - S B|| (about 1e-3 here: the left residual of a Gauss-Jordan inverse at kappa = 1e8)
    assert err["refine"] < 1e-12 and err["auto"] == err["refine"] and err["refine2"] < 1e-15, err
    assert err["inverse"] > 50 * err["refine"], err


@pytest.mark.parametrize("p", [3, 8])
def test_refined_sub_solve_on_the_p_sweep(p):
    """The refined solve on BASELINE configs[1] meshes: agrees with the oracle at north_star's tolerance like the stored
    inverse does, and is at least as close to the extended-precision solution."""
    mesh = plex.UnitSquareMesh(4, 4)
    form, bcs, disc, A = _problem(mesh, p, "scalar")
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    ref = o.apply(x)
    off, gtol = o.patches()
    truth = np.zeros(disc.total_dofs, dtype=np.longdouble)
    for i in range(off.size - 1):
        g = gtol[off[i]:off[i + 1]]
        if g.size:
            truth[g] += _solve_extended(A[g][:, g].toarray(), x[g])
    truth[disc.global_bc_nodes] = x[disc.global_bc_nodes]
    errs = {}
    for mode in ("inverse", "refine"):
        pc = make_gpu(form, bcs, ssc_sub_solve=mode)
        pc.setUp()
        y = np.zeros_like(x)
        pc.apply(x, y)
        assert relerr(y, ref) < TOL
        assert entry_err(y, ref, o.applyScale(x)) < TOL
        errs[mode] = float(np.abs(y - truth).max() / np.abs(truth).max())
    assert errs["refine"] < 1e-15 and errs["refine"] <= errs["inverse"] * 1.01, errs


def test_pageable_host_vectors_are_registered_once_and_give_the_same_result():
    """Host-space applies on plain numpy (pageable) vectors, the kind a PETSc Vec hands over: the library page-locks them
    on first use (cached by address), results equal those from pinned vectors; -ssc_register_host_vectors false leaves
    them alone; after reset the vectors are usable as ordinary memory again."""
    from ssc_b200 import PinnedArray
    form, bcs, disc, A = _problem(plex.TensorPlex((20, 16, 16)), 3, "scalar")
    assert disc.total_dofs * 8 > (1 << 20)
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    ref = o.apply(x)
    for reg in (True, False):
        pc = make_gpu(form, bcs, ssc_host_pipeline=-4, ssc_register_host_vectors=reg)
        pc.setUp()
        xs = [x.copy(), x.copy()]
        ys = [np.empty_like(x), np.empty_like(x)]
        for rep in range(3):
            for xv, yv in zip(xs, ys):
                yv[:] = np.nan
                pc.apply(xv, yv)
                assert relerr(yv, ref) < TOL
        assert pc.size("host_vectors_registered") == (4 if reg else 0)
        xh, yh = PinnedArray(pc, x.size), PinnedArray(pc, x.size)
        xh.array[:] = x
        pc.apply(xh.array, yh.array)
        assert relerr(yh.array, ys[0]) < 1e-14
        pc.reset()
        xs[0][:] = 0.0                           # plain memory again
        del pc


def test_unsaved_operators_report_a_singular_block_and_option_state_is_consistent():
    """(i) with -pc_patch_save_operators false a block that cannot be factorised inside PCApply used to add
    garbage silently; now the next synchronising call raises the reference's error code.  (ii) the storage mode cannot be
    switched after the first SetUp.  (iii) multiplicative sweeps reject the combinations they do not implement at SetUp."""
    import scipy.sparse as sp
    from ssc_b200 import PC, SSCError
    A = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 4.0, 0.0], [0.0, 0.0, 1.0]]))
    pc = PC(device=0)
    pc.setOption("pc_patch_save_operators", False)
    pc.setPatches([0, 2, 3], [0, 1, 2], 3)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()                                       # nothing is factorised yet
    x, y = np.ones(3), np.zeros(3)
    with pytest.raises(SSCError) as e:
        pc.apply(x, y)
    assert e.value.ierr == 71
    with pytest.raises(SSCError) as e:
        pc.setOption("pc_patch_save_operators", True)
    assert e.value.ierr == 73
    form, bcs, disc, A = _problem(plex.TensorPlex((3, 3, 3)), 2, "scalar")
    for extra in (dict(ssc_deterministic=True), dict(pc_patch_save_operators=False)):
        pc = make_gpu(form, bcs, pc_patch_multiplicative=True, **extra)
        with pytest.raises(SSCError) as e:
            pc.setUp()
        assert e.value.ierr == 56


def test_new_values_on_the_same_pattern():
    """SSCPatchSetOperatorValues: the update path of a nonlinear solve (new values, same nonzero pattern) re-uploads only the
    values; the result is that of a full setOperatorCSR with the new values."""
    from ssc_b200 import SSCError
    form, bcs, disc, A = _problem(plex.TensorPlex((4, 3, 3)), 3, "scalar")
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    ref = o.apply(x)
    free = np.setdiff1d(np.arange(x.size), disc.global_bc_nodes)
    pc = make_gpu(form, bcs)
    pc._compute_op = None
    with pytest.raises((SSCError, ValueError)):
        pc.setOperatorValues(A.data)                  # no pattern yet
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, ref) < TOL
    for scale in (4.0, 0.5):
        pc.setOperatorValues(scale * A.data)
        pc.setUp()
        pc.apply(x, y)
        assert relerr(scale * y[free], ref[free]) < TOL


def test_very_large_patch_takes_the_unlimited_paths():
    """nothing on the path may cap the patch size below what the reference's sub-KSP takes.  One patch of
    1500 dofs (a whole small operator, the shape of a coarse space handed to the patch engine): the extract kernel's shared
    memory is no longer tied to n, the factorisation falls back to the in-place global-memory Gauss-Jordan once the blocked
    kernel's panels stop fitting shared memory, and the apply kernel walks several rows per thread."""
    import scipy.sparse as sp
    from ssc_b200 import PC
    from oracle.oracle import OraclePatchPC
    n = 1500
    rng = np.random.default_rng(3)
    M = sp.random(n, n, density=0.01, random_state=rng, format="csr")
    A = (M + M.T + sp.diags(np.full(n, 12.0))).tocsr()
    A.sort_indices()
    B = sp.block_diag([A, sp.identity(7)]).tocsr()
    off, gtol, N = np.array([0, n, n + 7]), np.arange(n + 7), n + 7
    o = OraclePatchPC()
    o.setPatches(off, gtol, N)
    o.setOperatorCSR(B)
    o.setUpOperators()
    x = rhs_vector(N)
    ref = o.applyLocal(x)
    for sub in ("lu", "cholesky"):
        pc = PC(device=0)
        pc.setOption("sub_pc_type", sub)
        pc.setPatches(off, gtol, N)
        pc.setOperatorCSR(B.indptr, B.indices, B.data)
        pc.setUp()
        y = np.zeros(N)
        pc.apply(x, y)
        assert relerr(y, ref) < TOL


@pytest.mark.parametrize("order", ["ascending", "scrambled"])
def test_setup_overlapped_with_a_large_upload(order):
    """An operator whose values exceed the staging threshold (64 MB) travels through the pinned ring with milestones, and
    extract + factor of a patch start once the rows it reads have arrived (SSCPatchSetUp).  Patches in row order take
    their milestones one by one; scrambled patches of a large class are held back by a running maximum, and a small
    class waits for its last patch.  Either way the result is the oracle's, also after new values on the same pattern."""
    import scipy.sparse as sp
    from ssc_b200 import PC
    from oracle.oracle import OraclePatchPC
    N, band = 200_000, 22
    rng = np.random.default_rng(11)
    diags = [rng.uniform(-1.0, 1.0, N - abs(k)) for k in range(-band, band + 1)]
    diags[band] = np.full(N, 3.0 * band)
    A = sp.diags(diags, list(range(-band, band + 1)), format="csr")
    A.sort_indices()
    assert A.nnz * 8 >= 64 << 20
    starts_big = rng.integers(0, N - 64, 6000)                   # a class of 6000 patches with 24 dofs: per-milestone ranges
    starts_small = rng.integers(0, N - 64, 900)                  # a class of 900 patches with 40 dofs: one range
    if order == "ascending":
        starts_big.sort()
        starts_small.sort()
    lists = [s + np.sort(rng.choice(60, 24, replace=False)) for s in starts_big] + [s + np.sort(rng.choice(60, 40, replace=False)) for s in starts_small]
    off = np.concatenate([[0], np.cumsum([len(l) for l in lists])])
    gtol = np.concatenate(lists)
    o = OraclePatchPC()
    o.setPatches(off, gtol, N)
    o.setOperatorCSR(A)
    o.setUpOperators()
    x = rhs_vector(N)
    ref = o.applyLocal(x)
    pc = PC(device=0)
    pc.setPatches(off, gtol, N)
    pc.setOperatorCSR(A.indptr, A.indices, A.data)
    pc.setUp()
    y = np.zeros(N)
    pc.apply(x, y)
    assert relerr(y, ref) < TOL
    from ssc_b200._lib import lib
    assert lib.SSCPatchReleaseOperator(pc.handle) == 0        # the pattern's host arrays are handed back: a values-only set-up must not read them
    pc.setOperatorValues(2.0 * A.data)
    pc.setUp()
    pc.apply(x, y)
    assert relerr(2.0 * y, ref) < TOL
