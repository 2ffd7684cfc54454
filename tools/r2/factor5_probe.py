"""Round-2 probe of the register-resident DMMA factor kernel (ssc_factor_variant 5): correctness, then timing."""
import sys, os
import numpy as np
import scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_gpu, make_oracle, relerr
from ssc_b200 import PC, fem, plex, SSCError
from ssc_b200.patch import setup_patch_pc
from oracle.oracle import OraclePatchPC

V5 = int(os.environ.get("FVAR", "5"))
ok = True
# (a) FE blocks, every size class that the kernel takes
for mesh, p in ((plex.UnitSquareMesh(6, 6), 4), (plex.UnitSquareMesh(6, 6), 5), (plex.UnitSquareMesh(6, 6), 6), (plex.UnitSquareMesh(6, 6), 7),
                (plex.TensorPlex((4, 4, 4)), 3), (plex.TensorPlex((5, 6)), 4)):
    form, bcs = fem.poisson_problem(mesh, p, "scalar")
    disc = fem.discretisation_from_form(form, bcs)
    A = form.assemble(bcs)
    o = make_oracle(disc, A)
    x = np.random.default_rng(20260101).uniform(-1, 1, disc.total_dofs)
    ref = o.apply(x)
    for sub in ("lu", "cholesky"):
        pc = make_gpu(form, bcs, ssc_factor_variant=V5, sub_pc_type=sub)
        pc.setUp()
        y = np.zeros_like(x); pc.apply(x, y)
        e = relerr(y, ref)
        ok &= e < 1e-12
        print("FE dim %d p=%d %-8s nmax=%3d relerr %.2e" % (mesh.dim, p, sub, np.diff(pc.getPatches()[0]).max(), e), flush=True)
# (b) blocks that need pivoting
for n in (33, 64, 65, 96, 97, 100, 128):
    rng = np.random.default_rng(n)
    N = 4 * n
    A = sp.random(N, N, density=min(1.0, 40.0 / N), random_state=rng, format="lil")
    A.setdiag(0.0)
    perm = rng.permutation(N)
    for i in range(N):
        A[i, perm[i]] = 4.0 + rng.uniform()
    A = A.tocsr(); A.sort_indices()
    lists = [np.sort(rng.choice(N, size=n, replace=False)) for _ in range(6)]
    lists = [np.unique(np.concatenate([l[: n // 2], perm[l[: n // 2]]]))[:n] for l in lists]
    lists = [l for l in lists if np.linalg.cond(A[l][:, l].toarray()) < 1e6]
    off = np.concatenate(([0], np.cumsum([l.size for l in lists]))); gtol = np.concatenate(lists)
    o = OraclePatchPC(); o.setPatches(off, gtol, N); o.setOperatorCSR(A); o.setUpOperators()
    pc = PC(device=0); pc.setOption("ssc_factor_variant", V5); pc.setPatches(off, gtol, N); pc.setOperatorCSR(A.indptr, A.indices, A.data); pc.setUp()
    x = np.random.default_rng(1).uniform(-1, 1, N); y = np.zeros(N); pc.apply(x, y)
    e = relerr(y, o.applyLocal(x))
    worst = max(np.abs(pc.getPatchMatrix(i, inverse=True) @ A[l][:, l].toarray() - np.eye(l.size)).max() for i, l in enumerate(lists))
    ok &= e < 1e-9 and worst < 1e-8
    print("pivoting n=%3d sizes %s relerr %.2e  max|SB-I| %.2e  cond max %.1e" % (n, sorted(set(l.size for l in lists)), e, worst, max(np.linalg.cond(A[l][:, l].toarray()) for l in lists)), flush=True)
# (c) singular blocks are reported, early and late failures (exactly singular in floating point: repeated integer rows)
for n, dup in ((40, (0, 1)), (100, (98, 99)), (128, (5, 77)), (128, None)):
    rng = np.random.default_rng(7)
    M = rng.integers(-3, 4, (n, n)).astype(float) + 8.0 * np.eye(n)
    if dup is None:
        M[:, 3] = 0.0
    else:
        M[dup[1]] = M[dup[0]]
    A = sp.block_diag([sp.csr_matrix(M), sp.identity(n)]).tocsr()
    for sub in ("lu", "cholesky"):
        pc = PC(device=0); pc.setOption("ssc_factor_variant", V5); pc.setOption("sub_pc_type", sub)
        pc.setPatches([0, n, 2 * n], np.arange(2 * n), 2 * n); pc.setOperatorCSR(A.indptr, A.indices, A.data)
        try:
            pc.setUp(); print("singular n=%d %s %s: NOT reported" % (n, dup, sub)); ok = ok and sub == "cholesky"
        except SSCError as e:
            print("singular n=%d %s %s: ierr %d failed_patches %d" % (n, dup, sub, e.ierr, pc.size("failed_patches")))
            ok &= e.ierr == 71 and pc.size("failed_patches") == 1
print("CORRECTNESS", "OK" if ok else "FAILED", flush=True)
# (d) timing, 64^3 Q3 from one shared element matrix
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
Ke = fem.element_matrix_tensor(V)
x = np.random.default_rng(0).uniform(-1, 1, disc.total_dofs)
y0 = None
for sub in ("lu", "cholesky"):
    for variant in (4 if sub == "lu" else 3, V5):
        pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
        pc.setElementMatrices(Ke)
        pc.setOption("sub_pc_type", sub); pc.setOption("ssc_factor_variant", variant)
        pc.setUp()
        y = np.zeros_like(x); pc.apply(x, y)
        if y0 is None:
            y0 = y
        print("%-8s variant %d  factor %.1f ms  assemble %.1f ms   rel diff %.2e" % (sub, variant, pc.eventTime("PCPATCHSolve"), pc.eventTime("PCPATCHComputeOp"),
                                                                              np.linalg.norm(y - y0) / np.linalg.norm(y0)), flush=True)
        pc.destroy()
