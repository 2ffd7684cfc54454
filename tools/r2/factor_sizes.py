"""Factor kernels across patch sizes (element-matrix route): variant 1 (register tiles), 2 (L2-resident DMMA)."""
import sys, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc
cases = [((64, 64, 64), 3), ((1200, 1200), 4), ((700, 700), 5), ((500, 500), 6), ((400, 400), 7), ((350, 350), 8)]
for shape, p in cases:
    mesh = plex.TensorPlex(shape); V = fem.FunctionSpace(mesh, p); disc = fem.Discretisation([V])
    Ke = fem.element_matrix_tensor(V)
    for sub in ("lu", "cholesky"):
        out = []
        for variant in (1, 2, 3):
            pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
            pc.setElementMatrices(Ke)
            pc.setOption("sub_pc_type", sub); pc.setOption("ssc_factor_variant", variant)
            pc.setUp()
            out.append("v%d %.1f ms" % (variant, pc.eventTime("PCPATCHSolve")))
            n = int(np.diff(pc.getPatches()[0]).max()); npatch = pc.size("npatch")
            pc.destroy()
        print("Q%d %s (%d patches, max n = %d) %-8s %s" % (p, "x".join(map(str, shape)), npatch, n, sub, "  ".join(out)), flush=True)
