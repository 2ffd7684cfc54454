"""Factor kernels across patch sizes (element-matrix route, tensor Q_p hexes): python tools/gpu_factor_sizes.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc

for N, p in ((96, 1), (80, 2), (64, 3), (24, 4)):
    mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, p); disc = fem.Discretisation([V])
    Ke = fem.element_matrix_tensor(V)
    for sub in ("lu", "cholesky"):
        out = []
        for variant in (1, 2, 3):
            pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
            pc.setElementMatrices(Ke)
            pc.setOption("sub_pc_type", sub); pc.setOption("ssc_factor_variant", variant)
            pc.setUp()
            out.append("v%d %.1f ms" % (variant, pc.eventTime("PCPATCHSolve")))
            n = int(np.diff(pc.getPatches()[0]).max())
            pc.destroy()
        print("Q%d %d^3 (max n = %d) %-8s %s" % (p, N, n, sub, "  ".join(out)), flush=True)
