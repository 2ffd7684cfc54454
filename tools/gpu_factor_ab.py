"""A/B of the factor kernels on one box (64^3 Q3 by default): python tools/gpu_factor_ab.py [N]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
Ke = fem.element_matrix_tensor(V)
x = np.random.default_rng(0).uniform(-1, 1, disc.total_dofs)
y0 = None
for sub in ("lu", "cholesky"):
    for variant in (3, 4):
        pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
        pc.setElementMatrices(Ke)
        pc.setOption("sub_pc_type", sub); pc.setOption("ssc_factor_variant", variant // 10 if variant >= 10 else variant); pc.setOption("ssc_factor_remap", variant < 10)
        pc.setUp()
        y = np.zeros_like(x); pc.apply(x, y)
        if y0 is None:
            y0 = y
        print("%-8s variant %d  factor %.1f ms  assemble %.1f ms   rel diff %.2e" % (sub, variant, pc.eventTime("PCPATCHSolve"), pc.eventTime("PCPATCHComputeOp"),
                                                                              np.linalg.norm(y - y0) / np.linalg.norm(y0)), flush=True)
        pc.destroy()
