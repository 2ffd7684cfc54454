import sys, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex, synth, SSCError
from ssc_b200.patch import setup_patch_pc
for c in (32, 40):
    mesh = plex.TensorPlex((c, c, c)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
    csr = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
    for sub, var in (("lu", 1), ("lu", 0), ("cholesky", 1), ("lu", 1)):
        pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
        pc.setOperatorCSR(*csr)
        pc.setOption("sub_pc_type", sub); pc.setOption("ssc_factor_variant", var)
        try:
            pc.setUp(); print(c, sub, var, "ok")
        except SSCError as e:
            off, gtol = pc.getPatches()
            sizes = np.diff(off)
            bad = pc.getFailedPatches()
            print(c, sub, var, "FAILED", bad.size, "sizes:", dict(zip(*np.unique(sizes[bad], return_counts=True))), "first", bad[:5], "last", bad[-5:])
        pc.destroy()
