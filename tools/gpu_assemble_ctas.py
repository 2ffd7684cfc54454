"""Assemble kernel: blocks in flight per SM (L2 working set) at 64^3 Q3: python tools/gpu_assemble_ctas.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((64, 64, 64)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
Ke = fem.element_matrix_tensor(V)
for k in (16, 8, 6, 4, 3, 2, 16):
    pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
    pc.setElementMatrices(Ke); pc.setOption("ssc_assemble_ctas_per_sm", k)
    pc.setUp()
    print("ctas per SM %2d  assemble %.2f ms" % (k, pc.eventTime("PCPATCHComputeOp")), flush=True)
    pc.destroy()
