import sys, os, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex, synth
from ssc_b200.patch import setup_patch_pc
N = 64
mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
csr = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
for rb in (8, 4, 2):
    pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
    pc.setOperatorCSR(*csr)
    pc.setOption("ssc_extract_row_buffers", rb)
    pc.setUp()
    print("row buffers", rb, "extract ms", pc.eventTime("PCPATCHComputeOp"), flush=True)
    pc.setUp()
    print("   second setUp: extract ms", pc.eventTime("PCPATCHComputeOp"), "wall lists/upload/kernels", pc.eventTime("SSCSetupLists"), pc.eventTime("SSCSetupUploadOperator"), pc.eventTime("SSCSetupKernels"), flush=True)
    pc.destroy()
