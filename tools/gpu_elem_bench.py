"""Set-up of the 64^3 Q3 configuration from element matrices (shared and per cell) next to the CSR route."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex, synth
from ssc_b200.patch import setup_patch_pc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
Ke = fem.element_matrix_tensor(V)
x = np.random.default_rng(0).uniform(-1, 1, disc.total_dofs)
ys = []
VAR = [("csr", {}), ("shared Ke", {}), ("per-cell Ke", {}), ("shared Ke", {"ssc_builder_threads": 4}), ("shared Ke", {"ssc_builder_threads": 16}), ("shared Ke", {"ssc_builder_threads": 32})]
for name, opts in VAR:
    pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
    if name == "csr":
        pc.setOperatorCSR(*synth.tensor_poisson_csr(V, disc.ghost_bc_nodes))
    elif name == "shared Ke":
        pc.setElementMatrices(Ke)
    else:
        pc.setElementMatrices(np.broadcast_to(Ke, (mesh.num_cells,) + Ke.shape).copy())
    for k, v in opts.items():
        pc.setOption(k, v)
    t0 = time.perf_counter(); pc.setUp(); t1 = time.perf_counter()
    ev = {k: pc.eventTime(k) for k in ("PCPATCHComputeOp", "PCPATCHSolve", "SSCSetupLists", "PCPATCHCreate", "SSCListsSort", "SSCListsPack", "SSCListsUpload", "SSCSetupUploadOperator", "SSCSetupAllocBlocks")}
    y = np.zeros_like(x); pc.apply(x, y); ys.append(y)
    print("%-12s %s setUp wall %.3f s  %s" % (name, opts, t1 - t0, {k: round(v, 1) for k, v in ev.items()}), flush=True)
    pc.destroy()
for y in ys[1:]:
    print("rel diff vs csr route %.2e" % (np.linalg.norm(y - ys[0]) / np.linalg.norm(ys[0])))
