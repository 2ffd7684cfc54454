"""First set-up wall time at 64^3 Q3 and re-set-up time against the patch-builder thread count (0 = the library's split between builder and upload staging).
Run on a GPU box: python tools/r2/setup_sweep.py"""
import os, sys, time, types
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import bench
from ssc_b200 import PC
from ssc_b200.patch import setup_patch_pc

args = bench.parse()
V, disc, slab, info = bench.build_problem(args, 0, 1)
csr = bench.operator_csr(V, disc)
for tb in [0, 0, 0, 16, 12, 8, 0, 0]:      # 0: the library's own split (builder = cores - staging threads)
    pc = PC(device=0)
    if tb:
        pc.setOption("ssc_builder_threads", tb)
    setup_patch_pc(pc, disc, None)
    pc._compute_op = None
    pc.setOperatorCSR(*csr)
    t0 = time.time()
    pc.setUp(); pc.synchronize()
    t = time.time() - t0
    ev = {k: round(pc.eventTime(k)) for k in ("PCPATCHCreate", "SSCListsPack", "SSCListsUpload", "SSCSetupLists", "SSCSetupUploadOperator", "SSCSetupKernels")}
    print("builder %2d: setup %.3f s  %s" % (tb, t, ev), flush=True)
    pc.setOperatorValues(csr[2])
    t0 = time.time()
    pc.setUp(); pc.synchronize()
    print("            re-setup %.3f s  kernels %d ms (extract %.1f + factor %.1f)" % (time.time() - t0, pc.eventTime("SSCSetupKernels"), pc.eventTime("PCPATCHComputeOp"), pc.eventTime("PCPATCHSolve")), flush=True)
    pc.destroy() if hasattr(pc, "destroy") else None
    del pc
