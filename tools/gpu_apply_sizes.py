"""Apply kernel across patch sizes: achieved fraction of the measured HBM copy bandwidth (algorithmic bytes of SURVEY 8(d)).
python tools/gpu_apply_sizes.py"""
import sys, json, os
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex
from ssc_b200.patch import setup_patch_pc
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6537.6) if os.path.exists("MEASURED_PEAKS.json") else 6537.6

def case(name, mesh, p, opts=()):
    V = fem.FunctionSpace(mesh, p); disc = fem.Discretisation([V])
    pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
    pc.setElementMatrices(fem.element_matrix_tensor(V))
    for k, v in opts: pc.setOption(k, v)
    pc.setUp()
    x = DeviceVec(pc, disc.total_dofs); y = DeviceVec(pc, disc.total_dofs); x.set(np.random.default_rng(0).uniform(-1, 1, disc.total_dofs))
    for _ in range(3): pc.apply(x, y)
    e0, e1 = pc.eventCreate(), pc.eventCreate()
    flush = pc.size("block_bytes") < (1 << 30)
    tot = 0.0; reps = 20
    for _ in range(reps):
        if flush: pc.flushL2()
        pc.eventRecord(e0); pc.apply(x, y); pc.eventRecord(e1); pc.synchronize()
        tot += pc.eventElapsed(e0, e1)
    ms = tot / reps
    b = pc.size("bytes_apply")
    print("%-22s n_max %4d  dofs %9d  bytes %.2f GB  %.3f ms  %.2f Gdof/s  frac of measured HBM %.3f %s" % (name, int(np.diff(pc.getPatches()[0]).max()), disc.total_dofs, b / 1e9, ms, disc.total_dofs / ms / 1e6, b / (ms * 1e-3) / 1e9 / peak, dict(opts) or ""), flush=True)
    pc.destroy()

for opts in ((),):
    case("3D Q1 160^3", plex.TensorPlex((160, 160, 160)), 1, opts)
    case("2D Q2 1536^2", plex.TensorPlex((1536, 1536)), 2, opts)
    case("2D Q3 1024^2", plex.TensorPlex((1024, 1024)), 3, opts)
    case("3D Q2 112^3", plex.TensorPlex((112, 112, 112)), 2, opts)
