"""A/B timing of the packed-symmetric apply variants on one box (64^3 Q3): python tools/gpu_sym_ab.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex, synth
from ssc_b200.patch import setup_patch_pc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
csr = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
rng = np.random.default_rng(0)
xh = rng.uniform(-1, 1, disc.total_dofs)


def run(name, opts):
    pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
    pc.setOperatorCSR(*csr)
    for k, v in opts.items():
        pc.setOption(k, v)
    pc.setUp()
    x = DeviceVec(pc, disc.total_dofs); y = DeviceVec(pc, disc.total_dofs); x.set(xh)
    for _ in range(3):
        pc.apply(x, y)
    best = 1e9
    for rep in range(3):
        e0, e1 = pc.eventCreate(), pc.eventCreate(); pc.eventRecord(e0)
        for _ in range(20):
            pc.apply(x, y)
        pc.eventRecord(e1); pc.synchronize()
        best = min(best, pc.eventElapsed(e0, e1) / 20)
    print("%-28s %.3f ms" % (name, best), flush=True)
    yh = y.get().copy()
    pc.destroy()
    return yh


sym = {"sub_pc_type": "cholesky", "ssc_symmetric_storage": True}
y0 = run("general", {})
for name, extra in (("packed one-copy", {"ssc_symmetric_two_chunk": False}),
                    ("packed two-chunk", {"ssc_symmetric_two_chunk": True}),
                    ("packed persistent", {"ssc_symmetric_persistent": True})):
    y = run(name, dict(sym, **extra))
    print("   rel diff vs general %.2e" % (np.abs(y - y0).max() / np.abs(y0).max()))
