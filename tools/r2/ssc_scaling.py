"""ssc.SSC (patch + P1 coarse) at benchmark sizes: CG iteration counts flat in h, coarse-solve cost per apply.

    python tools/r2/ssc_scaling.py [cells ...]        (default 16 32 64; 3D Q3 Poisson, vertex-star patches)

Patch PC from the assembled CSR; coarse PC = SSCLowOrder* with -lo_ksp_type cg -lo_pc_type jacobi (device-resident PCG on
the Q1 operator), --gamg: the same CG preconditioned by the multigrid cycle, --one-cycle: -lo_ksp_type preonly -lo_pc_type gamg; the composite is formed on the device (ssc_b200/ssc_pc.py::SSC.applyDevice).  The outer CG runs on the
host with a scipy CSR matvec (the KSP is PETSc's job in the reference, ssc/ssc.py:38-49; tests/test_p_independence.py:44-60).
"""
import sys
import time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex, synth
from ssc_b200.lo import LowOrderCore
from ssc_b200.patch import setup_patch_pc


def run(c, p=3, rtol=1e-7):
    mesh = plex.TensorPlex((c, c, c))
    Vk, V1 = fem.FunctionSpace(mesh, p), fem.FunctionSpace(mesh, 1)
    dk, d1 = fem.Discretisation([Vk]), fem.Discretisation([V1])
    csr = synth.tensor_poisson_csr(Vk, dk.ghost_bc_nodes)
    A = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(Vk.node_count,) * 2)
    r1, c1, v1 = synth.tensor_poisson_csr(V1, d1.ghost_bc_nodes)
    A1 = sp.csr_matrix((v1, c1, r1), shape=(V1.node_count,) * 2)
    pc = PC(0); setup_patch_pc(pc, dk, None); pc._compute_op = None
    pc.setOperatorCSR(*csr)
    pc.setUp()
    lo = LowOrderCore(0)
    lo.setOption("lo_ksp_type", "preonly" if "--one-cycle" in sys.argv else "cg"); lo.setOption("lo_pc_type", "gamg" if ("--gamg" in sys.argv or "--one-cycle" in sys.argv) else "jacobi"); lo.setOption("lo_ksp_rtol", 1e-10)
    if "--no-persistent" in sys.argv:
        lo.setOption("lo_ksp_cg_persistent", "false")
    lo.setRestriction(synth.tensor_p1_restriction(Vk, V1, dk.ghost_bc_nodes, d1.ghost_bc_nodes, c, p))
    lo.setBcs(dk.global_bc_nodes)
    lo.setOperator(A1)
    t_lo = time.perf_counter(); lo.setUp(); t_lo = time.perf_counter() - t_lo
    n = Vk.node_count
    dx, dy, dw = (DeviceVec(pc, n) for _ in range(3))

    def M(r, coarse=True):
        dx.set(r)
        pc.apply(dx, dy)
        if coarse:
            lo.applyDevice(dx, dw)
            pc.synchronize()
            lo.axpyDevice(1.0, dw, dy)
            lo.synchronize()
        else:
            pc.synchronize()
        return dy.get()
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, n); b[dk.global_bc_nodes] = 0.0
    out = {"ssc": -1, "patch_only": -1}
    for coarse in (() if TIME_ONLY else (True, False)):
        x = np.zeros(n); r = b.copy(); z = M(r, coarse); pd = z.copy(); rz = r @ z; r0 = np.linalg.norm(r); its = 0
        while np.linalg.norm(r) > rtol * r0 and its < 500:
            q = A @ pd; alpha = rz / (pd @ q); x += alpha * pd; r -= alpha * q
            z = M(r, coarse); rz_new = r @ z; pd = z + (rz_new / rz) * pd; rz = rz_new; its += 1
        out["ssc" if coarse else "patch_only"] = its
    # timing of the applies (device-resident)
    e0, e1 = pc.eventCreate(), pc.eventCreate()
    dx.set(b)
    for _ in range(3):
        M(b)
    t = time.perf_counter()
    for _ in range(10):
        pc.apply(dx, dy); lo.applyDevice(dx, dw); pc.synchronize(); lo.axpyDevice(1.0, dw, dy); lo.synchronize()
    t_ssc = (time.perf_counter() - t) / 10
    t = time.perf_counter()
    for _ in range(10):
        pc.apply(dx, dy)
    pc.synchronize()
    t_patch = (time.perf_counter() - t) / 10
    last, total, solves = lo.iterations()
    print("Q%d %3d^3: dofs %9d coarse dofs %7d | CG its  ssc.SSC %3d  patch only %3d | apply ms  patch %.3f  patch+coarse %.3f | coarse CG its/solve %.1f | coarse set-up %.3f s"
          % (p, c, n, V1.node_count, out["ssc"], out["patch_only"], t_patch * 1e3, t_ssc * 1e3, total / max(solves, 1), t_lo), flush=True)
    pc.destroy()


if __name__ == "__main__":
    TIME_ONLY = "--time-only" in sys.argv          # skip the host-side outer CG (a scipy matvec per iteration: minutes at 64^3)
    for c in [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [16, 32, 64]:
        run(c)
