"""A/B of running the small size classes beside the big one (ssc_overlap_classes): python tools/gpu_overlap_ab.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex
from ssc_b200.patch import setup_patch_pc

def case(name, disc, Ke, flush):
    for overlap in (False, True, False, True):
        pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
        pc.setElementMatrices(Ke)
        pc.setOption("ssc_overlap_classes", overlap)
        pc.setUp()
        x = DeviceVec(pc, disc.total_dofs); y = DeviceVec(pc, disc.total_dofs); x.set(np.random.default_rng(0).uniform(-1, 1, disc.total_dofs))
        for _ in range(3): pc.apply(x, y)
        e0, e1 = pc.eventCreate(), pc.eventCreate()
        tot = 0.0
        for _ in range(30):
            if flush: pc.flushL2()
            pc.eventRecord(e0); pc.apply(x, y); pc.eventRecord(e1); pc.synchronize()
            tot += pc.eventElapsed(e0, e1)
        print("%-22s overlap=%-5s %.4f ms per apply" % (name, overlap, tot / 30), flush=True)
        pc.destroy()

mesh = plex.TensorPlex((64, 64, 64)); V = fem.FunctionSpace(mesh, 3)
case("Q3 64^3", fem.Discretisation([V]), fem.element_matrix_tensor(V), False)
mesh = plex.TensorPlex((24, 24, 24)); V = fem.FunctionSpace(mesh, 3)
case("Q3 24^3", fem.Discretisation([V]), fem.element_matrix_tensor(V), True)
mesh = plex.BoxMesh(32, 32, 32); V = fem.FunctionSpace(mesh, 2, 3)
case("elasticity tets 32^3", fem.Discretisation([V]), fem.elasticity_element_matrices(V), True)
mesh = plex.RectangleMesh(256, 256, 1, 1); V = fem.FunctionSpace(mesh, 4)
case("P4 triangles 256^2", fem.Discretisation([V]), fem.assemble_poisson_simplex(V)[1], True)
