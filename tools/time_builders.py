import sys, time, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((64, 64, 64)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
for dev in (True, False, True, False):
    pc = PC(0); setup_patch_pc(pc, disc, None); pc.setOption("ssc_device_patch_construction", dev)
    t = time.time(); pc.buildPatches(); dt = time.time() - t
    print("device" if dev else "host", "%.3f s" % dt, pc.size("lists_from_device"), pc.size("num_gtol"))
    pc.destroy()
