"""Concurrent pinned H2D / D2H bandwidth per rank (torchrun), with and without NUMA binding: torchrun ... tools/mgpu_pcie.py [bind]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch.distributed as dist
rank, nranks, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
if bind:
    sys.path.insert(0, ".")
    import bench
    n = bench.bind_near_gpu(local)
dist.init_process_group("gloo", rank=rank, world_size=nranks)
from ssc_b200 import PC, DeviceVec, PinnedArray
pc = PC(device=local)
pc.setPatches([0, 1], [0], 1)
N = 7189057
h = PinnedArray(pc, N); d = DeviceVec(pc, N)
h.array[:] = 1.0
import ctypes as C
from ssc_b200 import _lib
lib = _lib.lib
from ssc_b200.pcpatch import MEM_DEVICE, MEM_HOST
hp = h.array.ctypes.data_as(C.c_void_p)
def copy(kind):
    if kind == 0:
        lib.SSCMemcpy(pc.handle, d.ptr, hp, C.c_int64(8 * N), MEM_DEVICE, MEM_HOST)
    else:
        lib.SSCMemcpy(pc.handle, hp, d.ptr, C.c_int64(8 * N), MEM_HOST, MEM_DEVICE)
for kind, name in ((0, "H2D"), (1, "D2H")):
    for _ in range(3):
        copy(kind)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        copy(kind)
    pc.synchronize()
    dt = (time.perf_counter() - t0) / 20
    out = [None] * nranks
    dist.all_gather_object(out, round(8 * N / dt / 1e9, 1))
    if rank == 0:
        print("bind=%s %s GB/s per rank (all ranks copying at once): %s  affinity %d cpus" % (bind, name, out, len(os.sched_getaffinity(0))), flush=True)
dist.barrier(); dist.destroy_process_group()
