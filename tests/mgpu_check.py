"""Multi-GPU parity check, one process per GPU:  torchrun --nproc-per-node N tests/mgpu_check.py [cells] [degree]

Every rank builds its z-slab, sets the star forest and the library's NCCL communicator, applies the GPU PC, and rank 0
compares the gathered result with the single-rank CPU oracle on the global mesh (tolerance 1e-12 relative)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    import scipy.sparse as sp
    from ssc_b200 import PC, partition, synth, fem, plex
    from ssc_b200.patch import setup_patch_pc
    from ssc_b200.pcpatch import get_unique_id
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pou = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    p2p = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    slabs = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # ssc_host_pipeline: negative forces the overlapped host-space apply on a small case
    rank, nranks, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo", rank=rank, world_size=nranks)

    def allgather(obj):
        out = [None] * nranks
        dist.all_gather_object(out, obj)
        return out
    axis = int(sys.argv[6]) if len(sys.argv) > 6 else 0       # slabs cut along x (default) or z
    cells, lengths = [c, c, c], [1, 1, 1]
    cells[axis], lengths[axis] = c * nranks, nranks
    slab = partition.Slab(tuple(cells), p, rank, nranks, axis=axis)
    pc = PC(device=local)
    setup_patch_pc(pc, slab.discretisation, None)
    pc._compute_op = None
    pc.setOperatorCSR(*synth.tensor_poisson_csr(slab.space, slab.discretisation.ghost_bc_nodes))
    if pou:
        pc.setOption("pc_patch_partition_of_unity", True)
    if slabs:
        pc.setOption("ssc_host_pipeline", slabs)
    sf = partition.star_forest(slab.global_ids, slab.nowned, rank, nranks, allgather)
    pc.setStarForest(slab.nowned, *sf)
    if p2p:
        partition.connect_peers(pc, sf, rank, allgather)
    else:
        uid = [get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pc.commInit(nranks, rank, uid[0])
    pc.setUp()
    xlat = np.random.default_rng(11).uniform(-1, 1, int(np.prod(slab.global_shape)))
    x = np.ascontiguousarray(xlat[slab.global_ids[:slab.nowned]])
    y = np.zeros_like(x)
    for _ in range(3):
        y[:] = np.nan
        pc.apply(x, y)
    pc.synchronize()
    # the same through device-resident vectors (with a forced slab schedule this is the path whose neighbour barriers are split
    # into signal / collect halves around the ghost-touching slabs)
    from ssc_b200 import DeviceVec
    dx, dy = DeviceVec(pc, x.size), DeviceVec(pc, x.size)
    dx.set(x)
    for _ in range(3):
        pc.apply(dx, dy)
    pc.synchronize()
    ydev = dy.get()
    dev_diff = float(np.abs(ydev - y).max() / max(np.abs(y).max(), 1e-300))
    assert dev_diff < 1e-13, "device-resident apply differs from the host-space apply: %g" % dev_diff
    parts = allgather((slab.global_ids[:slab.nowned], y))
    ok = True
    if rank == 0:
        from helpers import make_oracle
        mesh = plex.TensorPlex(tuple(cells), lengths=tuple(lengths))
        V = fem.FunctionSpace(mesh, p)
        disc = fem.Discretisation([V])
        r, ci, v = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
        o = make_oracle(disc, sp.csr_matrix((v, ci, r), shape=(V.node_count,) * 2), partition_of_unity=pou)
        lat = np.empty(V.node_count, dtype=np.int64)
        lat[V.lattice_node.ravel()] = np.arange(V.node_count)
        yref = np.empty(V.node_count)
        yref[lat] = o.apply(xlat[lat])
        ygpu = np.full(V.node_count, np.nan)
        for gid, yy in parts:
            ygpu[gid] = yy
        err = np.linalg.norm(ygpu - yref) / np.linalg.norm(yref)
        ok = bool(err < 1e-12)
        print("mgpu_check ranks=%d cells=%d^2x%d p=%d pou=%d exchange=%s host_pipeline=%d axis=%d relerr=%.3e %s" % (nranks, c, c * nranks, p, pou, "p2p" if p2p else "nccl", slabs, axis, err, "OK" if ok else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
