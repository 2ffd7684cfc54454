"""Mesh partitions for one-process-per-GPU runs: who owns which patch, and the owner<->ghost dof map.

The reference shards by mesh partition: a rank builds patches only around the points it owns
(ssc/libssc.c:500-521), needs one layer of vertex overlap so every owned star is local
(tests/test_jacobi_equivalent.py:9), fills ghost values before the patch loop (SFBcast, :1323-1324) and sums
overlap contributions into the owners after it (SFReduce, :1399-1400).  PetscSF is replaced by explicit
per-neighbour index lists (``star_forest``), handed to SSCPatchSetStarForest.
"""
import numpy as np

from . import fem, plex
from .plex import IntType


class Brick(object):
    """Rank ``rank`` of a ``grid = (gx, gy, gz)`` brick partition of a structured (cx, cy, cz) box of tensor cells, degree p.

    Vertex planes are dealt out in contiguous ranges along every axis; a rank owns the vertices of its index box and
    every mesh entity whose lowest vertex it owns, builds patches only around owned vertices (ssc/libssc.c:500-521), and
    its local mesh holds every cell that touches an owned vertex (one layer of vertex overlap,
    tests/test_jacobi_equivalent.py:9).  Local dofs are numbered owned first, then ghosts.  ``global_ids`` maps local
    node -> node of the global lattice (used only to match ghosts with their owners at set-up).  Ranks are numbered with
    the last axis fastest.  A 2 x 2 x 2 grid gives every rank 7 neighbours (faces, edges, the corner); slabs are the
    special case of a grid that is 1 along two axes.
    """

    def __init__(self, cells, p, rank, grid):
        cells = tuple(int(c) for c in cells)
        grid = tuple(int(g) for g in grid)
        nranks = int(np.prod(grid))
        if not 0 <= rank < nranks:
            raise ValueError("rank %d outside the %s grid" % (rank, grid))
        coord = np.unravel_index(rank, grid)
        lo, own_lo, own_hi, local = [], [], [], []
        for a in range(3):
            ca, ga, ra = cells[a], grid[a], int(coord[a])
            if ga > ca:
                raise ValueError("more ranks than cell layers along axis %d" % a)
            k0 = ra * ca // ga
            k1 = (ra + 1) * ca // ga if ra < ga - 1 else ca + 1       # owned vertex planes [k0, k1)
            l, h = max(k0 - 1, 0), min(k1 - 1, ca - 1)                  # local cell layers [l, h]
            lo.append(l); local.append(h - l + 1); own_lo.append(k0 - l); own_hi.append(k1 - 1 - l)
        h = 1.0 / min(cells)                                        # cubic cells whose size does not depend on the partition
        self.h = h
        mesh = plex.TensorPlex(tuple(local), lengths=tuple(c * h for c in local))
        mesh.mark_ghost_box(own_lo, own_hi)
        V = fem.FunctionSpace(mesh, p)
        self.mesh, self.space, self.rank, self.nranks, self.grid, self.coord = mesh, V, rank, nranks, grid, tuple(int(c) for c in coord)
        self.nowned = V.owned_node_count
        L = list(np.indices(V.lattice_shape).reshape(3, -1))
        for a in range(3):
            L[a] = L[a] + p * lo[a]
        gshape = tuple(p * c + 1 for c in cells)
        gid = np.ravel_multi_index(tuple(L), gshape)
        self.global_ids = np.empty(V.node_count, dtype=IntType)
        self.global_ids[V.lattice_node.ravel()] = gid
        onb = np.zeros(gid.size, dtype=bool)
        for a in range(3):
            onb |= (L[a] == 0) | (L[a] == gshape[a] - 1)
        bc_nodes = np.sort(V.lattice_node.ravel()[onb])
        self.discretisation = fem.Discretisation([V], bc_nodes=[bc_nodes], owned_nodes=[self.nowned])
        self.global_shape = gshape
        # vertex planes at which this partition cuts the mesh, per axis (the first plane owned by every rank but the first)
        self.cuts = tuple(tuple(r * cells[a] // grid[a] for r in range(1, grid[a])) for a in range(3))


class Slab(Brick):
    """Rank ``rank`` of ``nranks`` slabs cut along ``axis``: a Brick partition with a grid that is 1 along the other axes.

    ``axis = 0`` (the slowest axis of the cell, patch and dof numbering) keeps a rank's interface dofs and the patches
    that touch ghosts contiguous at the two ends of its numbering -- what the overlapped host-space apply wants;
    ``axis = 2`` scatters them through the whole numbering (every x-range of the slab touches the interface).
    """

    def __init__(self, cells, p, rank, nranks, axis=2):
        grid = [1, 1, 1]
        grid[axis] = nranks
        Brick.__init__(self, cells, p, rank, tuple(grid))
        self.axis = axis


class ZSlab(Slab):
    def __init__(self, cells, p, rank, nranks):
        Slab.__init__(self, cells, p, rank, nranks, axis=2)


class XSlab(Slab):
    def __init__(self, cells, p, rank, nranks):
        Slab.__init__(self, cells, p, rank, nranks, axis=0)


def star_forest(global_ids, nowned, rank, nranks, allgather):
    """Per-neighbour send/receive lists from local->global dof ids.

    ``allgather(obj)`` returns the list of every rank's ``obj`` (torch.distributed.all_gather_object in real runs).
    Returns (self_leaf, self_root, neigh, send_off, send_root, recv_off, recv_leaf) in the layout of
    SSCPatchSetStarForest: the ghosts a rank receives from neighbour q appear in the order of that rank's own ghost
    list, and the owner lists its send indices in the same order.
    """
    global_ids = np.asarray(global_ids, dtype=IntType)
    nlocal = global_ids.size
    ghost_gid = global_ids[nowned:]
    all_ghosts = allgather(ghost_gid)
    owned_gid = global_ids[:nowned]
    order = np.argsort(owned_gid, kind="stable")
    sorted_owned = owned_gid[order]
    found = {}
    send = {}
    for q in range(nranks):
        if q == rank or all_ghosts[q].size == 0:
            continue
        pos = np.searchsorted(sorted_owned, all_ghosts[q])
        pos[pos >= nowned] = max(nowned - 1, 0)
        hit = sorted_owned[pos] == all_ghosts[q] if nowned else np.zeros(all_ghosts[q].size, dtype=bool)
        if hit.any():
            found[q] = hit
            send[q] = order[pos[hit]].astype(IntType)
    all_found = allgather(found)
    recv = {}
    for q in range(nranks):
        if q != rank and rank in all_found[q]:
            recv[q] = (nowned + np.nonzero(all_found[q][rank])[0]).astype(IntType)
    claimed = np.zeros(nlocal - nowned, dtype=np.int64)
    for q, idx in recv.items():
        claimed[idx - nowned] += 1
    if nlocal > nowned and not np.all(claimed == 1):
        raise ValueError("rank %d: %d ghost dofs have no unique owner" % (rank, int((claimed != 1).sum())))
    neigh = sorted(set(send) | set(recv))
    send_off, recv_off = [0], [0]
    send_root, recv_leaf = [], []
    for q in neigh:
        s = send.get(q, np.zeros(0, dtype=IntType))
        r = recv.get(q, np.zeros(0, dtype=IntType))
        send_root.append(s)
        recv_leaf.append(r)
        send_off.append(send_off[-1] + s.size)
        recv_off.append(recv_off[-1] + r.size)
    cat = lambda xs: np.concatenate(xs).astype(IntType) if xs else np.zeros(0, dtype=IntType)   # noqa: E731
    ident = np.arange(nowned, dtype=IntType)
    return (ident, ident.copy(), np.array(neigh, dtype=np.int32), np.array(send_off, dtype=IntType), cat(send_root),
            np.array(recv_off, dtype=IntType), cat(recv_leaf))


def simulate_allgather(objs_by_rank):
    """In-process stand-in for all_gather_object: call with the list of per-rank objects."""
    return lambda obj, _all=objs_by_rank: list(_all)


def connect_peers(pc, sf, rank, allgather):
    """Switch ``pc`` (star forest ``sf`` already set) to the NVLink peer exchange: gather every rank's IPC handles and
    index lists, pick out what concerns this rank's neighbours, and call SSCPatchPeerConnect."""
    _, _, neigh, send_off, send_root, recv_off, recv_leaf = sf
    nowned = int(sf[0].size)
    mine = {"handles": pc.peerExport(), "neigh": [int(q) for q in neigh], "send_off": send_off, "send_root": send_root,
            "recv_off": recv_off, "recv_leaf": recv_leaf, "nowned": nowned}
    world = allgather(mine)
    handles, slots, rleaf, rroot = [], [], [], []
    for q in mine["neigh"]:
        other = world[q]
        kq = other["neigh"].index(rank)
        handles.append(other["handles"])
        slots.append(kq)
        rleaf.append(other["recv_leaf"][other["recv_off"][kq]:other["recv_off"][kq + 1]] - other["nowned"])   # ghost slot numbers
        rroot.append(other["send_root"][other["send_off"][kq]:other["send_off"][kq + 1]])
    cat = lambda xs: np.concatenate(xs).astype(IntType) if xs else np.zeros(0, dtype=IntType)   # noqa: E731
    pc.peerConnect(handles, slots, cat(rleaf), cat(rroot))
