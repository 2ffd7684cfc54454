"""The N>1 host logic on CPU: world_size-2 (and 3) `gloo` processes build their z-slabs, derive the star forest with
all_gather_object exactly as bench.py does, run the patch loop with the oracle, exchange halos over gloo, and must
reproduce the single-rank result ("oracle with k partitions must equal the 1-partition result")."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C, P = 4, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_for(disc, V):
    import scipy.sparse as sp
    from oracle.oracle import OraclePatchPC
    from ssc_b200 import synth
    o = OraclePatchPC()
    o.setFromDiscretisation(disc)
    o.setOptions()
    o.setUpPatches()
    r, c, v = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
    o.setOperatorCSR(sp.csr_matrix((v, c, r), shape=(V.node_count,) * 2))
    o.setUpOperators()
    return o


def _worker(rank, nranks, port, outdir, axis=2, parity_leg=False, grid=None, cells_override=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from ssc_b200 import partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)

    def allgather(obj):
        out = [None] * nranks
        dist.all_gather_object(out, obj)
        return out
    cells = [C, C, C]
    cells[axis] = C * nranks
    if grid is not None:
        cells = list(cells_override)
        slab = partition.Brick(tuple(cells), P, rank, grid)
    else:
        slab = partition.Slab(tuple(cells), P, rank, nranks, axis=axis)
    sl, sr, neigh, so, sroot, ro, rleaf = partition.star_forest(slab.global_ids, slab.nowned, rank, nranks, allgather)
    o = _oracle_for(slab.discretisation, slab.space)
    xlat = np.random.default_rng(7).uniform(-1, 1, int(np.prod(slab.global_shape)))     # same on every rank
    x = xlat[slab.global_ids[:slab.nowned]]
    xl = np.zeros(slab.space.node_count)
    xl[sl] = x[sr]

    def exchange(send_chunks, recv_sizes):
        reqs, bufs = [], []
        for k, q in enumerate(neigh):
            t = torch.from_numpy(np.ascontiguousarray(send_chunks[k]))
            reqs.append(dist.isend(t, int(q)))
            b = torch.zeros(recv_sizes[k], dtype=torch.float64)
            reqs.append(dist.irecv(b, int(q)))
            bufs.append(b)
        for r in reqs:
            r.wait()
        return [b.numpy() for b in bufs]
    # halo fill (SFBcast, ssc/libssc.c:1323): owners send, ghosts receive
    got = exchange([x[sroot[so[k]:so[k + 1]]] for k in range(len(neigh))], [int(ro[k + 1] - ro[k]) for k in range(len(neigh))])
    for k in range(len(neigh)):
        xl[rleaf[ro[k]:ro[k + 1]]] = got[k]
    yl = o.applyLocal(xl)
    # overlap sum (SFReduce, :1399): ghosts send, owners add
    y = np.zeros(slab.nowned)
    y[sr] += yl[sl]
    got = exchange([yl[rleaf[ro[k]:ro[k + 1]]] for k in range(len(neigh))], [int(so[k + 1] - so[k]) for k in range(len(neigh))])
    for k in range(len(neigh)):
        np.add.at(y, sroot[so[k]:so[k + 1]], got[k])
    bc = slab.discretisation.global_bc_nodes
    y[bc] = x[bc]
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), y=y, gid=slab.global_ids[:slab.nowned])
    if parity_leg:
        # bench.py's multi-rank parity leg on the same data: correct results pass, a result without the overlap sum fails
        import argparse
        import json
        import bench
        args = argparse.Namespace(cells=C, degree=P)
        y_no_sum = np.zeros(slab.nowned)
        y_no_sum[sr] += yl[sl]
        y_no_sum[bc] = x[bc]
        xg = np.array(x)
        if rank == 1:
            xg[sroot[so[0]]] += 1.0                       # "stale ghost": what the neighbour saw differs from what the owner holds
        res = bench.interface_parity(args, dist, rank, nranks, slab, x, {"good": y, "no_overlap_sum": y_no_sum})
        res2 = bench.interface_parity(args, dist, rank, nranks, slab, xg, {"stale_ghost": y})
        if rank == 0:
            json.dump({"a": res, "b": res2}, open(os.path.join(outdir, "parity.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nranks,axis", [(2, 2), (3, 2), (2, 0), (3, 0), (2, 1)])
def test_slabs_over_gloo_match_single_rank(nranks, axis, tmp_path):
    """Slabs cut along z (interface dofs scattered through the local numbering) and along x (contiguous at the ends:
    the layout the overlapped multi-rank host apply is built for)."""
    import torch.multiprocessing as mp
    from ssc_b200 import fem, plex
    mp.spawn(_worker, args=(nranks, _free_port(), str(tmp_path), axis), nprocs=nranks, join=True)
    cells, lengths = [C, C, C], [1, 1, 1]
    cells[axis], lengths[axis] = C * nranks, nranks
    mesh = plex.TensorPlex(tuple(cells), lengths=tuple(lengths))
    V = fem.FunctionSpace(mesh, P)
    disc = fem.Discretisation([V])
    o = _oracle_for(disc, V)
    lat_of_node = np.empty(V.node_count, dtype=np.int64)
    lat_of_node[V.lattice_node.ravel()] = np.arange(V.node_count)
    xlat = np.random.default_rng(7).uniform(-1, 1, V.node_count)
    y = o.apply(xlat[lat_of_node])
    ylat = np.empty_like(y)
    ylat[lat_of_node] = y
    seen = np.zeros(V.node_count, dtype=int)
    for r in range(nranks):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        seen[d["gid"]] += 1
        assert np.abs(d["y"] - ylat[d["gid"]]).max() <= 1e-14 * np.abs(ylat).max()
    assert np.all(seen == 1)          # every global dof is owned by exactly one rank


@pytest.mark.parametrize("nranks", [2, 3])
def test_bench_parity_leg_accepts_correct_and_rejects_broken_exchanges(nranks, tmp_path):
    """bench.py's N>1 parity leg (oracle on the 6-cell sub-mesh around every rank interface): correct multi-rank results
    pass at 1e-12; a result whose overlap sum was skipped, or whose halo was filled with other values than the owner's,
    is rejected."""
    import json
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(nranks, _free_port(), str(tmp_path), 0, True), nprocs=nranks, join=True)
    r = json.load(open(os.path.join(str(tmp_path), "parity.json")))
    a, b = r["a"], r["b"]
    assert a["cut_planes"] == nranks - 1 and a["dofs_checked"] == (nranks - 1) * (4 * P + 1) * (C * P + 1) ** 2
    assert a["max_rel_err_vs_oracle"]["good"] < 1e-12 and a["max_entry_err_vs_oracle"]["good"] < 1e-12
    assert a["max_rel_err_vs_oracle"]["no_overlap_sum"] > 1e-3 and not a["ok"]
    assert b["max_rel_err_vs_oracle"]["stale_ghost"] > 1e-6 and not b["ok"]


@pytest.mark.parametrize("grid,cells", [((2, 2, 1), (6, 5, 4)), ((1, 2, 2), (4, 6, 6)), ((2, 2, 2), (4, 4, 4))])
def test_bricks_over_gloo_match_single_rank(grid, cells, tmp_path):
    """Brick partitions (SURVEY.md 8e names 2 x 2 x 2 bricks): ranks with face, edge and corner neighbours; a ghost dof on a
    brick edge is owned by exactly one of the up to 7 neighbours, and the overlap sum collects contributions from all."""
    import torch.multiprocessing as mp
    from ssc_b200 import fem, plex
    nranks = int(np.prod(grid))
    mp.spawn(_worker, args=(nranks, _free_port(), str(tmp_path), 0, False, grid, cells), nprocs=nranks, join=True)
    h = 1.0 / min(cells)
    mesh = plex.TensorPlex(tuple(cells), lengths=tuple(c * h for c in cells))
    V = fem.FunctionSpace(mesh, P)
    disc = fem.Discretisation([V])
    o = _oracle_for(disc, V)
    lat_of_node = np.empty(V.node_count, dtype=np.int64)
    lat_of_node[V.lattice_node.ravel()] = np.arange(V.node_count)
    xlat = np.random.default_rng(7).uniform(-1, 1, V.node_count)
    y = o.apply(xlat[lat_of_node])
    ylat = np.empty_like(y)
    ylat[lat_of_node] = y
    seen = np.zeros(V.node_count, dtype=int)
    for r in range(nranks):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        seen[d["gid"]] += 1
        assert np.abs(d["y"] - ylat[d["gid"]]).max() <= 1e-13 * np.abs(ylat).max()
    assert np.all(seen == 1)


def test_bench_parity_leg_on_bricks(tmp_path):
    """The parity leg on a 2 x 2 x 1 brick partition: two cut planes (one along x, one along y), windows that span two
    ranks each and meet at the brick edge."""
    import json
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(4, _free_port(), str(tmp_path), 0, True, (2, 2, 1), (8, 8, 4)), nprocs=4, join=True)
    r = json.load(open(os.path.join(str(tmp_path), "parity.json")))
    a, b = r["a"], r["b"]
    assert a["cut_planes"] == 2 and a["dofs_checked"] == 2 * (4 * P + 1) * (8 * P + 1) * (4 * P + 1)
    assert a["max_rel_err_vs_oracle"]["good"] < 1e-12 and a["max_entry_err_vs_oracle"]["good"] < 1e-12
    assert a["max_rel_err_vs_oracle"]["no_overlap_sum"] > 1e-3 and not a["ok"]
    assert b["max_rel_err_vs_oracle"]["stale_ghost"] > 1e-6 and not b["ok"]
