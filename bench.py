#!/usr/bin/env python
"""bench.py -- PCApply throughput of the vertex-star patch preconditioner (BASELINE.json metric).

Workload (config.workload): 3D Q3 Poisson on a 64^3 hex mesh per GPU, vertex-star patches, additive, fp64
(BASELINE.json configs[2]).  With N GPUs the mesh is (64*N, 64, 64) split into N x-slabs (weak scaling; x is the slowest
axis of the numbering, so a rank's interface dofs are contiguous at the ends of its vectors): every rank
builds its own slab (owned vertices + one cell layer of ghosts), patches never cross ranks, and the only exchange is
the halo fill before and the overlap sum after the patch kernel (SFBcast/SFReduce of ssc/libssc.c:1323, :1399).

A "step" is one PCApply.  value = dofs of the global problem / max-over-ranks device time per step, vectors resident in
HBM.  e2e = the same through PatchPC.apply with pinned HOST vectors (H2D of x and D2H of y inside the timed region).
roofline = algorithmic bytes of SURVEY.md 8(d) / device time of the fused apply kernel(s), CUDA events on the PC's stream.
Inputs (31.7 GB of patch blocks) are far larger than the 126 MB L2, so no L2 flush is needed between steps.

--impl reference times the CPU restatement of libssc's PCApply loop (oracle/, all host threads) on a bounded sample
of the same workload: the reference itself needs PETSc and cannot be built here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PCApply dofs/s (3D Q3 Poisson vertex-star patches, additive, fp64)"


# ----------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ssc_b200", choices=["ssc_b200", "reference"])
    ap.add_argument("--cells", type=int, default=64, help="hex cells per direction per GPU")
    ap.add_argument("--degree", type=int, default=3)
    ap.add_argument("--workload", default="poisson_hex", choices=["poisson_hex", "elasticity_tets"],
                    help="poisson_hex = BASELINE.json configs[2] (default); elasticity_tets = configs[3]: vector P2 linear elasticity on a Freudenthal tet mesh (N=1 only)")
    ap.add_argument("--operator", default="csr", choices=["csr", "elements"],
                    help="source of the patch operators: the assembled CSR (extract kernel, default) or element matrices (the reference's own route; hex workloads, no CSR is built: the CPU baseline and the packed-symmetric extra are skipped)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-setup-serial", action="store_true", help="CPU baseline: factorise the window's patch operators on one core too (the reference's set-up is serial per rank); adds ~10 s")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak (default): --cells^3 per GPU; strong: ONE --cells^3 mesh split over the GPUs")
    ap.add_argument("--partition", default="slab", choices=["slab", "brick"], help="multi-GPU partition: x-slabs (default) or bricks (2x1x1, 2x2x1, 2x2x2)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the extra single-GPU measurements of BASELINE.json configs[3] (elasticity on tets) and configs[4] (Q4) that the default N=1 run appends")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-rank parity leg (oracle on the patches around every rank interface)")
    ap.add_argument("--apply-variant", type=int, default=None, help="ssc_apply_variant (kernel selection; default = library default)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="multi-GPU halo exchange: NVLink peer memory (default) or NCCL send/recv")
    ap.add_argument("--factor-panel", type=int, default=None)
    ap.add_argument("--no-spd-packed", action="store_true", help="skip the extra measurements at N=1: packed-symmetric (cholesky) storage, the refined sub-solve and the ssc.SSC composite (patch + P1 coarse PC)")
    ap.add_argument("--deterministic", action="store_true", help="ssc_deterministic: slot-and-sum instead of fp64 atomics")
    ap.add_argument("--factor-ctas", type=int, default=None)
    ap.add_argument("--host-pipeline", type=int, default=None, help="ssc_host_pipeline (slabs of the overlapped host-space apply)")
    ap.add_argument("--pipeline-taper", type=int, default=None, help="ssc_host_pipeline_taper (0: equal slabs, 1: slabs growing from both ends)")
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def bind_near_gpu(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, before any pinned host memory is allocated, so
    the staging buffers of the host<->device copies sit on the GPU's own NUMA node (8 ranks sharing one node's memory
    controllers halve each other's PCIe throughput).  Best effort: silently does nothing where NVML or the CPU set is
    unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i for i in range(64 * words) if (int(mask[i // 64]) >> (i % 64)) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def build_problem(args, rank, nranks):
    """Local slab of the (c*nranks, c, c) Q_p problem: discretisation, CSR provider, SF description."""
    from ssc_b200 import fem, plex
    from ssc_b200 import partition
    t0 = time.time()
    c, p = args.cells, args.degree
    if args.workload == "elasticity_tets":
        if nranks != 1:
            raise SystemExit("elasticity_tets is a single-GPU workload")
        mesh = plex.BoxMesh(c, c, c)
        V = fem.FunctionSpace(mesh, 2, 3)
        disc = fem.Discretisation([V])
        return V, disc, None, {"t_mesh": time.time() - t0}
    if nranks == 1:
        mesh = plex.TensorPlex((c, c, c))
        V = fem.FunctionSpace(mesh, p)
        disc = fem.Discretisation([V])
        slab = None
    else:
        # slabs are cut along x, the slowest axis of the numbering: a rank's interface dofs are contiguous at the ends of its vectors
        grid = {"slab": (nranks, 1, 1), "brick": {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(nranks, (nranks, 1, 1))}[args.partition]
        cells = (c, c, c) if args.scaling == "strong" else tuple(c * g for g in grid)
        slab = partition.Brick(cells, p, rank, grid)
        V, disc = slab.space, slab.discretisation
    info = {"t_mesh": time.time() - t0}
    return V, disc, slab, info


def operator_csr(V, disc):
    from ssc_b200 import synth
    if V.mesh.cell_type == "simplex":
        from ssc_b200 import fem
        A, _ = fem.assemble_elasticity_simplex(V)
        A = fem.apply_bcs(A, disc.ghost_bc_nodes)
        return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
    synth.set_threads(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))   # torchrun pins OMP_NUM_THREADS=1
    return synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)


def window_oracle(args, nthreads):
    """The CPU oracle on a 6-cell-thick window (6, c, c) of the hex workload, everything built by oracle/ and the synthetic
    FE generator alone (no product library on this path): mesh, operator, dof lists (oracle restatement of
    ssc/libssc.c:452-900), LU factors.  The window's vertex planes 1..5 carry patches identical to those of any interior
    plane of the full mesh, planes 0 and 6 those of its two boundary planes.  Returns (oracle, V, times)."""
    from oracle import oracle
    from ssc_b200 import fem, plex, synth
    c, p = args.cells, args.degree
    h = 1.0 / c
    t0 = time.perf_counter()
    mesh = plex.TensorPlex((min(6, c), c, c), lengths=(min(6, c) * h, c * h, c * h))
    V = fem.FunctionSpace(mesh, p)
    disc = fem.Discretisation([V])
    synth.set_threads(max(1, nthreads))
    csr = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
    t1 = time.perf_counter()
    o = oracle.OraclePatchPC()
    o.setFromDiscretisation(disc)
    o.setOptions()
    o.setUpPatches()
    t2 = time.perf_counter()
    o.setOperatorCSRArrays(*csr)
    o.setUpOperators(nthreads=1 if args.cpu_setup_serial else nthreads)
    t3 = time.perf_counter()
    return o, V, {"mesh_and_operator_s": t1 - t0, "patch_lists_s": t2 - t1, "operators_and_lu_s": t3 - t2,
                  "operators_and_lu_threads": 1 if args.cpu_setup_serial else nthreads}, csr


def hex_census(c, p):
    """(number of patches, sum n, sum n^2) of the c^3 Q_p vertex-star problem with Dirichlet faces (sizes are
    products of the 1D sizes 2p-1 (interior vertex) and p-1 (boundary vertex))."""
    one = np.array([p - 1] + [2 * p - 1] * (c - 1) + [p - 1], dtype=np.float64)
    return (c + 1) ** 3, float(one.sum()) ** 3, float((one ** 2).sum()) ** 3


def hex_census_n3(c, p):
    one = np.array([p - 1] + [2 * p - 1] * (c - 1) + [p - 1], dtype=np.float64)
    return float((one ** 3).sum()) ** 3


def cpu_baseline(args, nthreads=None, x_lat=None, y_lat=None):
    """The reference's CPU path restated (oracle/ssc_oracle.c: gather, LU solve with precomputed factors, scatter-add --
    the patch loop of ssc/libssc.c:1342-1388) timed on the window of `window_oracle` (7 vertex planes of (c+1)^2 patches)
    and scaled to the whole mesh by sum n^2 (the loop streams the factors).  Reported on all host threads and on ONE
    core (the reference is single-threaded per rank), with the oracle's own set-up times scaled by patch count.
    x_lat / y_lat: optional global input and GPU result as lattice arrays (X, Y, Z) for the parity sample."""
    nthreads = nthreads or os.cpu_count() or 1
    c, p = args.cells, args.degree
    o, V, times, synth_csr = window_oracle(args, nthreads)
    off, _ = o.patches()
    sizes = np.diff(off).astype(np.float64)
    npatch_all, sum_n_all, sum_n2_all = hex_census(c, p)
    frac = float((sizes ** 2).sum() / sum_n2_all)
    L = np.indices(V.lattice_shape).reshape(3, -1)
    node = V.lattice_node.ravel()
    k = c // 2                                   # the window sits mid-mesh: cells [k-3, k+3)
    if x_lat is not None:
        xs = np.empty(V.node_count)
        xs[node] = x_lat[L[0] + (k - 3) * p, L[1], L[2]] if c > 6 else x_lat[L[0], L[1], L[2]]
    else:
        xs = np.random.default_rng(20260101).uniform(-1, 1, V.node_count)
    # the reference's set-up is serial per rank: time dof lists (already serial above) and extraction + LU on ONE core for a slice of
    # 1024 mid-window patches, scaled to the whole mesh by sum n^3 (the LU) -- with the one-core apply this gives its set-up + first apply
    from oracle import oracle as _oracle
    _, gtol_all = o.patches()
    mid = max(0, sizes.size // 2 - 512)
    sl = slice(mid, min(sizes.size, mid + 1024))
    o1 = _oracle.OraclePatchPC()
    o1.setPatches(off[sl.start:sl.stop + 1] - off[sl.start], gtol_all[off[sl.start]:off[sl.stop]], V.node_count)
    o1.setOperatorCSRArrays(*synth_csr)
    t0 = time.perf_counter()
    o1.setUpOperators(nthreads=1)
    lu1 = time.perf_counter() - t0
    n3_frac = float((sizes[sl] ** 3).sum() / hex_census_n3(c, p))
    times["operators_and_lu_one_core_s_full_estimate"] = lu1 / n3_frac
    yo = o.apply(xs, nthreads=nthreads)           # warm-up; also the parity sample below
    parity = None
    if y_lat is not None:
        good = (L[0] >= p) & (L[0] <= (min(6, c) - 1) * p) if c > 6 else np.ones(node.size, dtype=bool)
        yg = y_lat[L[0] + (k - 3) * p, L[1], L[2]] if c > 6 else y_lat[L[0], L[1], L[2]]
        d = np.abs(yg[good] - yo[node][good])
        scale = o.applyScale(xs)[node][good]
        err = float(d.max() / np.abs(yo[node][good]).max())
        eerr = float("inf") if np.any(d[scale == 0] != 0) else float((d[scale > 0] / scale[scale > 0]).max())
        parity = {"dofs_checked": int(good.sum()), "max_rel_err_vs_oracle": err, "max_entry_err_vs_oracle": eerr, "ok": bool(err < 1e-12 and eerr < 1e-12)}
    reps, t = 0, 0.0
    while t < 8.0 and reps < 20:
        t0 = time.perf_counter()
        o.apply(xs, nthreads=nthreads)
        t += time.perf_counter() - t0
        reps += 1
    per = t / reps
    t0 = time.perf_counter()
    o.apply(xs, nthreads=1)
    per1 = time.perf_counter() - t0
    ndofs = (c * p + 1) ** 3
    pfrac = sizes.size / npatch_all
    return {"value": ndofs / (per / frac), "unit": "dofs/s", "cores": int(nthreads), "kind": "port", "extrapolated": True,
            "sample": "window of %d x %d x %d cells mid-mesh = %d patches of %d (%.2f%% of sum n^2), %d applies, scaled by sum n^2; CPU restatement of libssc PCApply "
                      "(oracle/ssc_oracle.c) with its own dof lists and LU factors; nothing of the product library on this path" % (min(6, c), c, c, sizes.size, npatch_all, 100 * frac, reps),
            "ms_per_step_full_estimate": per / frac * 1e3,
            "one_core": {"value": ndofs / (per1 / frac), "unit": "dofs/s", "cores": 1, "ms_per_step_full_estimate": per1 / frac * 1e3,
                         "note": "the reference is single-threaded per MPI rank (SURVEY.md 8d)"},
            "setup": {"note": "oracle set-up on the window, scaled to the whole mesh by patch count (dof lists) and by sum n^3 ~ patch count (LU); "
                              "the reference factorises lazily inside the first KSPSolve (ssc/libssc.c:1374), so its set-up cost = these + the first apply",
                      "window": times, "patch_lists_s_full_estimate": times["patch_lists_s"] / pfrac,
                      "operators_and_lu_s_full_estimate": times["operators_and_lu_s"] / pfrac,
                      "operators_and_lu_one_core_s_full_estimate": times["operators_and_lu_one_core_s_full_estimate"],
                      "setup_plus_first_apply_s_full_estimate_one_core": times["patch_lists_s"] / pfrac + times["operators_and_lu_one_core_s_full_estimate"] + per1 / frac},
            "parity": parity}


def interface_parity(args, dist, rank, nranks, slab, x_owned, results):
    """Multi-rank parity against the CPU oracle (SCALE runs prove the halo fill and the overlap sum themselves).

    The reference's contract across ranks is exactness of SFBcast / SFReduce (ssc/libssc.c:1323-1324, :1399-1400): the
    result must not depend on the partition.  For every plane at which the partition cuts the mesh (vertex plane k along
    axis a: slabs have N - 1 of them, 2 x 2 x 2 bricks three) rank 0 builds the 6-cell-thick window [k-3, k+3) across the
    whole mesh on its own, runs the oracle (its own dof lists, its own LU solves) on it with the global input restricted
    to it, and compares every dof whose patches all lie within two cells of the cut (vertex planes k-2 .. k+2: their
    stars are complete inside the window, so their patch problems are those of the global mesh) -- owned dofs of all
    ranks meeting there, dofs that were ghosts on some of them, and the overlap sums in between.  `results` maps a label
    to this rank's owned result vector (device-resident apply, host-space apply)."""
    p = args.degree
    gshape = slab.global_shape
    gcells = tuple((g - 1) // p for g in gshape)
    cuts = [(a, k) for a in range(3) for k in slab.cuts[a]]
    gid = np.asarray(slab.global_ids[:slab.nowned])
    XYZ = np.unravel_index(gid, gshape)
    sel = np.zeros(gid.size, dtype=bool)
    for a, k in cuts:
        sel |= (XYZ[a] >= (k - 3) * p) & (XYZ[a] <= (k + 3) * p)
    mine = {"gid": gid[sel], "x": np.asarray(x_owned)[sel]}
    for name, y in results.items():
        mine[name] = np.asarray(y)[sel]
    parts = _allgather(dist, mine, nranks)
    if rank != 0:
        return None
    from oracle import oracle
    from ssc_b200 import fem, plex, synth
    allg = np.concatenate([q["gid"] for q in parts])
    order = np.argsort(allg)
    allg = allg[order]
    if np.any(np.diff(allg) == 0):
        raise SystemExit("interface parity: a dof is owned by two ranks")
    vals = {name: np.concatenate([q[name] for q in parts])[order] for name in ["x"] + list(results)}
    h = slab.h
    nth = os.cpu_count() or 1
    windows = {}                                   # window shape -> (oracle, V, lattice indices, node numbers): congruent windows share one set of oracle factors

    def window(shape):
        if shape not in windows:
            mesh = plex.TensorPlex(shape, lengths=tuple(c * h for c in shape))
            V = fem.FunctionSpace(mesh, p)
            disc = fem.Discretisation([V])
            synth.set_threads(nth)
            csr = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
            o = oracle.OraclePatchPC()
            o.setFromDiscretisation(disc)
            o.setOptions()
            o.setUpPatches()
            o.setOperatorCSRArrays(*csr)
            o.setUpOperators(nthreads=nth)
            windows[shape] = (o, V, np.indices(V.lattice_shape).reshape(3, -1), V.lattice_node.ravel())
        return windows[shape]
    out = {"cut_planes": len(cuts), "dofs_checked": 0, "patches_checked": 0,
           "method": "CPU oracle on the 6-cell window around every plane at which the partition cuts the mesh; dofs whose patches all lie within two cells of it"}
    worst = {name: 0.0 for name in results}
    worst_entry = {name: 0.0 for name in results}
    for a, k in cuts:
        if k < 3 or k + 3 > gcells[a]:
            raise SystemExit("interface parity: cut plane %d along axis %d is closer than 3 cells to the mesh boundary" % (k, a))
        others = [b for b in range(3) if b != a]
        o, V, L, node = window((6, gcells[others[0]], gcells[others[1]]))
        G = [None, None, None]
        G[a], G[others[0]], G[others[1]] = L[0] + (k - 3) * p, L[1], L[2]
        g = np.ravel_multi_index(tuple(G), gshape)
        pos = np.searchsorted(allg, g)
        if np.any(pos >= allg.size) or np.any(allg[np.minimum(pos, allg.size - 1)] != g):
            raise SystemExit("interface parity: the ranks did not deliver every dof around cut plane %d of axis %d" % (k, a))
        xs = np.empty(V.node_count)
        xs[node] = vals["x"][pos]
        yo = o.apply(xs, nthreads=nth)
        scale = o.applyScale(xs)
        good = node[(L[0] >= p) & (L[0] <= 5 * p)]
        out["dofs_checked"] += int(good.size)
        out["patches_checked"] += int(5 * (gcells[others[0]] + 1) * (gcells[others[1]] + 1))
        for name in results:
            yg = np.empty(V.node_count)
            yg[node] = vals[name][pos]
            d = np.abs(yg[good] - yo[good])
            worst[name] = max(worst[name], float(d.max() / np.abs(yo[good]).max()))
            sc = scale[good]
            if np.any(d[sc == 0] != 0):
                worst_entry[name] = float("inf")
            else:
                worst_entry[name] = max(worst_entry[name], float((d[sc > 0] / sc[sc > 0]).max()))
    out["max_rel_err_vs_oracle"] = worst
    out["max_entry_err_vs_oracle"] = worst_entry
    out["ok"] = bool(all(v < 1e-12 for v in worst.values()) and all(v < 1e-12 for v in worst_entry.values()))
    return out


# ----------------------------------------------------------------------------------------------
def run_reference(args, rank, nranks):
    if rank != 0:
        return
    if args.workload != "poisson_hex":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU restatement arm is wired for the poisson_hex workload (BASELINE.json configs[2]) only"}))
        return
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args) if base is None or i >= args.warmup else base
        if i >= args.warmup:
            vals.append(base["value"])
        if len(vals) >= 3:           # bounded: a step here is seconds of CPU work
            break
    v = float(np.median(vals)) if vals else base["value"]
    # dofs/s of the host cores does not depend on how many slabs the job has: at N GPUs the same throughput is quoted for
    # the N-times larger problem (the whole box's cores were already used for one slab), and a step takes N times longer
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "dofs/s", "n_gpus": args.gpus, "steps": len(vals),
           "warmup": args.warmup, "ms_per_step": (args.cells * args.degree + 1) ** 3 * (1 if args.scaling == "strong" else nranks) / v * 1e3, "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, nranks),
           "cpu_baseline": dict(base, value=v),
           "e2e": {"value": v, "unit": "dofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference (ssc/libssc.c) needs PETSc and cannot be built here; this is the CPU restatement (oracle port) on the host cores"}
    print(json.dumps(out))


def workload_config(args, nranks):
    c, p = args.cells, args.degree
    if args.workload == "elasticity_tets":
        return {"workload": "3D linear elasticity, vector P2 on %dx%dx%d cubes split into tets, vertex-star patches (n = 45 interior), additive" % (c, c, c),
                "cells_per_gpu": 6 * c ** 3, "degree": 2, "parallelism": "single GPU", "exchange": None,
                "l2": "L2 flushed between timed steps" }
    strong = args.scaling == "strong" and nranks > 1
    grid = {"slab": (nranks, 1, 1), "brick": {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(nranks, (nranks, 1, 1))}[args.partition]
    cells = (c, c, c) if strong or nranks == 1 else tuple(c * g for g in grid)
    per_gpu = c ** 3 // nranks if strong else c ** 3
    note = "" if nranks == 1 else " (%s %dx%dx%d, %s)" % ("x-slabs" if args.partition == "slab" else "bricks", grid[0], grid[1], grid[2],
                                                         "one mesh split over the GPUs" if strong else "%d^3 per GPU" % c)
    return {"workload": "3D Q%d Poisson, %dx%dx%d hex cells%s, vertex-star patches, additive, stored inverse (sub_pc_type lu)" % (p, cells[0], cells[1], cells[2], note),
            "cells_per_gpu": per_gpu, "degree": p, "parallelism": "patches by mesh %s %dx%dx%d" % (args.partition, grid[0], grid[1], grid[2]), "exchange": getattr(args, "exchange_used", None),
            "operator": "assembled CSR -> extract kernel" if args.operator == "csr" else "one shared element matrix -> assemble kernel (no CSR)",
            "l2": "inputs larger than L2 (patch blocks %.1f GB per GPU), no flush" % (8 * ((2 * p - 1) ** 3) ** 2 * (c - 1) ** 3 / 1e9 / (nranks if strong else 1))}


def main():
    args = parse()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` without a launcher: start the N ranks ourselves, exactly as the contract's torchrun line does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != nranks:
        raise SystemExit("--gpus %d but the launcher started %d rank(s)" % (args.gpus, nranks))
    if args.impl == "reference":
        run_reference(args, rank, nranks)
        return
    ncpus_bound = bind_near_gpu(local_rank) if nranks > 1 else 0
    dist = None
    if nranks > 1:
        import torch.distributed as dist          # rendezvous + barrier only (gloo); the data path is the library's own NCCL
        dist.init_process_group("gloo", rank=rank, world_size=nranks)

    from ssc_b200 import PC, DeviceVec, PinnedArray, partition
    from ssc_b200.patch import setup_patch_pc
    from ssc_b200.pcpatch import get_unique_id

    V, disc, slab, info = build_problem(args, rank, nranks)
    t0 = time.time()
    elements = args.operator == "elements"
    csr = None if elements else operator_csr(V, disc)
    info["t_assemble"] = time.time() - t0

    pc = PC(device=local_rank)
    exchange = None
    if nranks > 1:                               # N = 1: the library splits the cores between the patch builder and the upload itself
        pc.setOption("ssc_builder_threads", max(1, min(32, (os.cpu_count() or 1) // nranks)))
    setup_patch_pc(pc, disc, None)
    pc._compute_op = None
    if elements:
        from ssc_b200 import fem
        if V.mesh.cell_type == "simplex":
            pc.setElementMatrices(fem.elasticity_element_matrices(V))   # one matrix per tet
        else:
            pc.setElementMatrices(fem.element_matrix_tensor(V))      # all cells of the tensor mesh are congruent: one shared matrix
    else:
        pc.setOperatorCSR(*csr)
    if args.apply_variant is not None:
        pc.setOption("ssc_apply_variant", args.apply_variant)
    if args.deterministic:
        pc.setOption("ssc_deterministic", True)
    if args.factor_ctas is not None:
        pc.setOption("ssc_factor_ctas", args.factor_ctas)
    if args.factor_panel is not None:
        pc.setOption("ssc_factor_panel", args.factor_panel)
    if args.host_pipeline is not None:
        pc.setOption("ssc_host_pipeline", args.host_pipeline)
    if args.pipeline_taper is not None:
        pc.setOption("ssc_host_pipeline_taper", bool(args.pipeline_taper))
    if nranks > 1:
        sf = partition.star_forest(slab.global_ids, slab.nowned, rank, nranks,
                                   lambda obj: _allgather(dist, obj, nranks))
        pc.setStarForest(slab.nowned, *sf)
        exchange = args.exchange
        if exchange == "p2p":
            try:
                partition.connect_peers(pc, sf, rank, lambda obj: _allgather(dist, obj, nranks))
                ok = 1
            except Exception as exc:                      # CUDA IPC not permitted in this container: say so and use NCCL
                ok = 0
                sys.stderr.write("rank %d: peer exchange unavailable (%s), using NCCL\n" % (rank, exc))
            if min(_allgather(dist, ok, nranks)) == 0:
                exchange = "nccl"
                if ok:
                    raise RuntimeError("peer exchange connected on this rank but not on all ranks")
        if exchange == "nccl":
            uid = [get_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            pc.commInit(nranks, rank, uid[0])
    args.exchange_used = exchange if nranks > 1 else None
    t0 = time.time()
    pc.setUp()
    pc.synchronize()
    info["t_setup_wall"] = time.time() - t0
    first = {k: pc.eventTime(k) for k in ("PCPATCHCreate", "PCPATCHComputeOp", "PCPATCHSolve", "SSCSetupLists", "SSCSetupUploadOperator", "SSCSetupAllocBlocks", "SSCSetupKernels", "SSCSetupFreeOperator", "SSCListsSort", "SSCListsPack", "SSCListsUpload", "SSCListsTail")}
    # the second set-up (new values, same shape: what every Newton step of a nonlinear solve pays, PatchPC.update ssc/patch.py:269-270)
    if not elements:
        pc.setOperatorValues(csr[2])          # same nonzero pattern: only the values travel again
    t0 = time.time()
    pc.setUp()
    pc.synchronize()
    info["t_resetup_wall"] = time.time() - t0
    resetup = {"resetup_wall_s": info["t_resetup_wall"], "resetup_upload_operator_ms": pc.eventTime("SSCSetupUploadOperator"), "resetup_lists_ms": pc.eventTime("SSCSetupLists"),
               "resetup_kernels_wall_ms": pc.eventTime("SSCSetupKernels"), "resetup_extract_ms": pc.eventTime("PCPATCHComputeOp"), "resetup_factor_ms": pc.eventTime("PCPATCHSolve")}
    if nranks > 1:
        pc.releaseOperator()         # the CSR was only borrowed until SetUp returned; N=1 keeps it for the CPU baseline
        csr = None
    setup = {"note": "setup_wall_s = SSCPatchSetUp with the CSR in pageable host memory: patch construction on the host, H2D of dof lists and CSR, extract, factor" if not elements else "setup_wall_s = SSCPatchSetUp from one shared element matrix: patch construction on the host, H2D of dof lists / cell lists / cell-node maps, assemble kernel (extract_ms), factor", "create_patches_ms": first["PCPATCHCreate"], "extract_ms": first["PCPATCHComputeOp"],
             "factor_ms": first["PCPATCHSolve"], "lists_ms": first["SSCSetupLists"], "upload_operator_ms": first["SSCSetupUploadOperator"],
             "lists_sort_ms": first["SSCListsSort"], "lists_pack_ms": first["SSCListsPack"], "lists_upload_ms": first["SSCListsUpload"], "lists_tail_ms": first["SSCListsTail"],
             "alloc_blocks_ms": first["SSCSetupAllocBlocks"], "kernels_wall_ms": first["SSCSetupKernels"], "free_operator_ms": first["SSCSetupFreeOperator"], "setup_wall_s": info["t_setup_wall"],
             "assemble_wall_s": info["t_assemble"], "mesh_wall_s": info["t_mesh"]}
    setup.update(resetup)
    if setup["extract_ms"] > 0 and not elements:
        setup["extract_gbs"] = pc.size("bytes_extract") / (setup["extract_ms"] * 1e-3) / 1e9      # algorithmic bytes of SURVEY.md 8(d)
    if setup["factor_ms"] > 0:
        setup["factor_tflops"] = pc.size("flops_factor") / (setup["factor_ms"] * 1e-3) / 1e12     # 2 n^3 per block (fp64)
    nowned = pc.size("nowned")
    nglobal = nowned
    if nranks > 1:
        tot = [None] * nranks
        dist.all_gather_object(tot, nowned)
        nglobal = int(sum(tot))

    rng = np.random.default_rng(20260101 + rank)
    xh = PinnedArray(pc, nowned)
    yh = PinnedArray(pc, nowned)
    xh.array[:] = rng.uniform(-1, 1, nowned)
    dx, dy = DeviceVec(pc, nowned), DeviceVec(pc, nowned)
    dx.set(xh.array)

    def barrier():
        pc.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        pc.apply(dx, dy)
    barrier()
    e0, e1 = pc.eventCreate(), pc.eventCreate()
    launches0 = pc.size("gpu_launches")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    pc.kernelTimerReset(True)
    barrier()
    pc.eventRecord(e0)
    flush = args.workload != "poisson_hex" or pc.size("block_bytes") < (1 << 30)     # small working sets: flush L2 between steps
    if flush:
        ms = 0.0
        for _ in range(args.steps):
            pc.flushL2()
            pc.eventRecord(e0)
            pc.apply(dx, dy)
            pc.eventRecord(e1)
            ms += pc.eventElapsed(e0, e1)
    else:
        for _ in range(args.steps):
            pc.apply(dx, dy)
        pc.eventRecord(e1)
        ms = pc.eventElapsed(e0, e1)
    barrier()
    kms, kn = pc.kernelTimerRead()
    pc.kernelTimerReset(False)
    launches = pc.size("gpu_launches") - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_all = _max_over_ranks(dist, ms, nranks)
    ms_step = ms_all / args.steps

    # ---- end to end through the python PC surface with host vectors -------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(args.warmup):
            pc.apply(xh.array, yh.array)
        barrier()
        pc.eventRecord(e0)
        for _ in range(args.steps):
            pc.apply(xh.array, yh.array)
        pc.eventRecord(e1)
        ems = _max_over_ranks(dist, pc.eventElapsed(e0, e1), nranks) / args.steps
        barrier()
        e2e = {"value": nglobal / (ems * 1e-3), "unit": "dofs/s", "h2d_bytes_per_step": 8 * nowned, "d2h_bytes_per_step": 8 * nowned,
               "ms_per_step": ems, "api": "ssc_b200.PC.apply(x, y) on pinned host vectors (SSCPatchApply, SSC_MEM_HOST)", "cpus_bound_near_gpu": ncpus_bound}

    # ---- the same with PAGEABLE caller vectors (what a PETSc Vec hands over): the library page-locks them on first use ----
    e2e_pageable = None
    if not args.no_e2e:
        xp, yp = np.array(xh.array), np.empty(nowned)
        for _ in range(args.warmup):
            pc.apply(xp, yp)
        barrier()
        pc.eventRecord(e0)
        for _ in range(args.steps):
            pc.apply(xp, yp)
        pc.eventRecord(e1)
        pms = _max_over_ranks(dist, pc.eventElapsed(e0, e1), nranks) / args.steps
        barrier()
        e2e_pageable = {"value": nglobal / (pms * 1e-3), "unit": "dofs/s", "ms_per_step": pms, "host_vectors_registered": pc.size("host_vectors_registered"),
                        "max_abs_diff_vs_pinned": float(np.abs(yp - yh.array).max()),
                        "api": "ssc_b200.PC.apply(x, y) on plain numpy (pageable) vectors; cudaHostRegister on first use, cached by address"}

    # ---- multi-rank parity leg: the oracle on the patches around every rank interface -----------------------
    parity = None
    if nranks > 1 and args.workload == "poisson_hex" and not args.no_parity:
        res = {"device_resident": dy.get()}
        if e2e is not None:
            res["host_vectors"] = np.array(yh.array)
        parity = interface_parity(args, dist, rank, nranks, slab, xh.array, res)
        dist.barrier()          # rank 0 spent seconds in the oracle: the neighbour barriers of the next applies spin for 4 s at most

    # size-independent property at full size (all N): the additive operator is symmetric for an SPD problem, so
    # <M a, b> == <a, M b> over the whole (distributed) vector; a broken halo fill or overlap sum breaks it
    rng2 = np.random.default_rng(777 + rank)
    a_h, b_h = rng2.uniform(-1, 1, nowned), rng2.uniform(-1, 1, nowned)
    da, db, dMa, dMb = (DeviceVec(pc, nowned) for _ in range(4))
    da.set(a_h)
    db.set(b_h)
    pc.apply(da, dMa)
    pc.apply(db, dMb)
    pc.synchronize()
    Ma, Mb = dMa.get(), dMb.get()
    bc = disc.global_bc_nodes
    free = np.ones(nowned, dtype=bool)
    free[bc[bc < nowned]] = False                      # Dirichlet rows are passed through, not part of the symmetric operator
    dots = np.array([float(Ma[free] @ b_h[free]), float(a_h[free] @ Mb[free])])
    if dist is not None:
        parts = _allgather(dist, dots, nranks)
        dots = np.sum(parts, axis=0)
    symmetry = {"<Ma,b>": float(dots[0]), "<a,Mb>": float(dots[1]), "rel_diff": float(abs(dots[0] - dots[1]) / max(abs(dots[0]), 1e-300))}
    for v in (da, db, dMa, dMb):
        v.free()
    if rank != 0:
        return
    if symmetry["rel_diff"] > 1e-10:
        raise SystemExit("symmetry check failed: %r" % (symmetry,))
    if parity is not None and not parity["ok"]:
        raise SystemExit("multi-rank PCApply differs from the CPU oracle around the rank interfaces: %r" % (parity,))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_apply = pc.size("bytes_apply")
    kernel_ms = kms / max(kn, 1)
    achieved = bytes_apply / (kernel_ms * 1e-3) / 1e9
    read_peak = pc.measureReadBandwidth(5)      # plain streaming read of the same blocks on this very GPU
    traffic = None
    try:            # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this workload (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "apply_traffic.json")))
        if tj.get("cells") == args.cells and tj.get("degree") == args.degree:
            traffic = tj["dram_bytes_per_apply"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "ssc_apply_pipe_kernel<1,8> (+ ssc_apply_stream_kernel<8> for the small classes: all size-class launches of one apply; traffic = ncu capture of the dominant launch)", "kernel_ms": kernel_ms,
                "algorithmic_bytes": bytes_apply, "stream_read_gbs_this_gpu": read_peak, "frac_of_stream_read": achieved / read_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
    out = {"metric": METRIC.replace("Q3", "Q%d" % args.degree) if args.workload == "poisson_hex" else "PCApply dofs/s (3D vector P2 elasticity on tets, vertex-star patches, additive, fp64)", "value": nglobal / (ms_step * 1e-3), "unit": "dofs/s", "n_gpus": nranks, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling if nranks > 1 else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args, nranks),
           "dofs": nglobal, "patches_per_gpu": pc.size("npatch"), "hbm_gbs_step": bytes_apply * nranks / (ms_step * 1e-3) / 1e9,
           "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "setup": setup, "symmetry_check": symmetry}
    if nranks > 1:
        out["exchange_ms_per_apply"] = ms_step - kernel_ms      # halo push, the two neighbour barriers and the fold: whatever of a step is not the patch kernels (rank 0's kernels, max-over-ranks step)
    if parity is not None:
        out["parity"] = parity
    if e2e_pageable is not None:
        out["e2e_pageable"] = e2e_pageable
    if flush:
        out["config"]["l2"] = "working set %.2f GB: L2 flushed (256 MB write) before every timed step, each step timed on its own" % (pc.size("block_bytes") / 1e9)
    if nranks == 1 and not args.no_spd_packed and args.workload == "poisson_hex" and not elements:
        try:
            out["spd_packed"] = spd_packed_line(args, disc, csr, dx, dy, peak)
        except Exception as exc:                     # an extra, never allowed to take the main line down
            out["spd_packed"] = {"error": str(exc)}
    if nranks == 1 and not args.no_spd_packed and args.workload == "poisson_hex" and not elements:
        try:
            out["sub_solve_refine"] = refine_line(args, disc, csr, dx, dy, peak)
        except Exception as exc:
            out["sub_solve_refine"] = {"error": str(exc)}
    if nranks == 1 and not args.no_spd_packed and args.workload == "poisson_hex":
        try:
            out["ssc_composite"] = composite_line(args, V, disc, pc, dx, dy)
        except Exception as exc:
            out["ssc_composite"] = {"error": str(exc)}
    if nranks == 1 and not args.no_cpu_baseline and args.workload == "poisson_hex":
        pc.apply(dx, dy)
        pc.synchronize()
        lat = V.lattice_node                       # (X, Y, Z) lattice -> dof number
        out["cpu_baseline"] = cpu_baseline(args, x_lat=np.asarray(xh.array)[lat], y_lat=dy.get()[lat])
        if out["cpu_baseline"]["parity"] and not out["cpu_baseline"]["parity"]["ok"]:
            raise SystemExit("PCApply differs from the CPU oracle on the sampled dofs: %r" % (out["cpu_baseline"]["parity"],))
    if nranks == 1 and args.workload == "poisson_hex" and args.cells == 64 and args.degree == 3 and not args.no_other_configs:
        # BASELINE.json configs[3] and configs[4] measured by the same run (single GPU, sizes that set up in seconds), each
        # through its own `bench.py` process once this one has released the device
        for v in (dx, dy):
            v.free()
        pc.destroy()
        out["other_configs"] = other_configs(args)
    print(json.dumps(out))


def other_configs(args):
    res = {}
    runs = {"elasticity_tets_p2_vector_24^3": ["--workload", "elasticity_tets", "--cells", "24", "--no-e2e"],
            "poisson_hex_q4_32^3_from_element_matrices": ["--degree", "4", "--cells", "32", "--operator", "elements", "--no-e2e"]}
    for name, extra in runs.items():
        cmd = [sys.executable, os.path.abspath(__file__), "--steps", str(max(5, min(args.steps, 20))), "--warmup", "3", "--no-cpu-baseline", "--no-spd-packed", "--no-other-configs"] + extra
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            if p.returncode != 0 or not line:
                res[name] = {"error": (p.stderr or "no output")[-400:]}
                continue
            d = json.loads(line[-1])
            res[name] = {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "dofs", "patches_per_gpu", "config", "gpu_launches")}
            res[name]["roofline"] = {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "algorithmic_bytes", "kernel_ms")}
            res[name]["setup"] = {k: d["setup"].get(k) for k in ("setup_wall_s", "extract_ms", "factor_ms", "factor_tflops")}
            res[name]["symmetry_check_rel_diff"] = d["symmetry_check"]["rel_diff"]
        except Exception as exc:                      # an extra, never allowed to take the main line down
            res[name] = {"error": str(exc)}
    return res


def spd_packed_line(args, disc, csr, dx, dy, peak):
    """Same workload with -sub_pc_type cholesky -ssc_symmetric_storage: only the lower triangle of each (symmetric)
    inverse is stored and streamed, through one bulk asynchronous copy per patch.  Its own algorithmic byte count
    (4*sum n(n+1) in place of 8*sum n^2) -- reported beside, never mixed with, the general-mode line."""
    from ssc_b200 import PC, DeviceVec
    from ssc_b200.patch import setup_patch_pc
    pc2 = PC(device=dx.pc.device)
    setup_patch_pc(pc2, disc, None)
    pc2._compute_op = None
    pc2.setOperatorCSR(*csr)
    pc2.setOption("sub_pc_type", "cholesky")
    pc2.setOption("ssc_symmetric_storage", True)
    t0 = time.time()
    pc2.setUp()
    pc2.synchronize()
    wall = time.time() - t0
    x2, y2 = DeviceVec(pc2, dx.n), DeviceVec(pc2, dx.n)
    x2.set(dx.get())
    for _ in range(args.warmup):
        pc2.apply(x2, y2)
    pc2.synchronize()
    e0, e1 = pc2.eventCreate(), pc2.eventCreate()
    pc2.eventRecord(e0)
    for _ in range(args.steps):
        pc2.apply(x2, y2)
    pc2.eventRecord(e1)
    ms = pc2.eventElapsed(e0, e1) / args.steps
    ref = dy.get()
    got = y2.get()
    nbytes = pc2.size("bytes_apply")
    res = {"ms_per_step": ms, "value": dx.n / (ms * 1e-3), "unit": "dofs/s", "algorithmic_bytes": nbytes,
           "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / peak,
           "block_bytes": pc2.size("block_bytes"), "setup_wall_s": wall, "factor_ms": pc2.eventTime("PCPATCHSolve"),
           "rel_diff_vs_general_mode": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
           "options": "-sub_pc_type cholesky -ssc_symmetric_storage true"}
    pc2.destroy()
    return res


def composite_line(args, V, disc, pc, dx, dy):
    """ssc.SSC of ssc/ssc.py:38-49 on the same workload: the patch PC of the main line + the P1 coarse PC (ssc/lo.py), summed on
    the device.  Coarse solver: one smoothed-aggregation multigrid cycle (-lo_ksp_type preonly -lo_pc_type gamg); the two PCs run
    on their own streams.  Host-timed around `steps` composite applies with device-resident vectors (a synchronize on both sides)."""
    from ssc_b200 import DeviceVec, fem, synth
    from ssc_b200.lo import LowOrderCore
    c, p = args.cells, args.degree
    V1 = fem.FunctionSpace(V.mesh, 1)
    d1 = fem.Discretisation([V1])
    r1, c1, v1 = synth.tensor_poisson_csr(V1, d1.ghost_bc_nodes)
    import scipy.sparse as sp
    A1 = sp.csr_matrix((v1, c1, r1), shape=(V1.node_count,) * 2)
    lo = LowOrderCore(dx.pc.device)
    lo.setOption("lo_ksp_type", "preonly")
    lo.setOption("lo_pc_type", "gamg")
    lo.setRestriction(synth.tensor_p1_restriction(V, V1, disc.ghost_bc_nodes, d1.ghost_bc_nodes, c, p))
    lo.setBcs(disc.global_bc_nodes)
    lo.setOperator(A1)
    t0 = time.time()
    lo.setUp()
    t_setup = time.time() - t0
    dw = DeviceVec(pc, dx.n)

    def apply():
        pc.apply(dx, dy)
        lo.applyDevice(dx, dw)
        pc.synchronize()
        lo.axpyDevice(1.0, dw, dy)

    for _ in range(args.warmup):
        apply()
    lo.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        apply()
    lo.synchronize()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    return {"ms_per_step": ms, "value": dx.n / (ms * 1e-3), "unit": "dofs/s", "coarse_dofs": int(V1.node_count), "coarse_setup_s": t_setup,
            "options": "ssc.SSC = patch PC + P1PC, -lo_ksp_type preonly -lo_pc_type gamg (one V(1,1) cycle of smoothed-aggregation multigrid)",
            "timing": "host clock around %d applies, device-resident vectors" % args.steps}


def refine_line(args, disc, csr, dx, dy, peak):
    """Same workload with -ssc_sub_solve refine: stored-inverse product + one refinement step against the kept patch
    operators, residual in double-double (the accuracy of the reference's LU solve and better, at three passes over the
    blocks).  Its own algorithmic byte count (8 * 3 sum n^2 in place of 8 sum n^2), reported beside the default line."""
    from ssc_b200 import PC, DeviceVec
    from ssc_b200.patch import setup_patch_pc
    pc2 = PC(device=dx.pc.device)
    setup_patch_pc(pc2, disc, None)
    pc2._compute_op = None
    pc2.setOperatorCSR(*csr)
    pc2.setOption("ssc_sub_solve", "refine")
    pc2.setUp()
    pc2.synchronize()
    x2, y2 = DeviceVec(pc2, dx.n), DeviceVec(pc2, dx.n)
    x2.set(dx.get())
    for _ in range(args.warmup):
        pc2.apply(x2, y2)
    pc2.synchronize()
    e0, e1 = pc2.eventCreate(), pc2.eventCreate()
    pc2.eventRecord(e0)
    for _ in range(args.steps):
        pc2.apply(x2, y2)
    pc2.eventRecord(e1)
    ms = pc2.eventElapsed(e0, e1) / args.steps
    ref, got = dy.get(), y2.get()
    nbytes = pc2.size("bytes_apply")
    res = {"ms_per_step": ms, "value": dx.n / (ms * 1e-3), "unit": "dofs/s", "algorithmic_bytes": nbytes,
           "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / peak,
           "kappa_max": pc2.eventTime("SSCKappaMax"), "max_rel_diff_vs_stored_inverse": float(np.abs(got - ref).max() / np.abs(ref).max()),
           "options": "-ssc_sub_solve refine (one step; -ssc_sub_solve auto refines only classes with kappa > 1e3: none on this workload)"}
    pc2.destroy()
    return res


def _allgather(dist, obj, nranks):
    out = [None] * nranks
    dist.all_gather_object(out, obj)
    return out


def _max_over_ranks(dist, v, nranks):
    if dist is None:
        return v
    out = [None] * nranks
    dist.all_gather_object(out, float(v))
    return max(out)


if __name__ == "__main__":
    main()
