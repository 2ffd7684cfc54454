"""DMPlex-shaped mesh topology held as plain arrays.

The reference reads its topology out of a PETSc ``DMPlex`` (``PCGetDM`` at
ssc/libssc.c:472, strata at :482-498, star/closure through
``DMPlexGetTransitiveClosure`` at :344, :348, :1460, and the ``pyop2_ghost``
label at :502).  PETSc is not available to this build, so the same
information travels as arrays: a cone table, its transpose (supports), the
depth strata and a ghost mask.  ``Plex.from_dmplex`` shows how a live
petsc4py DMPlex is flattened into this form.

Point numbering follows an interpolated DMPlex: cells first, then vertices,
then the intermediate entities from dimension ``dim-1`` down to 1.
"""
from itertools import combinations

import numpy as np

IntType = np.int64  # PetscInt is ``long`` in the reference (ssc/PatchPC.pyx:7)


def _csr_from_lists(n, rows, cols):
    """rows[i] -> cols[i] relation as (offsets, values), stable in input order."""
    order = np.argsort(rows, kind="stable")
    counts = np.bincount(rows, minlength=n)
    off = np.zeros(n + 1, dtype=IntType)
    np.cumsum(counts, out=off[1:])
    return off, np.ascontiguousarray(cols[order], dtype=IntType)


class Plex(object):
    """Topology container.  All arrays are int64 and C-contiguous."""

    def __init__(self, dim, cone_off, cones, depth_start, depth_end,
                 coords=None, ghost=None, cell_type="simplex"):
        self.dim = int(dim)
        self.cone_off = np.ascontiguousarray(cone_off, dtype=IntType)
        self.cones = np.ascontiguousarray(cones, dtype=IntType)
        self.pEnd = len(self.cone_off) - 1
        self.depth_start = np.ascontiguousarray(depth_start, dtype=IntType)
        self.depth_end = np.ascontiguousarray(depth_end, dtype=IntType)
        rows = np.repeat(np.arange(self.pEnd, dtype=IntType), np.diff(self.cone_off))
        # support(q) = points whose cone holds q
        self.supp_off, self.supports = _csr_from_lists(self.pEnd, self.cones, rows)
        self.coords = coords
        self.ghost = (np.zeros(self.pEnd, dtype=np.uint8) if ghost is None
                      else np.ascontiguousarray(ghost, dtype=np.uint8))
        self.cell_type = cell_type
        # filled by the mesh constructors, used by the function spaces
        self.cell_entities = {}
        self.entity_verts = {}

    # strata -----------------------------------------------------------
    def depth_stratum(self, depth):
        return int(self.depth_start[depth]), int(self.depth_end[depth])

    def height_stratum(self, height):
        return self.depth_stratum(self.dim - height)

    @property
    def cStart(self):
        return self.height_stratum(0)[0]

    @property
    def cEnd(self):
        return self.height_stratum(0)[1]

    @property
    def num_cells(self):
        return self.cEnd - self.cStart

    @property
    def num_vertices(self):
        s, e = self.depth_stratum(0)
        return e - s

    def point_dim(self):
        """Topological dimension of every point."""
        out = np.empty(self.pEnd, dtype=IntType)
        for d in range(self.dim + 1):
            s, e = self.depth_stratum(d)
            out[s:e] = d
        return out

    # traversal (small-scale, python; the C++ builder and the oracle have their own)
    def cone(self, p):
        return self.cones[self.cone_off[p]:self.cone_off[p + 1]]

    def support(self, p):
        return self.supports[self.supp_off[p]:self.supp_off[p + 1]]

    def closure(self, p, use_cone=True):
        seen = [int(p)]
        mark = {int(p)}
        i = 0
        while i < len(seen):
            q = seen[i]
            i += 1
            nxt = self.cone(q) if use_cone else self.support(q)
            for r in nxt:
                r = int(r)
                if r not in mark:
                    mark.add(r)
                    seen.append(r)
        return seen

    def star(self, p):
        return self.closure(p, use_cone=False)

    # boundary ---------------------------------------------------------
    def boundary_points(self):
        """Mask of points in the closure of exterior facets (Firedrake's "on_boundary")."""
        fs, fe = self.height_stratum(1)
        nsupp = np.diff(self.supp_off)[fs:fe]
        mask = np.zeros(self.pEnd, dtype=bool)
        front = np.arange(fs, fe, dtype=IntType)[nsupp == 1]
        while front.size:
            mask[front] = True
            idx = _ranges(self.cone_off[front], self.cone_off[front + 1])
            if idx.size == 0:
                break
            front = np.unique(self.cones[idx])
        return mask

    # live PETSc adapter (cannot run in this container: petsc4py is absent)
    @classmethod
    def from_dmplex(cls, dm):
        """Flatten a petsc4py ``DMPlex`` (what ``mesh._plex`` is at ssc/patch.py:214)."""
        pStart, pEnd = dm.getChart()
        assert pStart == 0
        cone_off = np.zeros(pEnd + 1, dtype=IntType)
        cones = []
        for p in range(pEnd):
            c = dm.getCone(p)
            cone_off[p + 1] = cone_off[p] + len(c)
            cones.append(np.asarray(c, dtype=IntType))
        cones = np.concatenate(cones) if cones else np.zeros(0, dtype=IntType)
        dim = dm.getDimension()
        ds = [dm.getDepthStratum(d)[0] for d in range(dim + 1)]
        de = [dm.getDepthStratum(d)[1] for d in range(dim + 1)]
        ghost = np.zeros(pEnd, dtype=np.uint8)
        if dm.hasLabel("pyop2_ghost"):
            for p in range(pEnd):
                if dm.getLabelValue("pyop2_ghost", p) >= 0:
                    ghost[p] = 1
        return cls(dim, cone_off, cones, ds, de, ghost=ghost)


def _ranges(lo, hi):
    """Concatenate arange(lo[i], hi[i]) for all i, vectorised."""
    n = hi - lo
    tot = int(n.sum())
    if tot == 0:
        return np.zeros(0, dtype=IntType)
    start = np.repeat(lo - np.concatenate(([0], np.cumsum(n)[:-1])), n)
    return (np.arange(tot, dtype=IntType) + start).astype(IntType)


# ----------------------------------------------------------------------
# simplicial meshes from cell -> vertex connectivity
# ----------------------------------------------------------------------
def _row_keys(a, base):
    key = np.zeros(a.shape[0], dtype=np.int64)
    for j in range(a.shape[1]):
        key = key * base + a[:, j]
    return key


def simplex_plex(dim, cell_verts, coords):
    """Interpolated simplicial Plex.

    ``cell_verts`` (ncell, dim+1) vertex indices.  Vertices of every cell are
    sorted ascending so that all cells see a shared edge or face with the same
    parametrisation (no per-cell orientation tables are needed).
    """
    cv = np.sort(np.asarray(cell_verts, dtype=IntType), axis=1)
    nc = cv.shape[0]
    nv = int(coords.shape[0])
    base = nv + 1
    # entity tables by dimension
    ent_verts = {0: np.arange(nv, dtype=IntType)[:, None], dim: cv}
    cell_ents = {0: cv.copy(), dim: np.arange(nc, dtype=IntType)[:, None]}
    keys = {}
    for k in range(1, dim):
        combos = list(combinations(range(dim + 1), k + 1))
        allv = np.stack([cv[:, c] for c in combos], axis=1)  # (nc, ncomb, k+1)
        flat = allv.reshape(-1, k + 1)
        key = _row_keys(flat, base)
        ukey, first, inv = np.unique(key, return_index=True, return_inverse=True)
        ent_verts[k] = flat[first]
        cell_ents[k] = inv.reshape(nc, len(combos)).astype(IntType)
        keys[k] = ukey
    # point numbering: cells, vertices, then dim-1 ... 1
    start = {dim: 0, 0: nc}
    nxt = nc + nv
    for k in range(dim - 1, 0, -1):
        start[k] = nxt
        nxt += ent_verts[k].shape[0]
    pEnd = nxt
    counts = np.zeros(pEnd, dtype=IntType)
    cone_chunks = {}
    for k in range(1, dim + 1):
        ev = ent_verts[k]
        n = ev.shape[0]
        counts[start[k]:start[k] + n] = k + 1
        if k == 1:
            cone_chunks[k] = (ev + start[0]).reshape(-1)
        elif k == dim:
            cone_chunks[k] = (cell_ents[dim - 1] + start[dim - 1]).reshape(-1)
        else:
            sub = np.stack([np.delete(ev, j, axis=1) for j in range(k, -1, -1)], axis=1)
            kk = _row_keys(sub.reshape(-1, k), base)
            loc = np.searchsorted(keys[k - 1], kk)
            cone_chunks[k] = (loc + start[k - 1]).astype(IntType)
    cone_off = np.zeros(pEnd + 1, dtype=IntType)
    np.cumsum(counts, out=cone_off[1:])
    cones = np.zeros(int(cone_off[-1]), dtype=IntType)
    for k in range(1, dim + 1):
        s = start[k]
        cones[cone_off[s]:cone_off[s + ent_verts[k].shape[0]]] = cone_chunks[k]
    ds = [start[d] for d in range(dim + 1)]
    de = [start[d] + ent_verts[d].shape[0] for d in range(dim + 1)]
    plex = Plex(dim, cone_off, cones, ds, de, coords=np.asarray(coords, dtype=float),
                cell_type="simplex")
    plex.cell_verts = cv
    for k in range(dim + 1):
        plex.cell_entities[k] = cell_ents[k] + start[k]
        plex.entity_verts[k] = ent_verts[k]
    return plex


def IntervalMesh(ncells, length=1.0):
    x = np.linspace(0.0, length, ncells + 1)[:, None]
    cv = np.stack([np.arange(ncells), np.arange(1, ncells + 1)], axis=1)
    return simplex_plex(1, cv, x)


def RectangleMesh(nx, ny, Lx=1.0, Ly=1.0, diagonal="left"):
    """Triangulated rectangle, each square split by one diagonal (Firedrake default "left")."""
    xs, ys = np.linspace(0, Lx, nx + 1), np.linspace(0, Ly, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    vid = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    a, b, c, d = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel(), vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    if diagonal == "left":   # diagonal from (i, j+1) to (i+1, j)
        cells = np.concatenate([np.stack([a, b, d], 1), np.stack([b, c, d], 1)])
    else:                    # diagonal from (i, j) to (i+1, j+1)
        cells = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    return simplex_plex(2, cells, coords)


def UnitSquareMesh(nx, ny, diagonal="left"):
    return RectangleMesh(nx, ny, 1.0, 1.0, diagonal)


def BoxMesh(nx, ny, nz, Lx=1.0, Ly=1.0, Lz=1.0):
    """Tetrahedral box: every cube split into six tets around the main diagonal (Kuhn/Freudenthal)."""
    xs, ys, zs = (np.linspace(0, L, n + 1) for L, n in ((Lx, nx), (Ly, ny), (Lz, nz)))
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    vid = np.arange((nx + 1) * (ny + 1) * (nz + 1)).reshape(nx + 1, ny + 1, nz + 1)

    def corner(dx, dy, dz):
        return vid[dx:nx + dx, dy:ny + dy, dz:nz + dz].ravel()
    cells = []
    from itertools import permutations
    for perm in permutations(range(3)):
        step = [0, 0, 0]
        verts = [corner(0, 0, 0)]
        for ax in perm:
            step[ax] = 1
            verts.append(corner(*step))
        cells.append(np.stack(verts, axis=1))
    return simplex_plex(3, np.concatenate(cells), coords)


# ----------------------------------------------------------------------
# tensor-product (interval / quadrilateral / hexahedral) structured meshes
# ----------------------------------------------------------------------
class TensorPlex(Plex):
    """Structured d-dimensional box of tensor cells.

    An entity is identified by the set S of axes it extends along and a
    multi-index; its dimension is |S|.  The cone of (S, i) is, for each a in
    S, the two entities (S - {a}, i) and (S - {a}, i + e_a).
    """

    def __init__(self, ncells, lengths=None, ghost_layers=None):
        n = tuple(int(x) for x in ncells)
        d = len(n)
        self.n = n
        self.lengths = tuple(1.0 for _ in n) if lengths is None else tuple(float(x) for x in lengths)
        groups = {k: [frozenset(c) for c in combinations(range(d), k)] for k in range(d + 1)}
        order = [d, 0] + list(range(d - 1, 0, -1))
        self.groups = []          # list of (S, shape, start) in point order
        self.group_of = {}
        nxt = 0
        ds, de = [0] * (d + 1), [0] * (d + 1)
        for k in order:
            ds[k] = nxt
            for S in groups[k]:
                shape = tuple(n[a] if a in S else n[a] + 1 for a in range(d))
                self.group_of[S] = len(self.groups)
                self.groups.append((S, shape, nxt))
                nxt += int(np.prod(shape))
            de[k] = nxt
        pEnd = nxt
        counts = np.zeros(pEnd, dtype=IntType)
        chunks = []
        for S, shape, start in self.groups:
            size = int(np.prod(shape))
            counts[start:start + size] = 2 * len(S)
            if not S:
                continue
            idx = np.indices(shape).reshape(d, -1)
            cols = []
            for a in sorted(S):
                T = S - {a}
                _, tshape, tstart = self.groups[self.group_of[T]]
                lo = np.ravel_multi_index(tuple(idx), tshape) + tstart
                up = idx.copy()
                up[a] += 1
                hi = np.ravel_multi_index(tuple(up), tshape) + tstart
                cols += [lo, hi]
            chunks.append((start, np.stack(cols, axis=1).reshape(-1)))
        cone_off = np.zeros(pEnd + 1, dtype=IntType)
        np.cumsum(counts, out=cone_off[1:])
        cones = np.zeros(int(cone_off[-1]), dtype=IntType)
        for start, vals in chunks:
            cones[cone_off[start]:cone_off[start] + vals.size] = vals
        vshape = tuple(x + 1 for x in n)
        grids = np.meshgrid(*[np.linspace(0, L, m) for L, m in zip(self.lengths, vshape)], indexing="ij")
        coords = np.stack([g.ravel() for g in grids], axis=1)
        Plex.__init__(self, d, cone_off, cones, ds, de, coords=coords, cell_type="tensor")

    def entity_point(self, S, idx):
        """Point ids of the entities (S, idx); idx is a (d, ...) integer array."""
        _, shape, start = self.groups[self.group_of[frozenset(S)]]
        return np.ravel_multi_index(tuple(idx), shape) + start

    def mark_ghost_box(self, lo, hi):
        """Label as ``pyop2_ghost`` every point whose lowest vertex lies outside the index box lo <= i <= hi (per axis):
        the owned points of a brick partition (an entity belongs to the rank that owns its lowest vertex)."""
        for S, shape, start in self.groups:
            idx = np.indices(shape).reshape(self.dim, -1)
            g = np.zeros(idx.shape[1], dtype=bool)
            for a in range(self.dim):
                g |= (idx[a] < lo[a]) | (idx[a] > hi[a])
            self.ghost[start:start + g.size] = g.astype(np.uint8)

    def mark_ghost_vertices(self, axis, lo, hi):
        """Label as ``pyop2_ghost`` every point that has a vertex outside [lo, hi] along ``axis``
        ... i.e. keep as owned exactly the points whose lowest vertex plane lies in [lo, hi]."""
        for S, shape, start in self.groups:
            idx = np.indices(shape).reshape(self.dim, -1)
            owner_plane = idx[axis]
            g = (owner_plane < lo) | (owner_plane > hi)
            self.ghost[start:start + g.size] = g.astype(np.uint8)


def UnitCubeHexMesh(n):
    return TensorPlex((n, n, n))


def UnitSquareQuadMesh(n):
    return TensorPlex((n, n))
