"""Synthetic finite-element problems: the inputs the reference takes from Firedrake.

``setup_patch_pc`` (ssc/patch.py:184-226) hands the C layer, per subspace, a
PetscSection (dofs per mesh point), a block size, and ``cell_node_list``; plus
the Dirichlet dof lists from ``bcdofs`` (ssc/patch.py:135-166).  Firedrake,
PyOP2 and TSFC are absent here, so this module produces the same objects for
continuous Lagrange spaces on the meshes of :mod:`ssc_b200.plex`, and
assembles the Poisson / elasticity operators to CSR (the reference
re-assembles patch matrices from element kernels; this build extracts them
from the assembled operator).
"""
from itertools import combinations
from math import comb

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

from .plex import IntType, TensorPlex


# ----------------------------------------------------------------------
# section numbering shared by both cell types
# ----------------------------------------------------------------------
def _number_points(pEnd, first_cell, rank_in_cell, ndof, ghost=None):
    """PetscSection (dof, off) over the chart: points are numbered in the order a
    cell-by-cell closure traversal first meets them (what Firedrake does for locality);
    points labelled ghost come after all owned ones (Firedrake: nodes < dof_dset.size are owned)."""
    if ghost is None or not np.any(ghost):
        order = np.lexsort((rank_in_cell, first_cell))
    else:
        order = np.lexsort((rank_in_cell, first_cell, ghost.astype(np.int64)))
    off = np.zeros(pEnd, dtype=IntType)
    c = np.cumsum(ndof[order])
    off[order] = c - ndof[order]
    return ndof.astype(IntType), off


def _positive_compositions(total, parts):
    """All tuples of ``parts`` positive integers summing to ``total``, lexicographic."""
    if parts == 1:
        return [(total,)] if total >= 1 else []
    out = []
    for first in range(1, total - parts + 2):
        out += [(first,) + rest for rest in _positive_compositions(total - first, parts - 1)]
    return out


class FunctionSpace(object):
    """Continuous Lagrange space of degree p with ``bs`` components per node."""

    def __init__(self, mesh, degree, bs=1, name="CG"):
        self.mesh = mesh
        self.degree = int(degree)
        self.bs = int(bs)
        self.value_size = self.bs
        if mesh.cell_type == "simplex":
            self._init_simplex()
        else:
            self._init_tensor()
        self.node_count = int(self.sec_dof.sum())
        self.dof_count = self.node_count * self.bs
        self.owned_node_count = int(self.sec_dof[mesh.ghost == 0].sum())

    # -- simplices -----------------------------------------------------
    def _init_simplex(self):
        m, p, d = self.mesh, self.degree, self.mesh.dim
        nc = m.num_cells
        per_dim = [comb(p - 1, k) for k in range(d + 1)]   # interior nodes of a k-simplex
        ndof = np.zeros(m.pEnd, dtype=IntType)
        first_cell = np.full(m.pEnd, np.iinfo(np.int64).max, dtype=IntType)
        rank = np.zeros(m.pEnd, dtype=IntType)
        r0 = 0
        cells = np.arange(nc, dtype=IntType)
        for k in range(d + 1):
            ce = m.cell_entities[k]
            s, e = m.depth_stratum(k)
            ndof[s:e] = per_dim[k]
            for j in range(ce.shape[1] - 1, -1, -1):
                np.minimum.at(first_cell, ce[:, j], cells)
            r0 += ce.shape[1]
        # rank of a point inside the closure of its first cell
        base = 0
        for k in range(d + 1):
            ce = m.cell_entities[k]
            for j in range(ce.shape[1]):
                hit = first_cell[ce[:, j]] == cells
                rank[ce[hit, j]] = base + j
            base += ce.shape[1]
        self.sec_dof, self.sec_off = _number_points(m.pEnd, first_cell, rank, ndof, m.ghost)
        # local lattice: basis functions ordered by entity dimension, entity, local dof
        alphas, ent = [], []
        for k in range(d + 1):
            combos = list(combinations(range(d + 1), k + 1))
            comps = _positive_compositions(p, k + 1)
            for ci, c in enumerate(combos):
                for li, comp_ in enumerate(comps):
                    a = [0] * (d + 1)
                    for pos, val in zip(c, comp_):
                        a[pos] = val
                    alphas.append(a)
                    ent.append((k, ci, li))
        self.alphas = np.array(alphas, dtype=IntType)
        self.nodes_per_cell = len(alphas)
        cnl = np.empty((nc, self.nodes_per_cell), dtype=IntType)
        for b, (k, ci, li) in enumerate(ent):
            cnl[:, b] = self.sec_off[m.cell_entities[k][:, ci]] + li
        self.cell_node_list = np.ascontiguousarray(cnl)

    def tabulate_simplex(self, lam):
        """Values and barycentric derivatives of all local basis functions at points ``lam`` (nq, d+1)."""
        p = self.degree
        nq, nb = lam.shape[0], self.nodes_per_cell
        amax = p
        f = np.ones((amax + 1, nq, lam.shape[1]))
        df = np.zeros_like(f)
        for a in range(1, amax + 1):
            g = (p * lam - (a - 1)) / a
            f[a] = f[a - 1] * g
            df[a] = df[a - 1] * g + f[a - 1] * (p / a)
        val = np.ones((nb, nq))
        dval = np.zeros((nb, nq, lam.shape[1]))
        for b in range(nb):
            fac = [f[self.alphas[b, i], :, i] for i in range(lam.shape[1])]
            dfac = [df[self.alphas[b, i], :, i] for i in range(lam.shape[1])]
            val[b] = np.prod(fac, axis=0)
            for i in range(lam.shape[1]):
                term = dfac[i].copy()
                for j in range(lam.shape[1]):
                    if j != i:
                        term = term * fac[j]
                dval[b, :, i] = term
        return val, dval

    # -- tensor cells --------------------------------------------------
    def _init_tensor(self):
        m, p, d = self.mesh, self.degree, self.mesh.dim
        assert isinstance(m, TensorPlex)
        n = m.n
        lshape = tuple(p * x + 1 for x in n)
        L = np.indices(lshape).reshape(d, -1)
        rem = L % p
        quo = L // p
        on = rem != 0                                   # axes the owning entity extends along
        code = np.zeros(L.shape[1], dtype=IntType)
        for a in range(d):
            code |= on[a].astype(IntType) << a
        point = np.empty(L.shape[1], dtype=IntType)
        ldof = np.zeros(L.shape[1], dtype=IntType)
        for S, shape, start in m.groups:
            c = sum(1 << a for a in S)
            sel = code == c
            if not sel.any():
                continue
            point[sel] = np.ravel_multi_index(tuple(quo[:, sel]), shape) + start
            li = np.zeros(int(sel.sum()), dtype=IntType)
            for a in sorted(S):
                li = li * (p - 1) + (rem[a, sel] - 1)
            ldof[sel] = li
        ndof = np.zeros(m.pEnd, dtype=IntType)
        for S, shape, start in m.groups:
            ndof[start:start + int(np.prod(shape))] = (p - 1) ** len(S)
        # first cell touching each point, and the point's rank inside that cell's closure
        first_cell = np.zeros(m.pEnd, dtype=IntType)
        rank = np.zeros(m.pEnd, dtype=IntType)
        gi = 0
        for S, shape, start in m.groups:
            idx = np.indices(shape).reshape(d, -1)
            cmin = np.stack([idx[a] if a in S else np.maximum(idx[a] - 1, 0) for a in range(d)])
            first_cell[start:start + idx.shape[1]] = np.ravel_multi_index(tuple(cmin), n)
            side = np.zeros(idx.shape[1], dtype=IntType)
            for a in range(d):
                if a not in S:
                    side = side * 2 + (idx[a] - cmin[a])
            rank[start:start + idx.shape[1]] = gi * (1 << d) + side
            gi += 1
        self.sec_dof, self.sec_off = _number_points(m.pEnd, first_cell, rank, ndof, m.ghost)
        self.lattice_shape = lshape
        self.lattice_node = (self.sec_off[point] + ldof).reshape(lshape)
        self.nodes_per_cell = (p + 1) ** d
        cidx = np.indices(n).reshape(d, -1)
        loc = np.indices((p + 1,) * d).reshape(d, -1)
        glob = cidx[:, :, None] * p + loc[:, None, :]
        self.cell_node_list = np.ascontiguousarray(self.lattice_node[tuple(glob)], dtype=IntType)

    # -- boundary conditions ------------------------------------------
    def boundary_nodes(self):
        """Nodes of DirichletBC(V, 0, "on_boundary"): nodes on points in the closure of exterior facets."""
        mask = self.mesh.boundary_points()
        pts = np.nonzero(mask & (self.sec_dof > 0))[0]
        if pts.size == 0:
            return np.zeros(0, dtype=IntType)
        from .plex import _ranges
        lo = self.sec_off[pts]
        return np.sort(_ranges(lo, lo + self.sec_dof[pts]))


def bcdofs(V, nodes, offset=0, owned_nodes=None):
    """Dirichlet dofs in the concatenated numbering (ssc/patch.py:135-166): all components of each node."""
    nodes = np.asarray(nodes, dtype=IntType)
    if owned_nodes is not None:
        nodes = nodes[nodes < owned_nodes]
    bs = V.bs
    return np.concatenate([nodes * bs + j for j in range(bs)]) + offset if nodes.size else np.zeros(0, IntType)


# ----------------------------------------------------------------------
# quadrature
# ----------------------------------------------------------------------
def simplex_quadrature(d, n):
    """Collapsed Gauss-Jacobi rule on the unit d-simplex, exact to degree 2n-1.  Returns barycentric points."""
    pts1 = []
    for k in range(d):
        x, w = roots_jacobi(n, d - 1 - k, 0)
        pts1.append(((x + 1) / 2, w / 2 ** (d - k)))
    grids = np.meshgrid(*[p[0] for p in pts1], indexing="ij")
    wg = np.meshgrid(*[p[1] for p in pts1], indexing="ij")
    u = [g.ravel() for g in grids]
    w = np.prod([g.ravel() for g in wg], axis=0)
    xi = []
    scale = np.ones_like(u[0])
    for k in range(d):
        xi.append(u[k] * scale)
        scale = scale * (1 - u[k])
    xi = np.stack(xi, axis=1)
    lam = np.concatenate([1 - xi.sum(axis=1, keepdims=True), xi], axis=1)
    return lam, w


def gll_points(p):
    """Gauss-Lobatto-Legendre points on [0, 1]."""
    if p == 1:
        return np.array([0.0, 1.0])
    x, _ = roots_jacobi(p - 1, 1, 1)
    return np.concatenate(([0.0], (x + 1) / 2, [1.0]))


def lagrange_1d(nodes, x):
    """Values and derivatives at ``x`` of the Lagrange basis on ``nodes``."""
    n = len(nodes)
    V = np.ones((n, len(x)))
    D = np.zeros((n, len(x)))
    for i in range(n):
        for j in range(n):
            if j == i:
                continue
            V[i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for k in range(n):
            if k == i:
                continue
            t = np.ones(len(x)) / (nodes[i] - nodes[k])
            for j in range(n):
                if j in (i, k):
                    continue
                t *= (x - nodes[j]) / (nodes[i] - nodes[j])
            D[i] += t
    return V, D


def reference_1d(p):
    """1-D reference stiffness and mass on [0, 1] for degree-p GLL Lagrange."""
    nodes = gll_points(p)
    xg, wg = np.polynomial.legendre.leggauss(p + 2)
    xg, wg = (xg + 1) / 2, wg / 2
    V, D = lagrange_1d(nodes, xg)
    return (D * wg) @ D.T, (V * wg) @ V.T


# ----------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------
def _simplex_geometry(mesh):
    cv = mesh.cell_verts
    X = mesh.coords[cv]                                    # (nc, d+1, gdim)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))  # columns v_j - v_0
    detJ = np.abs(np.linalg.det(J))
    Jinv = np.linalg.inv(J)
    return J, Jinv, detJ


def _coo_to_csr(n, rows, cols, vals):
    A = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def assemble_poisson_simplex(V):
    """Scalar stiffness matrix of inner(grad u, grad v)*dx on a simplicial space (node numbering)."""
    m, d = V.mesh, V.mesh.dim
    lam, w = simplex_quadrature(d, V.degree + 1)
    _, dval = V.tabulate_simplex(lam)
    dxi = dval[:, :, 1:] - dval[:, :, :1]                  # d/dxi_j = d/dlam_j - d/dlam_0
    S = np.einsum("q,iqa,jqb->abij", w, dxi, dxi)
    _, Jinv, detJ = _simplex_geometry(m)
    G = np.einsum("c,cak,cbk->cab", detJ, Jinv, Jinv)
    Ke = np.einsum("cab,abij->cij", G, S)
    cnl = V.cell_node_list
    rows = np.repeat(cnl[:, :, None], cnl.shape[1], axis=2)
    cols = np.repeat(cnl[:, None, :], cnl.shape[1], axis=1)
    return _coo_to_csr(V.node_count, rows, cols, Ke), Ke


def elasticity_element_matrices(V, lmbda=1.0, mu=1.0):
    """Element matrices of 2 mu eps(u):eps(v) + lambda div u div v on a vector simplicial space, one (nb d) x (nb d)
    matrix per cell, cell dofs node by node, component by component (what the reference's element kernel computes per
    cell, ssc/patch.py:28-62)."""
    m, d = V.mesh, V.mesh.dim
    assert V.bs == d
    lam, w = simplex_quadrature(d, V.degree + 1)
    _, dval = V.tabulate_simplex(lam)
    dxi = dval[:, :, 1:] - dval[:, :, :1]
    S = np.einsum("q,iqa,jqb->abij", w, dxi, dxi)
    _, Jinv, detJ = _simplex_geometry(m)
    # H[c, r, s, i, j] = int d_r phi_i d_s phi_j
    H = np.einsum("c,car,cbs,abij->crsij", detJ, Jinv, Jinv, S, optimize=True)
    nb = V.nodes_per_cell
    Ke = np.zeros((m.num_cells, nb, d, nb, d))
    tr = np.einsum("crrij->cij", H)
    for a in range(d):
        for b in range(d):
            Ke[:, :, a, :, b] = lmbda * H[:, a, b] + mu * H[:, b, a]
            if a == b:
                Ke[:, :, a, :, b] += mu * tr
    return Ke.reshape(m.num_cells, nb * d, nb * d)


def assemble_elasticity_simplex(V, lmbda=1.0, mu=1.0):
    """Linear elasticity 2 mu eps(u):eps(v) + lambda div u div v on a vector simplicial space (dof numbering).
    Returns the assembled operator and the element matrices it was summed from."""
    m, d = V.mesh, V.mesh.dim
    nb = V.nodes_per_cell
    Ke = elasticity_element_matrices(V, lmbda, mu)
    dofs = (V.cell_node_list[:, :, None] * d + np.arange(d)[None, None, :]).reshape(m.num_cells, nb * d)
    rows = np.repeat(dofs[:, :, None], nb * d, axis=2)
    cols = np.repeat(dofs[:, None, :], nb * d, axis=1)
    return _coo_to_csr(V.dof_count, rows, cols, Ke), Ke


def tensor_1d_operators(V):
    """Per axis: assembled 1-D stiffness and mass (dense, lattice numbering) for a tensor space."""
    p = V.degree
    Kh, Mh = reference_1d(p)
    out = []
    for a, n in enumerate(V.mesh.n):
        h = V.mesh.lengths[a] / n
        N = p * n + 1
        K = np.zeros((N, N))
        M = np.zeros((N, N))
        for e in range(n):
            s = slice(p * e, p * e + p + 1)
            K[s, s] += Kh / h
            M[s, s] += Mh * h
        out.append((K, M))
    return out


def assemble_poisson_tensor_scipy(V):
    """Kronecker-sum assembly through scipy (small meshes; the threaded generator in
    csrc/synth.cpp is the one used at benchmark sizes and is checked against this)."""
    ops = tensor_1d_operators(V)
    d = V.mesh.dim
    A = None
    for a in range(d):
        term = None
        for b in range(d):
            f = sp.csr_matrix(ops[b][0] if b == a else ops[b][1])
            term = f if term is None else sp.kron(term, f, format="csr")
        A = term if A is None else A + term
    perm = V.lattice_node.ravel()
    P = sp.csr_matrix((np.ones(perm.size), (perm, np.arange(perm.size))), shape=(perm.size,) * 2)
    A = (P @ A @ P.T).tocsr()
    A.sort_indices()
    return A


def element_matrix_tensor(V):
    """Element stiffness of one tensor cell (all cells are congruent), local lexicographic order."""
    p, d = V.degree, V.mesh.dim
    Kh, Mh = reference_1d(p)
    Ke = None
    for a in range(d):
        term = None
        for b in range(d):
            h = V.mesh.lengths[b] / V.mesh.n[b]
            f = Kh / h if b == a else Mh * h
            term = f if term is None else np.kron(term, f)
        Ke = term if Ke is None else Ke + term
    return Ke


def expand_block(A, bs):
    """Componentwise operator for a vector/tensor space: kron(A, I_bs) in interleaved dof numbering."""
    if bs == 1:
        return A
    B = sp.kron(A, sp.identity(bs), format="csr")
    B.sort_indices()
    return B


def apply_bcs(A, bc_dofs):
    """Zero the rows and columns of Dirichlet dofs and put one on their diagonal (Firedrake's convention)."""
    n = A.shape[0]
    keep = np.ones(n)
    keep[bc_dofs] = 0.0
    D = sp.diags(keep)
    B = (D @ A @ D + sp.diags(1.0 - keep)).tocsr()
    B.sort_indices()
    return B


# ----------------------------------------------------------------------
# the bundle PatchPC needs
# ----------------------------------------------------------------------
class Discretisation(object):
    """What ``setup_patch_pc`` gathers from the form and its function space (ssc/patch.py:184-226)."""

    def __init__(self, spaces, bc_nodes=None, owned_nodes=None):
        self.spaces = list(spaces)
        self.mesh = self.spaces[0].mesh
        self.bs = np.array([W.bs for W in self.spaces], dtype=IntType)
        self.offsets = np.append([0], np.cumsum([W.dof_count for W in self.spaces])).astype(IntType)
        if bc_nodes is None:
            bc_nodes = [W.boundary_nodes() for W in self.spaces]
        self.bc_nodes = bc_nodes
        ghost, glob = [], []
        goff = 0
        for i, (W, nodes) in enumerate(zip(self.spaces, bc_nodes)):
            ghost.append(bcdofs(W, nodes, self.offsets[i]))
            own = None if owned_nodes is None else owned_nodes[i]
            glob.append(bcdofs(W, nodes, goff, own))
            goff += (W.node_count if own is None else own) * W.bs
        self.ghost_bc_nodes = np.unique(np.concatenate(ghost)).astype(IntType)
        self.global_bc_nodes = np.unique(np.concatenate(glob)).astype(IntType)
        nc = self.mesh.num_cells
        # cell numbering section: plex cell point -> cell_node_list row (identity here)
        self.cell_numbering_dof = np.zeros(self.mesh.pEnd, dtype=IntType)
        self.cell_numbering_off = np.zeros(self.mesh.pEnd, dtype=IntType)
        self.cell_numbering_dof[self.mesh.cStart:self.mesh.cEnd] = 1
        self.cell_numbering_off[self.mesh.cStart:self.mesh.cEnd] = np.arange(nc)
        self.total_dofs = int(self.offsets[-1])

    def block_diag(self, mats):
        return sp.block_diag(mats, format="csr")


# ----------------------------------------------------------------------
# stand-ins for the Firedrake objects ssc/patch.py:231-253 digs the form and BCs out of
# ----------------------------------------------------------------------
class DirichletBC(object):
    """Homogeneous Dirichlet condition on ``nodes`` of subspace ``index`` (all components)."""

    def __init__(self, index, nodes):
        self.index = int(index)
        self.nodes = np.asarray(nodes, dtype=IntType)

    def __eq__(self, other):
        return isinstance(other, DirichletBC) and self.index == other.index and np.array_equal(self.nodes, other.nodes)

    def __ne__(self, other):
        return not self == other


class Form(object):
    """A bilinear form on a (mixed) space: knows its function spaces and how to assemble itself to CSR."""

    def __init__(self, spaces, assemble, owned_nodes=None):
        self.spaces = list(spaces)
        self._assemble = assemble
        self.owned_nodes = owned_nodes

    def ufl_domain(self):
        return self.spaces[0].mesh

    def assemble(self, bcs=()):
        """Local operator (rows for all local dofs) with Dirichlet rows/columns replaced by the identity."""
        return self._assemble(bcs)


class ImplicitMatrixContext(object):
    """What a matfree Firedrake operator carries (ssc/patch.py:233-243)."""

    def __init__(self, a, row_bcs, col_bcs=None):
        self.a = a
        self.row_bcs = list(row_bcs)
        self.col_bcs = list(row_bcs if col_bcs is None else col_bcs)
        self._csr = None

    def mult(self, mat, x, y):
        if self._csr is None:
            self._csr = self.a.assemble(self.row_bcs)
        y.array[:] = self._csr @ x.array


class _Problem(object):
    def __init__(self, bcs):
        self.bcs = list(bcs)


class SNESContext(object):
    """What sits in the DM's appctx for an assembled operator (ssc/patch.py:245-253)."""

    def __init__(self, J, bcs, Jp=None):
        self.J = J
        self.Jp = Jp
        self._problem = _Problem(bcs)


def discretisation_from_form(J, bcs):
    nodes = [[] for _ in J.spaces]
    for bc in bcs:
        nodes[bc.index].append(bc.nodes)
    merged = [np.unique(np.concatenate(n)) if n else np.zeros(0, dtype=IntType) for n in nodes]
    return Discretisation(J.spaces, bc_nodes=merged, owned_nodes=J.owned_nodes)


def poisson_problem(mesh, degree, kind="scalar"):
    """inner(grad u, grad v)*dx with zero Dirichlet data on the whole boundary, as in tests/test_jacobi_equivalent.py:25-50.
    kind: scalar | vector | tensor | mixed (P x Q x R).  Returns (Form, bcs)."""
    d = mesh.dim
    sizes = {"scalar": [1], "vector": [d], "tensor": [d * d], "mixed": [1, d, d * d]}[kind]
    spaces = [FunctionSpace(mesh, degree, bs) for bs in sizes]
    V0 = spaces[0]
    for W in spaces[1:]:                      # same scalar layout, different block size
        assert np.array_equal(W.cell_node_list, V0.cell_node_list)
    if mesh.cell_type == "simplex":
        A0, _ = assemble_poisson_simplex(V0)
    else:
        A0 = assemble_poisson_tensor_scipy(V0)
    bcs = [DirichletBC(i, W.boundary_nodes()) for i, W in enumerate(spaces)]

    def assemble(bcs_):
        A = sp.block_diag([expand_block(A0, W.bs) for W in spaces], format="csr")
        disc = discretisation_from_form(form, bcs_)
        return apply_bcs(A, disc.ghost_bc_nodes)
    form = Form(spaces, assemble)
    form.low_order = lambda: poisson_problem(mesh, 1, kind)          # the same form on the degree-1 space (ssc/lo.py:122-131)
    form.load_vector = lambda: load_vector(spaces)
    return form, bcs


def poisson_element_matrices(spaces):
    """Element matrices of poisson_problem's form: what the reference's compute-operator callback feeds to
    MatSetValues cell by cell (ssc/patch.py:28-62).  Cell dofs run subspace by subspace, node by node, component by
    component (the order of the per-cell dof table, ssc/libssc.c:846-889).  Returns (ncells, t, t) on simplices and
    one shared (t, t) matrix on tensor meshes (all cells congruent)."""
    V0 = spaces[0]
    if V0.mesh.cell_type == "simplex":
        K0 = assemble_poisson_simplex(V0)[1]
    else:
        K0 = element_matrix_tensor(V0)[None]
    nb = K0.shape[-1]
    t = sum(nb * W.bs for W in spaces)
    Ke = np.zeros((K0.shape[0], t, t))
    o = 0
    for W in spaces:
        blk = np.einsum("cij,ab->ciajb", K0, np.eye(W.bs)).reshape(K0.shape[0], nb * W.bs, nb * W.bs)
        Ke[:, o:o + nb * W.bs, o:o + nb * W.bs] = blk
        o += nb * W.bs
    return Ke if V0.mesh.cell_type == "simplex" else Ke[0]


def load_vector(spaces):
    """inner(Constant(1), v)*dx for every component (the right-hand side of the reference's tests)."""
    out = []
    for V in spaces:
        m = V.mesh
        if m.cell_type == "simplex":
            lam, w = simplex_quadrature(m.dim, V.degree + 1)
            val, _ = V.tabulate_simplex(lam)
            _, _, detJ = _simplex_geometry(m)
            loc = val @ w
            b = np.zeros(V.node_count)
            np.add.at(b, V.cell_node_list.ravel(), (detJ[:, None] * loc[None, :]).ravel())
        else:
            _, Mh = reference_1d(V.degree)
            loc = None
            for a in range(m.dim):
                f = Mh.sum(axis=1) * (m.lengths[a] / m.n[a])
                loc = f if loc is None else np.kron(loc, f)
            b = np.zeros(V.node_count)
            np.add.at(b, V.cell_node_list.ravel(), np.broadcast_to(loc, V.cell_node_list.shape).ravel())
        out.append(np.repeat(b, V.bs))
    return np.concatenate(out)
