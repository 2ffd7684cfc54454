"""Low-order (P1) coarse correction, SURVEY.md 8(f) N1: mirror of ssc/lo.py.

``P1PC`` re-poses the operator on the degree-1 space of the same mesh, builds the restriction matrix
R (P1 <- Pk, the transpose of the natural embedding, with Dirichlet rows/columns dropped: ssc/lo.py:83-102) and
applies ``y = R^T A1^{-1} R x`` with boundary values carried over (ssc/lo.py:187-198).  The host assembles R and
A1 (Firedrake's job in the reference); the products and the coarse solve run on the GPU through
``SSCLowOrder*`` in include/ssc_b200.h.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import fem, petsc_lite
from ._lib import CHKERR, MEM_DEVICE, MEM_HOST, as_i64, dp, ip, lib
from .patch import PCBase, _form_and_bcs


def embedding_matrix(Vk, V1):
    """(Vk.node_count x V1.node_count) matrix of the natural embedding P1 -> Pk: P1 basis functions evaluated at the
    Pk nodes, cell by cell with WRITE semantics like the reference's par_loop (ssc/lo.py:89-99)."""
    mesh = Vk.mesh
    if mesh.cell_type == "simplex":
        lam = Vk.alphas / float(Vk.degree)             # barycentric coordinates of the Pk nodes
        local = lam @ V1.alphas.T                      # P1 basis b = barycentric coordinate of its vertex
    else:
        d, p = mesh.dim, Vk.degree
        t = fem.gll_points(p)
        loc = np.indices((p + 1,) * d).reshape(d, -1)
        b = np.indices((2,) * d).reshape(d, -1)
        local = np.ones((loc.shape[1], b.shape[1]))
        for a in range(d):
            ta = t[loc[a]][:, None]
            local = local * np.where(b[a][None, :] == 1, ta, 1.0 - ta)
    rows = np.repeat(Vk.cell_node_list[:, :, None], V1.nodes_per_cell, axis=2).ravel()
    cols = np.repeat(V1.cell_node_list[:, None, :], Vk.nodes_per_cell, axis=1).ravel()
    vals = np.broadcast_to(local, (Vk.cell_node_list.shape[0],) + local.shape).ravel()
    E = sp.coo_matrix((vals, (rows, cols)), shape=(Vk.node_count, V1.node_count)).tocsr()
    cnt = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=E.shape).tocsr()
    E.data /= cnt.data                                  # every cell writes the same value
    E.eliminate_zeros()
    return E


def restriction_matrix(Pk_spaces, P1_spaces, Pk_bc_dofs, P1_bc_dofs):
    """ssc/lo.py:83-102 for a (mixed) space: block diagonal of kron(E^T, I_bs), Dirichlet rows and columns removed."""
    blocks = []
    for Vk, V1 in zip(Pk_spaces, P1_spaces):
        E = embedding_matrix(Vk, V1)
        blocks.append(sp.kron(E.T, sp.identity(Vk.bs), format="csr"))
    R = sp.block_diag(blocks, format="csr")
    k1 = np.ones(R.shape[0])
    k1[P1_bc_dofs] = 0.0
    kk = np.ones(R.shape[1])
    kk[Pk_bc_dofs] = 0.0
    R = (sp.diags(k1) @ R @ sp.diags(kk)).tocsr()
    R.eliminate_zeros()
    R.sort_indices()
    return R


def injection_dofs(Pk_spaces, P1_spaces):
    """For every dof of the (mixed) P1 space the dof of the Pk space that sits on the same vertex, same component: the rows of the
    embedding matrix that are unit vectors.  low = tmp[injection_dofs] is `low.interpolate(tmp)` of ssc/lo.py:155."""
    out, off = [], 0
    for Vk, V1 in zip(Pk_spaces, P1_spaces):
        E = embedding_matrix(Vk, V1).tocsr()
        single = np.flatnonzero((np.diff(E.indptr) == 1))
        single = single[np.abs(E.data[E.indptr[single]] - 1.0) < 1e-12]
        node_of = np.full(V1.node_count, -1, dtype=np.int64)
        node_of[E.indices[E.indptr[single]]] = single
        if (node_of < 0).any():
            raise ValueError("a P1 node has no coincident node in the high-order space")
        out.append(off + (node_of[:, None] * Vk.bs + np.arange(Vk.bs)[None, :]).ravel())
        off += Vk.dof_count
    return np.concatenate(out)


class LowOrderCore(object):
    """ctypes face of SSCLowOrder* (the coarse PC object on the device)."""

    def __init__(self, device=0):
        self.handle = C.c_void_p()
        CHKERR(lib.SSCLowOrderCreate(int(device), C.byref(self.handle)))
        self.nfine = 0

    def __del__(self):
        try:
            if self.handle:
                lib.SSCLowOrderDestroy(C.byref(self.handle))
        except Exception:
            pass

    def setOption(self, name, value):
        CHKERR(lib.SSCLowOrderSetOption(self.handle, name.encode(), str(value).encode()))

    def setRestriction(self, R):
        R = R.tocsr()
        rp, ci, v = as_i64(R.indptr), np.ascontiguousarray(R.indices, dtype=np.int32), np.ascontiguousarray(R.data, dtype=np.float64)
        CHKERR(lib.SSCLowOrderSetRestriction(self.handle, C.c_int64(R.shape[0]), C.c_int64(R.shape[1]), ip(rp),
                                             ci.ctypes.data_as(C.POINTER(C.c_int32)), dp(v)))
        self.nfine = R.shape[1]

    def setOperator(self, A):
        A = A.tocsr()
        rp, ci, v = as_i64(A.indptr), np.ascontiguousarray(A.indices, dtype=np.int32), np.ascontiguousarray(A.data, dtype=np.float64)
        CHKERR(lib.SSCLowOrderSetOperator(self.handle, C.c_int64(A.shape[0]), ip(rp), ci.ctypes.data_as(C.POINTER(C.c_int32)), dp(v)))

    def setBcs(self, bc):
        bc = as_i64(bc)
        CHKERR(lib.SSCLowOrderSetBcs(self.handle, C.c_int64(bc.size), ip(bc)))

    def setNearNullSpace(self, vecs):
        """Near-nullspace of the coarse operator for -lo_pc_type gamg: a sequence of vectors over the coarse dofs (or a k x n array)."""
        B = np.ascontiguousarray(np.atleast_2d(np.asarray(vecs, dtype=np.float64)))
        CHKERR(lib.SSCLowOrderSetNearNullSpace(self.handle, C.c_int64(B.shape[1] if B.size else 0), C.c_int(B.shape[0] if B.size else 0), dp(B)))

    def setUp(self):
        CHKERR(lib.SSCLowOrderSetUp(self.handle))

    def applyDevice(self, dx, dy):
        """Device-resident apply (DeviceVec in, DeviceVec out), asynchronous on the coarse PC's stream."""
        CHKERR(lib.SSCLowOrderApply(self.handle, dx.ptr, dy.ptr, MEM_DEVICE))

    def axpyDevice(self, a, dx, dy):
        CHKERR(lib.SSCLowOrderAXPY(self.handle, C.c_int64(dx.n), C.c_double(a), dx.ptr, dy.ptr))

    def synchronize(self):
        CHKERR(lib.SSCLowOrderSynchronize(self.handle))

    def iterations(self):
        """(last, total, solves) of the coarse CG (zeros with the dense coarse factorisation)."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        CHKERR(lib.SSCLowOrderGetIterations(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def apply(self, x, y):
        xa = x if isinstance(x, np.ndarray) else x.array
        ya = y if isinstance(y, np.ndarray) else y.array
        if xa.size != self.nfine or ya.size != self.nfine:
            raise ValueError("vector length does not match the restriction")
        CHKERR(lib.SSCLowOrderApply(self.handle, xa.ctypes.data_as(C.c_void_p), ya.ctypes.data_as(C.c_void_p), MEM_HOST))


class P1PC(PCBase):
    """ssc/lo.py:105-198."""

    device = 0

    def __init__(self, J=None, bcs=None):
        PCBase.__init__(self)
        self.J = J
        self.bcs = bcs

    def initialize(self, pc):
        if self.J is None:
            _, _, self.J, self.bcs, _ = _form_and_bcs(pc)
        J, bcs = self.J, self.bcs
        if not hasattr(J, "low_order"):
            raise NotImplementedError("the form does not know its low-order counterpart")
        lo_J, lo_bcs = J.low_order()                                 # ssc/lo.py:122-140: same form on P1, same boundary
        disc = fem.discretisation_from_form(J, bcs)
        lo_disc = fem.discretisation_from_form(lo_J, lo_bcs)
        self.lo_J, self.lo_bcs = lo_J, lo_bcs
        prefix = pc.getOptionsPrefix() or ""
        opts = petsc_lite.Options(prefix)
        self.lo = LowOrderCore(self.device)
        if len(lo_disc.spaces) == 1 and int(lo_disc.bs[0]) > 1:
            self.lo.setOption("lo_amg_block_size", int(lo_disc.bs[0]))     # vector-valued P1 space: multigrid aggregates nodes, not dofs
        for name in ("lo_mat_type", "lo_ksp_type", "lo_pc_type", "lo_ksp_rtol", "lo_ksp_max_it", "lo_ksp_cg_persistent", "lo_amg_block_size"):    # ssc/lo.py:141, :164-166
            if opts.hasName(name):
                self.lo.setOption(name, opts.getString(name))
        # near-nullspace of P, interpolated to P1 and attached to the coarse operator (ssc/lo.py:146-159): interpolation of a nodal
        # function to the vertices is the injection of its vertex values
        _, P = pc.getOperators()
        nearnullsp = P.getNearNullSpace() if hasattr(P, "getNearNullSpace") else None
        if nearnullsp is not None and nearnullsp.getVecs():
            inj = injection_dofs(J.spaces, lo_J.spaces)
            self.lo.setNearNullSpace([np.asarray(v.array)[inj] for v in nearnullsp.getVecs()])
        self.lo.setRestriction(restriction_matrix(J.spaces, lo_J.spaces, disc.ghost_bc_nodes, lo_disc.ghost_bc_nodes))
        self.lo.setBcs(disc.global_bc_nodes)                         # ssc/lo.py:175-183
        self.lo.setOperator(lo_J.assemble(lo_bcs))

    def update(self, pc):
        self.lo.setOperator(self.lo_J.assemble(self.lo_bcs))         # ssc/lo.py:185-186
        self.lo.setUp()

    def apply(self, pc, x, y):
        self.lo.apply(x, y)

    def applyTranspose(self, pc, x, y):
        raise NotImplementedError("Haven't coded lo-applyTranspose")  # ssc/lo.py:200-201

    def view(self, pc, viewer=None):
        if viewer is not None:
            viewer.printfASCII("Low-order PC\n")
