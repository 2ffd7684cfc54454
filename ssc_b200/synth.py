"""Python face of csrc/synth.cpp: assembled Q_p Poisson operators at benchmark sizes (test/bench support)."""
import ctypes as C
import os

import numpy as np

from . import fem

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libssc_synth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            from . import build
            build.build()
        _lib = C.CDLL(_SO)
    return _lib


def set_threads(n):
    _load().ssc_synth_set_threads(C.c_int(int(n)))


def tensor_poisson_csr(V, bc_nodes=None):
    """(indptr int64, indices int32, data float64) of the Q_p stiffness matrix on the tensor space V, in V's node
    numbering, with the rows and columns of ``bc_nodes`` replaced by the identity."""
    L = _load()
    d, p = V.mesh.dim, V.degree
    N = np.ascontiguousarray(V.lattice_shape, dtype=np.int64)
    ops = fem.tensor_1d_operators(V)
    K = [np.ascontiguousarray(o[0]) for o in ops]
    M = [np.ascontiguousarray(o[1]) for o in ops]
    perm = np.ascontiguousarray(V.lattice_node.ravel(), dtype=np.int64)
    n = perm.size
    isbc = None
    if bc_nodes is not None:
        mask = np.zeros(n, dtype=np.uint8)
        mask[bc_nodes] = 1
        isbc = np.ascontiguousarray(mask[perm])
    rowptr = np.zeros(n + 1, dtype=np.int64)
    i64p, f64p = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    L.ssc_synth_tensor_rowptr(C.c_int(d), N.ctypes.data_as(i64p), C.c_int(p), perm.ctypes.data_as(i64p), rowptr.ctypes.data_as(i64p))
    nnz = int(rowptr[-1])
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    PK = f64p * d
    L.ssc_synth_tensor_fill(C.c_int(d), N.ctypes.data_as(i64p), C.c_int(p), PK(*[k.ctypes.data_as(f64p) for k in K]),
                            PK(*[m.ctypes.data_as(f64p) for m in M]), perm.ctypes.data_as(i64p),
                            isbc.ctypes.data_as(C.c_void_p) if isbc is not None else None,
                            rowptr.ctypes.data_as(i64p), col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(f64p))
    return rowptr, col, val


def _embedding_1d(c, p):
    import scipy.sparse as sp
    t = fem.gll_points(p)
    rows, cols, vals = [], [], []
    for cell in range(c):
        for j in range(p + 1):
            i = cell * p + j
            rows += [i, i]
            cols += [cell, cell + 1]
            vals += [1.0 - t[j], t[j]]
    E = sp.coo_matrix((vals, (rows, cols)), shape=(c * p + 1, c + 1)).tocsr()
    cnt = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=E.shape).tocsr()
    E.data /= cnt.data
    E.eliminate_zeros()
    return E


def tensor_p1_restriction(Vk, V1, bck, bc1, c, p):
    """Restriction matrix (P1 dofs x Q_p dofs) of ssc/lo.py:83-102 on a c x c x c tensor mesh, built as a Kronecker product of
    the 1-D interpolation instead of the per-cell loop (benchmark sizes); rows / columns of Dirichlet nodes dropped."""
    import scipy.sparse as sp
    E1 = _embedding_1d(c, p)
    E = sp.kron(E1, sp.kron(E1, E1, format="csr"), format="csr")          # lattice ordering, x slowest
    inv_f = np.empty(Vk.node_count, dtype=np.int64)
    inv_f[Vk.lattice_node.ravel()] = np.arange(Vk.node_count)
    inv_c = np.empty(V1.node_count, dtype=np.int64)
    inv_c[V1.lattice_node.ravel()] = np.arange(V1.node_count)
    E = E[inv_f][:, inv_c]
    k1 = np.ones(V1.node_count)
    k1[bc1] = 0.0
    kk = np.ones(Vk.node_count)
    kk[bck] = 0.0
    R = (sp.diags(k1) @ E.T @ sp.diags(kk)).tocsr()
    R.eliminate_zeros()
    R.sort_indices()
    return R
