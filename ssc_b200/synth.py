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
