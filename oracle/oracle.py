"""ctypes face of the CPU oracle (oracle/ssc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_so = os.path.join(_here, "_build", "libssc_oracle.so")
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(_here, "ssc_oracle.c")
    if force or not os.path.exists(_so) or os.path.getmtime(_so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _here, "-s"])
    return _so


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_so):
            build()
        L = C.CDLL(_so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_error.restype = C.c_char_p
        for name in ("npatch", "vstart", "num_cells", "num_gtol", "num_dofs"):
            getattr(L, "oracle_" + name).restype = C.c_int64
            getattr(L, "oracle_" + name).argtypes = [C.c_void_p]
        for name in ("cell_dof", "cell_off", "cells", "gtol_dof", "gtol_off", "gtol", "dofs"):
            getattr(L, "oracle_" + name).restype = i64p
            getattr(L, "oracle_" + name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(i64p)


def _dp(a):
    return a.ctypes.data_as(f64p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class OracleError(RuntimeError):
    pass


class OraclePatchPC(object):
    """The reference's PC "patch" restated on the CPU; method names follow ssc/PatchPC.pyx:101-170."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.oracle_create())
        self._keep = []
        self.opts = dict(construct_type=0, dim=-1, codim=-1, exclude_subspace=-1, vankadim=-1,
                         partition_of_unity=0, symmetrise_sweep=0)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_destroy(self.h)
            self.h = None

    def _chk(self, ierr):
        if ierr:
            raise OracleError("%d: %s" % (ierr, self.L.oracle_error(self.h).decode()))

    def setPlex(self, plex):
        arrs = [_i64(a) for a in (plex.cone_off, plex.cones, plex.supp_off, plex.supports,
                                  plex.depth_start, plex.depth_end)]
        ghost = np.ascontiguousarray(plex.ghost, dtype=np.uint8)
        self._keep += arrs + [ghost]
        self._chk(self.L.oracle_set_plex(self.h, C.c_int64(plex.dim), C.c_int64(plex.pEnd), *[_ip(a) for a in arrs],
                                         ghost.ctypes.data_as(C.c_void_p)))

    def setPatchCellNumbering(self, dof, off):
        dof, off = _i64(dof), _i64(off)
        self._keep += [dof, off]
        self._chk(self.L.oracle_set_cell_numbering(self.h, _ip(dof), _ip(off)))

    def setPatchDiscretisationInfo(self, sections, bs, cellNodeMaps, subspaceOffsets, ghostBcNodes, globalBcNodes):
        n = len(sections)
        sd = [_i64(s[0]) for s in sections]
        so = [_i64(s[1]) for s in sections]
        cnm = [_i64(m) for m in cellNodeMaps]
        bs = _i64(bs)
        npc = _i64([m.shape[1] for m in cnm])
        off = _i64(subspaceOffsets)
        gb, glb = _i64(ghostBcNodes), _i64(globalBcNodes)
        self._keep += sd + so + cnm + [bs, npc, off, gb, glb]
        P = i64p * n
        self._chk(self.L.oracle_set_discretisation_info(
            self.h, C.c_int64(n), P(*[_ip(a) for a in sd]), P(*[_ip(a) for a in so]), _ip(bs), _ip(npc),
            P(*[_ip(a) for a in cnm]), _ip(off), C.c_int64(gb.size), _ip(gb), C.c_int64(glb.size), _ip(glb)))
        self.nlocal = int(off[-1])

    def setFromDiscretisation(self, disc):
        self.setPlex(disc.mesh)
        self.setPatchCellNumbering(disc.cell_numbering_dof, disc.cell_numbering_off)
        self.setPatchDiscretisationInfo([(W.sec_dof, W.sec_off) for W in disc.spaces], disc.bs,
                                        [W.cell_node_list for W in disc.spaces], disc.offsets,
                                        disc.ghost_bc_nodes, disc.global_bc_nodes)

    def setOptions(self, **kw):
        self.opts.update(kw)
        o = self.opts
        self._chk(self.L.oracle_set_options(self.h, int(o["construct_type"]), C.c_int64(o["dim"]), C.c_int64(o["codim"]),
                                            C.c_int64(o["exclude_subspace"]), C.c_int64(o["vankadim"]),
                                            int(o["partition_of_unity"]), int(o["symmetrise_sweep"])))

    def setUserPatches(self, patches, iteration_set):
        off = _i64(np.concatenate(([0], np.cumsum([len(p) for p in patches]))))
        pts = _i64(np.concatenate([np.asarray(p) for p in patches])) if patches else _i64([])
        it = _i64(iteration_set)
        self._keep += [off, pts, it]
        self._chk(self.L.oracle_set_user_patches(self.h, C.c_int64(len(patches)), _ip(off), _ip(pts),
                                                 C.c_int64(it.size), _ip(it)))
        self.setOptions(construct_type=2)

    def setUpPatches(self):
        self._chk(self.L.oracle_setup_patches(self.h))

    def setPatches(self, offsets, gtol, nlocal):
        offsets, gtol = _i64(offsets), _i64(gtol)
        self._chk(self.L.oracle_set_patches(self.h, C.c_int64(offsets.size - 1), _ip(offsets), _ip(gtol)))
        self.nlocal = int(nlocal)

    def _arr(self, name, n):
        p = getattr(self.L, "oracle_" + name)(self.h)
        return np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy() if n else np.zeros(0, dtype=np.int64)

    @property
    def npatch(self):
        return self.L.oracle_npatch(self.h)

    def patches(self):
        """(offsets[npatch+1], gtol[sum n]) -- gtolCounts/gtol of ssc/libssc.c:779-837."""
        n = self.npatch
        dof = self._arr("gtol_dof", n)
        return np.concatenate(([0], np.cumsum(dof))).astype(np.int64), self._arr("gtol", self.L.oracle_num_gtol(self.h))

    def cells(self):
        n = self.npatch
        dof = self._arr("cell_dof", n)
        return np.concatenate(([0], np.cumsum(dof))).astype(np.int64), self._arr("cells", self.L.oracle_num_cells(self.h))

    def dofs(self):
        return self._arr("dofs", self.L.oracle_num_dofs(self.h))

    def setOperatorCSR(self, A):
        rowptr, col, val = _i64(A.indptr), _i64(A.indices), np.ascontiguousarray(A.data, dtype=np.float64)
        self._keep += [rowptr, col, val]
        self._chk(self.L.oracle_set_operator_csr(self.h, C.c_int64(A.shape[0]), _ip(rowptr), _ip(col), _dp(val)))

    def setOperatorCSRArrays(self, rowptr, col32, val):
        rowptr = _i64(rowptr)
        col32 = np.ascontiguousarray(col32, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        self._keep += [rowptr, col32, val]
        self._chk(self.L.oracle_set_operator_csr32(self.h, C.c_int64(rowptr.size - 1), _ip(rowptr),
                                                   col32.ctypes.data_as(C.c_void_p), _dp(val)))

    def setUpOperators(self, keep_mat=False, nthreads=0):
        self._chk(self.L.oracle_setup_operators(self.h, int(keep_mat), int(nthreads)))

    def patchMatrix(self, p):
        off, _ = self.patches()
        n = int(off[p + 1] - off[p])
        out = np.zeros((n, n))
        self._chk(self.L.oracle_get_patch_matrix(self.h, C.c_int64(p), _dp(out)))
        return out

    def patchMatrixFromElements(self, p, Ke):
        off, _ = self.patches()
        n = int(off[p + 1] - off[p])
        out = np.zeros((n, n))
        Ke = np.ascontiguousarray(Ke, dtype=np.float64)
        self._chk(self.L.oracle_patch_matrix_from_elements(self.h, C.c_int64(p), _dp(Ke), _dp(out)))
        return out

    def applyLocal(self, localX, nthreads=1):
        localX = np.ascontiguousarray(localX, dtype=np.float64)
        y = np.zeros_like(localX)
        self._chk(self.L.oracle_apply_local(self.h, C.c_int64(localX.size), _dp(localX), _dp(y), int(nthreads)))
        return y

    def applyScale(self, x):
        """scale_i = sum over the patches containing local dof i of max|B_p^-1 R_p x| (tests/helpers.entry_err)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.zeros_like(x)
        self._chk(self.L.oracle_apply_scale(self.h, C.c_int64(x.size), _dp(x), _dp(s)))
        return s

    def multiplicityLocal(self, nlocal):
        y = np.zeros(nlocal)
        self._chk(self.L.oracle_multiplicity_local(self.h, C.c_int64(nlocal), _dp(y)))
        return y

    def apply(self, x, nthreads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._chk(self.L.oracle_apply(self.h, C.c_int64(x.size), _dp(x), _dp(y), int(nthreads)))
        return y


def _mult(self, x, order):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    order = _i64(order)
    self._chk(self.L.oracle_apply_multiplicative(self.h, C.c_int64(x.size), _dp(x), _dp(y), C.c_int64(order.size), _ip(order)))
    return y


OraclePatchPC.applyMultiplicative = _mult


def max_threads():
    return lib().oracle_max_threads()
