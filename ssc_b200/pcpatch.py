"""The PC of type "patch" as the host sees it.

Mirror of ``cdef class PC(PETSc.PC)`` in ssc/PatchPC.pyx:101-170 (``setPatchCellNumbering``,
``setPatchDiscretisationInfo``, ``setPatchComputeOperator``, ``setPatchUserConstructionOperator``) together
with the petsc4py PC methods ssc/patch.py calls on it (``setOptionsPrefix``, ``setOperators``, ``setDM``,
``setFromOptions``, ``setUp``, ``apply``, ``applyTranspose``, ``view``).  Every method marshals numpy arrays
into the C ABI of include/ssc_b200.h; all arithmetic happens in CUDA kernels behind it.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CHKERR, MEM_DEVICE, MEM_HOST, as_i64, dp, ip, lib
import importlib
import sys

# the Cython shim ssc_b200/PatchPC.pyx.  Importing a submodule binds it as the package attribute `PatchPC`; that name belongs
# to the python class (ssc_b200.PatchPC, resolved lazily by the package): exactly the collision of ssc/patch.py:12 +
# ssc/__init__.py:2 in the reference, where the package rebinds the name to the class.
cPatchPC = importlib.import_module(".PatchPC", __package__)
if getattr(sys.modules[__package__], "PatchPC", None) is cPatchPC:
    delattr(sys.modules[__package__], "PatchPC")

# option names of PCSetFromOptions_PATCH (ssc/libssc.c:1551-1620) and of the per-patch KSP prefix "sub_" (:1240-1241)
PATCH_OPTIONS = ("pc_patch_save_operators", "pc_patch_partition_of_unity", "pc_patch_multiplicative",
                 "pc_patch_construction_dim", "pc_patch_construction_codim", "pc_patch_construction_type",
                 "pc_patch_vanka_dim", "pc_patch_vanka_space", "pc_patch_sub_mat_type", "pc_patch_print_patches",
                 "pc_patch_symmetrise_sweep", "pc_patch_exclude_subspace", "sub_ksp_type", "sub_pc_type",
                 "sub_pc_factor_mat_solver_type", "sub_pc_factor_mat_ordering_type",
                 "ssc_keep_dofs_table", "ssc_keep_patch_operators", "ssc_builder_threads", "ssc_apply_variant", "ssc_factor_variant", "ssc_factor_panel", "ssc_extract_row_buffers", "ssc_sub_solve", "ssc_register_host_vectors", "ssc_refine_kappa", "ssc_refine_steps", "ssc_factor_ctas", "ssc_host_pipeline", "ssc_deterministic", "ssc_device_patch_construction", "ssc_symmetric_storage", "ssc_overlap_classes", "ssc_apply_stream_small", "ssc_host_pipeline_taper", "ssc_apply_stream_min_waves", "ssc_apply_stream_ctas")


class DeviceVec(object):
    """A vector resident in HBM (owned dofs of one rank)."""

    def __init__(self, pc, n):
        self.pc = pc
        self.n = int(n)
        p = C.c_void_p()
        CHKERR(lib.SSCDeviceMalloc(pc.handle, C.c_int64(8 * max(self.n, 1)), C.byref(p)))
        self.ptr = p

    def set(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.n
        CHKERR(lib.SSCMemcpy(self.pc.handle, self.ptr, host.ctypes.data_as(C.c_void_p), C.c_int64(8 * self.n), MEM_DEVICE, MEM_HOST))

    def get(self):
        out = np.empty(self.n)
        CHKERR(lib.SSCMemcpy(self.pc.handle, out.ctypes.data_as(C.c_void_p), self.ptr, C.c_int64(8 * self.n), MEM_HOST, MEM_DEVICE))
        return out

    def free(self):
        if self.ptr:
            lib.SSCDeviceFree(self.pc.handle, self.ptr)
            self.ptr = None


class PinnedArray(object):
    """Page-locked host buffer viewed as a numpy array (what a PETSc Vec with pinned storage offers)."""

    def __init__(self, pc, n):
        self.pc = pc
        p = C.c_void_p()
        CHKERR(lib.SSCHostAlloc(pc.handle, C.c_int64(8 * max(int(n), 1)), C.byref(p)))
        self.ptr = p
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(int(n),))

    def free(self):
        if self.ptr:
            self.array = None
            lib.SSCHostFree(self.pc.handle, self.ptr)
            self.ptr = None


def _host_array(v):
    """numpy view of a Vec-like (petsc4py Vec, petsc_lite Vec, numpy array)."""
    if isinstance(v, np.ndarray):
        return v
    if hasattr(v, "array"):
        return v.array
    return v.getArray()


class PC(object):
    def __init__(self, device=0, comm=None):
        self._core = cPatchPC.PC(int(device))          # owns the SSCPatchPC handle
        self.handle = C.c_void_p(self._core.handle)    # same pointer, for the ctypes introspection calls
        self.device = device
        self.comm = comm
        self.prefix = ""
        self.A = self.P = None
        self.attrs = {}
        self._keep = []
        self._user_op = None
        self._csr_keep = None
        self.options_db = None

    # ---- lifetime -------------------------------------------------------------------------
    def destroy(self):
        if self.handle:
            self._core.destroy()
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def reset(self):
        CHKERR(lib.SSCPatchReset(self.handle))

    @staticmethod
    def cast(pc):                      # ssc/PatchPC.pyx:103-108
        return pc

    # ---- petsc4py PC surface used by ssc/patch.py:259-279 ---------------------------------
    def create(self, comm=None):
        self.comm = comm
        return self

    def setOptionsPrefix(self, prefix):
        self.prefix = prefix or ""

    def getOptionsPrefix(self):
        return self.prefix

    def setOperators(self, A, P=None):
        self.A, self.P = A, (A if P is None else P)

    def getOperators(self):
        return self.A, self.P

    def setType(self, t):
        if t != "patch":
            raise ValueError("this PC only has type 'patch'")

    def setDM(self, plex):
        """``patch.setDM(mesh._plex)`` (ssc/patch.py:214); ``plex`` is a ssc_b200.plex.Plex (or petsc4py DMPlex)."""
        if not hasattr(plex, "cone_off"):
            from .plex import Plex
            plex = Plex.from_dmplex(plex)
        arrs = [as_i64(a) for a in (plex.cone_off, plex.cones, plex.depth_start, plex.depth_end)]
        ghost = np.ascontiguousarray(plex.ghost, dtype=np.uint8)
        self._core.setPlex(plex.dim, arrs[0], arrs[1], arrs[2], arrs[3], ghost)
        self.plex = plex

    def getDM(self):
        return self.plex

    def setAttr(self, k, v):
        self.attrs[k] = v

    def getAttr(self, k):
        return self.attrs.get(k)

    set_attr, get_attr = setAttr, getAttr

    def incrementTabLevel(self, n, parent=None):
        pass

    def setOption(self, name, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        self._core.setOption(name, value)

    def setFromOptions(self):
        """PCSetFromOptions_PATCH (ssc/libssc.c:1551-1620): read every known name under this PC's prefix."""
        if self.options_db is None:
            try:
                from petsc4py import PETSc
                self.options_db = PETSc.Options
            except ImportError:
                from .petsc_lite import Options
                self.options_db = Options
        opts = self.options_db(self.prefix)
        for name in PATCH_OPTIONS:
            if opts.hasName(name):
                self.setOption(name, opts.getString(name))
        t = opts.getString("pc_patch_construction_type", None)
        if t in ("user", "python") and self._user_op is not None:
            self._run_user_construction()

    # ---- ssc/PatchPC.pyx:110-170 -----------------------------------------------------------
    def setPatchCellNumbering(self, sec):
        dof, off = (as_i64(sec[0]), as_i64(sec[1])) if isinstance(sec, tuple) else _section_arrays(sec, self.plex.pEnd)
        self._core.setPatchCellNumbering(dof, off)

    def setPatchDiscretisationInfo(self, dms, bs, cellNodeMaps, subspaceOffsets, ghostBcNodes, globalBcNodes):
        """``dms[i]`` is the (dof, offset) pair of the default section of subspace i (or a petsc4py DM)."""
        n = len(dms)
        secs = [(as_i64(d[0]), as_i64(d[1])) if isinstance(d, tuple) else _section_arrays(d.getDefaultSection(), self.plex.pEnd)
                for d in dms]
        cnm = [np.ascontiguousarray(m, dtype=np.int64) for m in cellNodeMaps]
        self._core.setPatchDiscretisationInfo(secs, as_i64(bs), cnm, as_i64(subspaceOffsets), as_i64(ghostBcNodes), as_i64(globalBcNodes))

    def setPatchComputeOperator(self, operator, args=None, kargs=None):
        """The reference's callback assembles one patch matrix (ssc/PatchPC.pyx:158-163).  Here ``operator`` is
        called once per setUp and returns the assembled local operator as (indptr, indices, data[, memspace])."""
        self._compute_op = (operator, args or (), kargs or {})

    def setPatchUserConstructionOperator(self, operator, args=None, kargs=None):
        self._user_op = (operator, args or (), kargs or {})

    def _run_user_construction(self):
        op, args, kargs = self._user_op
        patches, iteration_set = op(self, *args, **kargs)     # ssc/PatchPC.pyx:85
        patches = [as_i64(p.getIndices() if hasattr(p, "getIndices") else p) for p in patches]
        it = as_i64(iteration_set.getIndices() if hasattr(iteration_set, "getIndices") else iteration_set)
        off = as_i64(np.concatenate(([0], np.cumsum([p.size for p in patches]))))
        pts = as_i64(np.concatenate(patches)) if patches else as_i64([])
        CHKERR(lib.SSCPatchSetUserPatches(self.handle, C.c_int64(len(patches)), ip(off), ip(pts), C.c_int64(it.size), ip(it)))

    def setPatches(self, offsets, gtol, nlocal, nowned=None, global_bc_nodes=()):
        offsets, gtol, bc = as_i64(offsets), as_i64(gtol), as_i64(global_bc_nodes)
        CHKERR(lib.SSCPatchSetPatches(self.handle, C.c_int64(offsets.size - 1), ip(offsets), ip(gtol), C.c_int64(nlocal),
                                      C.c_int64(nlocal if nowned is None else nowned), C.c_int64(bc.size), ip(bc)))

    def setStarForest(self, nowned, self_leaf, self_root, neigh, send_off, send_root, recv_off, recv_leaf):
        a = [as_i64(v) for v in (self_leaf, self_root, send_off, send_root, recv_off, recv_leaf)]
        ng = np.ascontiguousarray(neigh, dtype=np.int32)
        CHKERR(lib.SSCPatchSetStarForest(self.handle, C.c_int64(nowned), C.c_int64(a[0].size), ip(a[0]), ip(a[1]), C.c_int64(ng.size),
                                         ng.ctypes.data_as(_lib.i32p), ip(a[2]), ip(a[3]), ip(a[4]), ip(a[5])))

    def peerExport(self):
        buf = C.create_string_buffer(192)
        CHKERR(lib.SSCPatchPeerExport(self.handle, buf))
        return buf.raw

    def peerConnect(self, neigh_handles, slot_at_neigh, remote_leaf, remote_root):
        slots = np.ascontiguousarray(slot_at_neigh, dtype=np.int32)
        rl, rr = as_i64(remote_leaf), as_i64(remote_root)
        CHKERR(lib.SSCPatchPeerConnect(self.handle, b"".join(neigh_handles), slots.ctypes.data_as(_lib.i32p), ip(rl), ip(rr)))

    def commInit(self, nranks, rank, unique_id):
        CHKERR(lib.SSCPatchCommInit(self.handle, int(nranks), int(rank), unique_id))

    def setOperatorCSR(self, indptr, indices, data, memspace=MEM_HOST):
        if memspace == MEM_HOST:
            indptr = as_i64(indptr)
            indices = np.ascontiguousarray(indices, dtype=np.int32)
            data = np.ascontiguousarray(data, dtype=np.float64)
            self._csr_keep = (indptr, indices, data)
            self._core.setOperatorCSR(indptr, indices, data)
            return
        n, ptrs = indptr[0], [C.c_void_p(p) for p in (indptr[1], indices, data)]
        CHKERR(lib.SSCPatchSetOperatorCSR(self.handle, C.c_int64(n), ptrs[0], ptrs[1], ptrs[2], int(memspace)))

    def setOperatorValues(self, data):
        """New values on the pattern of the last setOperatorCSR (host arrays): the next setUp uploads only them."""
        data = np.ascontiguousarray(data, dtype=np.float64)
        if getattr(self, "_csr_keep", None) is None or len(self._csr_keep) != 3 or data.size != self._csr_keep[2].size:
            raise ValueError("setOperatorValues needs the pattern of a previous setOperatorCSR of the same size")
        if lib.SSCPatchSetOperatorValues(self.handle, data.ctypes.data_as(C.c_void_p)) != 0:
            # no device copy of the pattern to reuse (not set up yet, or dropped because device memory is tight): the pattern arrays
            # are still held here, so the whole operator goes over again
            self.setOperatorCSR(self._csr_keep[0], self._csr_keep[1], data)
            return
        self._csr_keep = (self._csr_keep[0], self._csr_keep[1], data)

    def setElementMatrices(self, Ke, memspace=MEM_HOST, ncells=None, dofs_per_cell=None):
        """The reference's source of the patch operators (PCPatchSetComputeOperator, ssc/PatchPC.pyx:158-163 with the
        element kernel of ssc/patch.py:28-62): Ke[cell] is the element matrix of that cell, rows/columns in the order
        of the per-cell dof table; a 2-D Ke is one matrix shared by every cell.  Replaces setOperatorCSR."""
        if memspace == MEM_HOST:
            Ke = np.ascontiguousarray(Ke, dtype=np.float64)
            if Ke.ndim not in (2, 3) or Ke.shape[-1] != Ke.shape[-2]:
                raise ValueError("element matrices must be (ncells, t, t) or one shared (t, t) matrix")
            self._csr_keep = (Ke,)
            t = Ke.shape[-1]
            nc, stride, ptr = (1, 0, Ke.ctypes.data) if Ke.ndim == 2 else (Ke.shape[0], t * t, Ke.ctypes.data)
        else:
            t, nc = int(dofs_per_cell), int(ncells or 1)
            stride, ptr = (0 if ncells is None else t * t), int(Ke)
        CHKERR(lib.SSCPatchSetElementMatrices(self.handle, C.c_int64(nc), C.c_int64(t), C.c_void_p(ptr), C.c_int64(stride), int(memspace)))

    def releaseOperator(self):
        """Drop the host CSR arrays kept alive for the library (they are only borrowed until setUp returns)."""
        self._csr_keep = None
        self._core.releaseOperator()
        CHKERR(lib.SSCPatchReleaseOperator(self.handle))

    # ---- setup / apply ---------------------------------------------------------------------
    def setUp(self):
        """PCSetUp_PATCH (ssc/libssc.c:1201): re-reads the operator values on every call."""
        if getattr(self, "_compute_op", None) is not None:
            op, args, kargs = self._compute_op
            csr = op(self, *args, **kargs)
            self.setOperatorCSR(*csr)
        elif self._csr_keep is None and self.P is not None and hasattr(self.P, "getValuesCSR"):
            self.setOperatorCSR(*self.P.getValuesCSR())
        self._core.setUp()

    def apply(self, x, y):
        if isinstance(x, DeviceVec):
            self._core.applyDevice(x.ptr.value, y.ptr.value)
            return
        xa, ya = _host_array(x), _host_array(y)
        if not (xa.flags.c_contiguous and ya.flags.c_contiguous and xa.dtype == np.float64 and ya.dtype == np.float64):
            raise ValueError("x and y must be contiguous float64")
        if xa.size != self.size("nowned") or ya.size != xa.size:
            raise ValueError("vector length %d does not match the PC (%d)" % (xa.size, self.size("nowned")))
        self._core.apply(xa, ya)

    def applyTranspose(self, x, y):
        self._core.applyTranspose()

    def synchronize(self):
        self._core.synchronize()

    def view(self, viewer=None):
        buf = C.create_string_buffer(4096)
        CHKERR(lib.SSCPatchView(self.handle, buf, C.c_int64(4096)))
        text = "PC Object: type patch\n" + buf.value.decode()
        if viewer is None:
            print(text, end="")
        else:
            viewer.printfASCII(text)
        return text

    # ---- introspection ---------------------------------------------------------------------
    def size(self, what):
        v = C.c_int64(0)
        CHKERR(lib.SSCPatchGetSize(self.handle, what.encode(), C.byref(v)))
        return v.value

    def buildPatches(self):
        CHKERR(lib.SSCPatchBuildPatches(self.handle))

    def getPrintout(self):
        n = C.c_int64(0)
        CHKERR(lib.SSCPatchGetPrintout(self.handle, None, C.c_int64(0), C.byref(n)))
        buf = C.create_string_buffer(n.value)
        CHKERR(lib.SSCPatchGetPrintout(self.handle, buf, n, None))
        return buf.value.decode()

    def getPatches(self):
        self.buildPatches()
        off = np.zeros(self.size("npatch") + 1, dtype=np.int64)
        gtol = np.zeros(self.size("num_gtol"), dtype=np.int64)
        CHKERR(lib.SSCPatchGetPatches(self.handle, ip(off), ip(gtol)))
        return off, gtol

    def getCells(self):
        self.buildPatches()
        off = np.zeros(self.size("npatch") + 1, dtype=np.int64)
        cells = np.zeros(self.size("num_cells"), dtype=np.int64)
        CHKERR(lib.SSCPatchGetCells(self.handle, ip(off), ip(cells)))
        return off, cells

    def getDofs(self):
        self.buildPatches()
        d = np.zeros(self.size("num_dofs"), dtype=np.int64)
        CHKERR(lib.SSCPatchGetDofs(self.handle, ip(d)))
        return d

    def getPatchMatrix(self, p, inverse=False):
        off, _ = self.getPatches()
        n = int(off[p + 1] - off[p])
        out = np.zeros((n, n))
        CHKERR(lib.SSCPatchGetPatchMatrix(self.handle, C.c_int64(p), 1 if inverse else 0, dp(out)))
        return out

    def getColours(self):
        c = np.zeros(self.size("npatch"), dtype=np.int64)
        CHKERR(lib.SSCPatchGetColours(self.handle, ip(c)))
        return c

    def getFailedPatches(self):
        n = C.c_int64(0)
        CHKERR(lib.SSCPatchGetFailedPatches(self.handle, None, C.c_int64(0), C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.int64)
        CHKERR(lib.SSCPatchGetFailedPatches(self.handle, ip(out), C.c_int64(n.value), None))
        return out[:n.value]

    def getDofWeights(self):
        w = np.zeros(self.size("nowned"))
        CHKERR(lib.SSCPatchGetDofWeights(self.handle, dp(w)))
        return w

    def eventTime(self, name):
        v = C.c_double(0)
        CHKERR(lib.SSCPatchGetEventTime(self.handle, name.encode(), C.byref(v)))
        return v.value

    def kernelTimerReset(self, enable=True):
        CHKERR(lib.SSCPatchKernelTimerReset(self.handle, int(enable)))

    def kernelTimerRead(self):
        ms, n = C.c_double(0), C.c_int64(0)
        CHKERR(lib.SSCPatchKernelTimerRead(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def measureReadBandwidth(self, reps=5):
        v = C.c_double(0)
        CHKERR(lib.SSCDeviceMeasureReadBandwidth(self.handle, int(reps), C.byref(v)))
        return v.value

    def flushL2(self):
        CHKERR(lib.SSCDeviceFlushL2(self.handle))

    # CUDA events on the PC's stream
    def eventCreate(self):
        e = C.c_void_p()
        CHKERR(lib.SSCEventCreate(self.handle, C.byref(e)))
        return e

    def eventRecord(self, e):
        CHKERR(lib.SSCEventRecord(self.handle, e))

    def eventElapsed(self, a, b):
        ms = C.c_double(0)
        CHKERR(lib.SSCEventElapsed(self.handle, a, b, C.byref(ms)))
        return ms.value


def _section_arrays(sec, pEnd):
    """(dof, offset) arrays of a petsc4py Section over the chart [0, pEnd)."""
    dof = np.zeros(pEnd, dtype=np.int64)
    off = np.zeros(pEnd, dtype=np.int64)
    s, e = sec.getChart()
    for p in range(s, e):
        dof[p] = sec.getDof(p)
        off[p] = sec.getOffset(p)
    return dof, off


def get_unique_id():
    buf = C.create_string_buffer(128)
    CHKERR(lib.SSCCommGetUniqueId(buf))
    return buf.raw
