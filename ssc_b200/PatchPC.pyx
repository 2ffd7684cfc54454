# cython: language_level=3, boundscheck=False, wraparound=False
"""Cython shim over the C ABI of include/ssc_b200.h -- the counterpart of ssc/PatchPC.pyx in the reference.

The reference's PatchPC.pyx subclasses petsc4py.PETSc.PC and marshals numpy arrays into libssc's setters
(ssc/PatchPC.pyx:101-170), with ``CHKERR`` turning a non-zero return code into ``RuntimeError(ierr)`` (:36-47).
petsc4py is not available to this build, so the class here owns the opaque SSCPatchPC handle itself; the methods keep
the reference's names and argument meaning.  The extern block is declared ``nogil`` like the reference's (:6, :14), and
``apply`` releases the GIL around the call.
"""
import numpy
cimport numpy
from libc.stdint cimport int64_t, int32_t, uintptr_t
from libc.stdlib cimport malloc, free

cdef extern from "ssc_b200.h" nogil:
    ctypedef int64_t SSCInt
    ctypedef struct SSCPatchPC_s
    ctypedef SSCPatchPC_s* SSCPatchPC
    const char *SSCGetLastError()
    int SSCPatchInitializePackage()
    int SSCPatchCreate(int device, SSCPatchPC *pc)
    int SSCPatchDestroy(SSCPatchPC *pc)
    int SSCPatchSetOption(SSCPatchPC pc, const char *name, const char *value)
    int SSCPatchSetPlex(SSCPatchPC pc, SSCInt dim, SSCInt pEnd, const SSCInt *coneOff, const SSCInt *cones,
                        const SSCInt *depthStart, const SSCInt *depthEnd, const unsigned char *ghost)
    int SSCPatchSetCellNumbering(SSCPatchPC pc, const SSCInt *dof, const SSCInt *off)
    int SSCPatchSetDiscretisationInfo(SSCPatchPC pc, SSCInt nsubspaces, const SSCInt *const *secDof,
                                      const SSCInt *const *secOff, const SSCInt *bs, const SSCInt *nodesPerCell,
                                      const SSCInt *const *cellNodeMap, const SSCInt *subspaceOffsets,
                                      SSCInt numGhostBcs, const SSCInt *ghostBcNodes,
                                      SSCInt numGlobalBcs, const SSCInt *globalBcNodes)
    int SSCPatchSetOperatorCSR(SSCPatchPC pc, SSCInt nrows, const SSCInt *rowptr, const int32_t *colidx,
                               const double *vals, int memspace)
    int SSCPatchSetUp(SSCPatchPC pc)
    int SSCPatchApply(SSCPatchPC pc, const double *x, double *y, int memspace)
    int SSCPatchApplyTranspose(SSCPatchPC pc, const double *x, double *y, int memspace)
    int SSCPatchSynchronize(SSCPatchPC pc)


from ssc_b200._lib import SSCError      # RuntimeError(ierr, message), cf. ssc/PatchPC.pyx:36-40


cdef inline int SETERR(int ierr) except -1 with gil:
    raise SSCError(ierr, SSCGetLastError().decode("utf-8", "replace"))


cdef inline int CHKERR(int ierr) except -1 nogil:
    if ierr == 0:
        return 0  # no error
    return SETERR(ierr)


cdef class PC:
    """Owner of one SSCPatchPC handle; ``handle`` is the pointer as an integer (shared with the ctypes introspection calls)."""
    cdef SSCPatchPC pc
    cdef list keep          # arrays the library borrows for the PC's lifetime (cellNodeMap, ssc/libssc.c:268)
    cdef object csr_keep

    def __cinit__(self, int device=0):
        self.pc = NULL
        self.keep = []
        self.csr_keep = None
        CHKERR( SSCPatchInitializePackage() )
        CHKERR( SSCPatchCreate(device, &self.pc) )

    def __dealloc__(self):
        if self.pc != NULL:
            SSCPatchDestroy(&self.pc)

    def destroy(self):
        if self.pc != NULL:
            CHKERR( SSCPatchDestroy(&self.pc) )

    property handle:
        def __get__(self):
            return <uintptr_t>self.pc

    def setOption(self, name, value):
        cdef bytes n = name.encode(), v = str(value).encode()
        CHKERR( SSCPatchSetOption(self.pc, n, v) )

    def setPlex(self, SSCInt dim,
                numpy.ndarray[SSCInt, ndim=1, mode="c"] coneOff,
                numpy.ndarray[SSCInt, ndim=1, mode="c"] cones,
                numpy.ndarray[SSCInt, ndim=1, mode="c"] depthStart,
                numpy.ndarray[SSCInt, ndim=1, mode="c"] depthEnd,
                numpy.ndarray[numpy.uint8_t, ndim=1, mode="c"] ghost):
        CHKERR( SSCPatchSetPlex(self.pc, dim, coneOff.shape[0] - 1, <SSCInt*>coneOff.data, <SSCInt*>cones.data,
                                <SSCInt*>depthStart.data, <SSCInt*>depthEnd.data, <unsigned char*>ghost.data) )

    def setPatchCellNumbering(self,
                              numpy.ndarray[SSCInt, ndim=1, mode="c"] dof,
                              numpy.ndarray[SSCInt, ndim=1, mode="c"] off):
        CHKERR( SSCPatchSetCellNumbering(self.pc, <SSCInt*>dof.data, <SSCInt*>off.data) )

    def setPatchDiscretisationInfo(self, sections,
                                   numpy.ndarray[SSCInt, ndim=1, mode="c"] bs,
                                   cellNodeMaps,
                                   numpy.ndarray[SSCInt, ndim=1, mode="c"] subspaceOffsets,
                                   numpy.ndarray[SSCInt, ndim=1, mode="c"] ghostBcNodes,
                                   numpy.ndarray[SSCInt, ndim=1, mode="c"] globalBcNodes):
        # same marshalling as ssc/PatchPC.pyx:113-156, with each DM replaced by its section's (dof, offset) arrays
        cdef:
            SSCInt numSubSpaces = bs.shape[0]
            SSCInt *cnodesPerCell = <SSCInt*>malloc(numSubSpaces * sizeof(SSCInt))
            const SSCInt **ccellNodeMaps = <const SSCInt**>malloc(numSubSpaces * sizeof(SSCInt*))
            const SSCInt **csecDof = <const SSCInt**>malloc(numSubSpaces * sizeof(SSCInt*))
            const SSCInt **csecOff = <const SSCInt**>malloc(numSubSpaces * sizeof(SSCInt*))
            SSCInt i
            numpy.ndarray[SSCInt, ndim=2, mode="c"] tmp
            numpy.ndarray[SSCInt, ndim=1, mode="c"] d
            numpy.ndarray[SSCInt, ndim=1, mode="c"] o
            int ierr
        self.keep = [cellNodeMaps, sections]
        for i in range(numSubSpaces):
            tmp = <numpy.ndarray?>(cellNodeMaps[i])
            ccellNodeMaps[i] = <SSCInt*>tmp.data
            cnodesPerCell[i] = tmp.shape[1]
            d = <numpy.ndarray?>(sections[i][0])
            o = <numpy.ndarray?>(sections[i][1])
            csecDof[i] = <SSCInt*>d.data
            csecOff[i] = <SSCInt*>o.data
        ierr = SSCPatchSetDiscretisationInfo(self.pc, numSubSpaces, csecDof, csecOff, <SSCInt*>bs.data, cnodesPerCell,
                                             ccellNodeMaps, <SSCInt*>subspaceOffsets.data,
                                             ghostBcNodes.shape[0], <SSCInt*>ghostBcNodes.data,
                                             globalBcNodes.shape[0], <SSCInt*>globalBcNodes.data)
        free(cnodesPerCell)
        free(ccellNodeMaps)
        free(csecDof)
        free(csecOff)
        CHKERR( ierr )

    def setOperatorCSR(self,
                       numpy.ndarray[SSCInt, ndim=1, mode="c"] indptr,
                       numpy.ndarray[int32_t, ndim=1, mode="c"] indices,
                       numpy.ndarray[double, ndim=1, mode="c"] data):
        self.csr_keep = (indptr, indices, data)      # borrowed by the library until the next setUp returns
        CHKERR( SSCPatchSetOperatorCSR(self.pc, indptr.shape[0] - 1, <SSCInt*>indptr.data, <int32_t*>indices.data,
                                       <double*>data.data, 0) )

    def releaseOperator(self):
        self.csr_keep = None

    def setUp(self):
        cdef int ierr
        with nogil:
            ierr = SSCPatchSetUp(self.pc)
        CHKERR( ierr )

    def apply(self, numpy.ndarray[double, ndim=1, mode="c"] x, numpy.ndarray[double, ndim=1, mode="c"] y):
        cdef int ierr
        cdef const double *xp = <const double*>x.data
        cdef double *yp = <double*>y.data
        with nogil:
            ierr = SSCPatchApply(self.pc, xp, yp, 0)
        CHKERR( ierr )

    def applyDevice(self, uintptr_t x, uintptr_t y):
        cdef int ierr
        with nogil:
            ierr = SSCPatchApply(self.pc, <const double*>x, <double*>y, 1)
        CHKERR( ierr )

    def applyTranspose(self):
        CHKERR( SSCPatchApplyTranspose(self.pc, NULL, NULL, 0) )

    def synchronize(self):
        cdef int ierr
        with nogil:
            ierr = SSCPatchSynchronize(self.pc)
        CHKERR( ierr )
