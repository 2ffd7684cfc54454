"""ctypes binding of the C ABI in include/ssc_b200.h.

Plays the part of the ``cdef extern from "libssc.h"`` block of ssc/PatchPC.pyx:14-28 and of its
``CHKERR`` (ssc/PatchPC.pyx:42-47): a non-zero return code becomes a ``RuntimeError`` carrying the code.
There is no Python or CPU fallback: if the shared library is missing, importing this module fails.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssc_b200.so")

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
MEM_HOST, MEM_DEVICE = 0, 1

if not os.path.exists(LIB_PATH):
    raise ImportError("%s is missing: build it with `python ssc_b200/build.py` (nvcc, sm_100a). "
                      "ssc_b200 has no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)
lib.SSCGetLastError.restype = C.c_char_p

_H = C.c_void_p
_sigs = {
    "SSCDeviceCount": [C.POINTER(C.c_int)],
    "SSCPatchInitializePackage": [],
    "SSCPatchCreate": [C.c_int, C.POINTER(_H)],
    "SSCPatchReset": [_H],
    "SSCPatchDestroy": [C.POINTER(_H)],
    "SSCPatchSetOption": [_H, C.c_char_p, C.c_char_p],
    "SSCPatchSetSaveOperators": [_H, C.c_int],
    "SSCPatchSetPartitionOfUnity": [_H, C.c_int],
    "SSCPatchSetSubMatType": [_H, C.c_char_p],
    "SSCPatchSetPlex": [_H, C.c_int64, C.c_int64, i64p, i64p, i64p, i64p, C.c_void_p],
    "SSCPatchSetCellNumbering": [_H, i64p, i64p],
    "SSCPatchSetDiscretisationInfo": [_H, C.c_int64, C.POINTER(i64p), C.POINTER(i64p), i64p, i64p, C.POINTER(i64p), i64p,
                                      C.c_int64, i64p, C.c_int64, i64p],
    "SSCPatchSetUserPatches": [_H, C.c_int64, i64p, i64p, C.c_int64, i64p],
    "SSCPatchSetPatches": [_H, C.c_int64, i64p, i64p, C.c_int64, C.c_int64, C.c_int64, i64p],
    "SSCPatchSetStarForest": [_H, C.c_int64, C.c_int64, i64p, i64p, C.c_int64, i32p, i64p, i64p, i64p, i64p],
    "SSCCommGetUniqueId": [C.c_char_p],
    "SSCPatchCommInit": [_H, C.c_int, C.c_int, C.c_char_p],
    "SSCPatchPeerExport": [_H, C.c_char_p],
    "SSCPatchPeerConnect": [_H, C.c_char_p, i32p, i64p, i64p],
    "SSCPatchSetOperatorCSR": [_H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int],
    "SSCPatchSetOperatorValues": [_H, C.c_void_p],
    "SSCPatchSetElementMatrices": [_H, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int],
    "SSCPatchReleaseOperator": [_H],
    "SSCPatchSetUp": [_H],
    "SSCPatchApply": [_H, C.c_void_p, C.c_void_p, C.c_int],
    "SSCPatchApplyTranspose": [_H, C.c_void_p, C.c_void_p, C.c_int],
    "SSCPatchSynchronize": [_H],
    "SSCPatchView": [_H, C.c_char_p, C.c_int64],
    "SSCPatchGetSize": [_H, C.c_char_p, i64p],
    "SSCPatchBuildPatches": [_H],
    "SSCPatchGetPrintout": [_H, C.c_char_p, C.c_int64, i64p],
    "SSCPatchGetPatches": [_H, i64p, i64p],
    "SSCPatchGetCells": [_H, i64p, i64p],
    "SSCPatchGetDofs": [_H, i64p],
    "SSCPatchGetPatchMatrix": [_H, C.c_int64, C.c_int, f64p],
    "SSCPatchGetDofWeights": [_H, f64p],
    "SSCPatchGetFailedPatches": [_H, i64p, C.c_int64, i64p],
    "SSCPatchGetColours": [_H, i64p],
    "SSCPatchGetEventTime": [_H, C.c_char_p, f64p],
    "SSCPatchKernelTimerReset": [_H, C.c_int],
    "SSCPatchKernelTimerRead": [_H, f64p, i64p],
    "SSCDeviceMalloc": [_H, C.c_int64, C.POINTER(C.c_void_p)],
    "SSCDeviceFree": [_H, C.c_void_p],
    "SSCHostAlloc": [_H, C.c_int64, C.POINTER(C.c_void_p)],
    "SSCHostFree": [_H, C.c_void_p],
    "SSCMemcpy": [_H, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int],
    "SSCDeviceFlushL2": [_H],
    "SSCDeviceMeasureReadBandwidth": [_H, C.c_int, f64p],
    "SSCEventCreate": [_H, C.POINTER(C.c_void_p)],
    "SSCEventRecord": [_H, C.c_void_p],
    "SSCEventElapsed": [_H, C.c_void_p, C.c_void_p, f64p],
    "SSCEventDestroy": [_H, C.c_void_p],
    "SSCLowOrderCreate": [C.c_int, C.POINTER(_H)],
    "SSCLowOrderDestroy": [C.POINTER(_H)],
    "SSCLowOrderSetOption": [_H, C.c_char_p, C.c_char_p],
    "SSCLowOrderSetRestriction": [_H, C.c_int64, C.c_int64, i64p, i32p, f64p],
    "SSCLowOrderSetOperator": [_H, C.c_int64, i64p, i32p, f64p],
    "SSCLowOrderSetBcs": [_H, C.c_int64, i64p],
    "SSCLowOrderSetNearNullSpace": [_H, C.c_int64, C.c_int, f64p],
    "SSCLowOrderSetUp": [_H],
    "SSCLowOrderApply": [_H, C.c_void_p, C.c_void_p, C.c_int],
    "SSCLowOrderGetIterations": [_H, i64p, i64p, i64p],
    "SSCLowOrderAXPY": [_H, C.c_int64, C.c_double, C.c_void_p, C.c_void_p],
    "SSCLowOrderSynchronize": [_H],
}
EXPORTS = sorted(_sigs) + ["SSCGetLastError"]
for _name, _args in _sigs.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = C.c_int


class SSCError(RuntimeError):
    """Raised like the reference's RuntimeError(ierr) (ssc/PatchPC.pyx:36-40); args = (code, message)."""

    def __init__(self, code, message):
        RuntimeError.__init__(self, code, message)
        self.ierr = code
        self.message = message

    def __str__(self):
        return "error code %d: %s" % (self.ierr, self.message)


def CHKERR(ierr):
    if ierr != 0:
        raise SSCError(ierr, lib.SSCGetLastError().decode("utf-8", "replace"))
    return 0


def device_count():
    n = C.c_int(0)
    CHKERR(lib.SSCDeviceCount(C.byref(n)))
    return n.value


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ip(a):
    return a.ctypes.data_as(i64p)


def dp(a):
    return a.ctypes.data_as(f64p)
