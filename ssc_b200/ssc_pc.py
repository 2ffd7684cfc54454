"""``ssc.SSC``: patch PC + P1 coarse PC combined additively (mirror of ssc/ssc.py:8-67).

The reference builds a PETSc PCCOMPOSITE with sub-PC 0 = "patch" and sub-PC 1 = python(P1PC) under the prefix
``<prefix>ssc_`` and leaves the composite type at PETSc's default (additive).  Here the composite is spelled out:
``y = patch(x) + lo(x)``; sub-PC options live under ``ssc_sub_0_`` and ``ssc_sub_1_`` as in
tests/test_p_independence.py:48-59.
"""
import numpy as np

from . import petsc_lite
from .lo import P1PC
from .patch import PCBase, _form_and_bcs, setup_patch_pc
from .pcpatch import PC as PatchPCCore, DeviceVec


class SSC(PCBase):

    device = 0

    def initialize(self, pc):
        A, P, J, bcs, ctx = _form_and_bcs(pc)
        mesh = J.ufl_domain()
        if getattr(getattr(mesh, "cell_set", None), "_extruded", False):
            raise NotImplementedError("Not implemented on extruded meshes")      # ssc/ssc.py:35-36
        prefix = (pc.getOptionsPrefix() or "") + "ssc_"                          # ssc/ssc.py:39
        patch = PatchPCCore(device=self.device)
        patch.setOptionsPrefix(prefix + "sub_0_")
        patch.setOperators(A, P)
        patch = setup_patch_pc(patch, J, bcs)                                    # ssc/ssc.py:44-45
        patch.setFromOptions()
        self.patch = patch
        self.lo = P1PC(J, bcs)                                                   # ssc/ssc.py:46-47
        self.lo.device = self.device
        self.lo_pc = petsc_lite.PC()
        self.lo_pc.setOptionsPrefix(prefix + "sub_1_")
        self.lo_pc.setOperators(A, P)
        self.lo_pc.setType("python")
        self.lo_pc.setPythonContext(self.lo)
        self._work = None

    def update(self, pc):
        self.patch.setUp()                                                       # combiner.setUp(), ssc/ssc.py:51-52
        self.lo_pc.setUp()

    def apply(self, pc, x, y):
        """PCCOMPOSITE additive (PETSc's default, ssc/ssc.py:38-49): y = patch(x) + lo(x).  The residual goes to the device
        once, both sub-PCs run there (the patch PC on its stream, the coarse PC beside it on its own), the sum is formed on
        the device and the result comes back once."""
        xa = x.array if hasattr(x, "array") else x
        ya = y.array if hasattr(y, "array") else y
        if self._work is None or self._work[0].n != xa.size:
            self._work = tuple(DeviceVec(self.patch, xa.size) for _ in range(3))
        dx, dy, dw = self._work
        dx.set(xa)
        self.applyDevice(dx, dy, dw)
        ya[:] = dy.get()

    def applyDevice(self, dx, dy, dw):
        """The same on device-resident vectors (dw: work vector); returns when dy is complete."""
        self.patch.apply(dx, dy)
        self.lo.lo.applyDevice(dx, dw)
        self.patch.synchronize()
        self.lo.lo.axpyDevice(1.0, dw, dy)
        self.lo.lo.synchronize()

    def applyTranspose(self, pc, x, y):
        self.patch.applyTranspose(x, y)

    def view(self, pc, viewer=None):
        if viewer is not None:
            viewer.printfASCII("Subspace correction PC\n")
            viewer.printfASCII("Combining (composite) PC:\n")
        self.patch.view(viewer)
        self.lo.view(self.lo_pc, viewer)
