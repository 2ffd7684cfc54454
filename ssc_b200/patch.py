"""``ssc.PatchPC``: the user-facing python PC (mirror of ssc/patch.py:184-279).

``-pc_type python -pc_python_type ssc.PatchPC`` makes PETSc instantiate this class with no arguments and call
``setUp`` (firedrake.PCBase: ``initialize`` once, then ``update``), ``apply`` and ``view``.  ``initialize`` finds
the form and boundary conditions exactly where the reference looks (python context of P, else the DM's appctx),
creates the inner PC of type "patch" with prefix ``<prefix>patch_`` and feeds it the mesh, the cell numbering,
the per-subspace sections / block sizes / cell-node maps, the Dirichlet dof lists and an operator source.

Difference from the reference, by design: the inner PC does not assemble patch matrices through a
per-patch callback into PETSc Mats (ssc/patch.py:210-213); it is given the assembled local operator in CSR and
cuts the dense patch blocks out of it on the GPU.
"""

import numpy as np

from . import fem, petsc_lite
from .pcpatch import PC as PatchPCCore

try:                                     # pragma: no cover - petsc4py is not installed in this build's containers
    from firedrake import PCBase
except ImportError:
    PCBase = petsc_lite.PCBase


def user_construction_op(pc, *args, **kwargs):
    """ssc/patch.py:169-181: resolve -pc_patch_construction_python_type to a callable returning (patches, order)."""
    prefix = pc.getOptionsPrefix()
    opts = pc.options_db(prefix) if pc.options_db is not None else petsc_lite.Options(prefix)
    usercode = opts.getString("pc_patch_construction_python_type", None)
    if usercode is None:
        raise ValueError("Must set %spc_patch_construction_python_type" % prefix)
    modname, funname = usercode.rsplit(".", 1)
    mod = __import__(modname, fromlist=[funname])
    fun = getattr(mod, funname)
    if isinstance(fun, type):
        fun = fun()
    return fun(pc, *args, **kwargs)


def _csr_of(J, bcs, P):
    """Assembled LOCAL operator as (indptr, indices, data): from P if it is assembled, else by assembling J.

    The extract kernel indexes rows and columns by local dof numbers (owned + ghost, the numbering of the dof lists) and
    needs the ghost rows as well.  ``Mat.getValuesCSR`` of a parallel (MPIAIJ) matrix returns the owned rows only, with
    GLOBAL column indices: this adapter is therefore for a sequential (one-rank) P; across ranks hand the local operator
    over explicitly (``setOperatorCSR`` with the locally assembled matrix, as bench.py does) or use the element-matrix
    route (``setElementMatrices``), which needs no assembled operator.  A matrix that is not square over the local dofs is
    rejected here rather than cut wrongly."""
    if P is not None and hasattr(P, "getType") and P.getType() != "python" and hasattr(P, "getValuesCSR"):
        comm_size = P.getComm().getSize() if hasattr(P, "getComm") and hasattr(P.getComm(), "getSize") else 1
        if comm_size > 1:
            raise ValueError("Mat.getValuesCSR of a parallel matrix gives owned rows with global columns; pass the locally "
                             "assembled operator with setOperatorCSR or use setElementMatrices")
        indptr, indices, data = P.getValuesCSR()
        if len(indices) and int(np.max(indices)) >= 2 ** 31:
            raise ValueError("column index %d does not fit the device's 32-bit dof index" % int(np.max(indices)))
        return indptr, indices, data
    A = J.assemble(bcs)
    return A.indptr, A.indices, A.data


def setup_patch_pc(patch, J, bcs):
    """ssc/patch.py:184-226."""
    patch = PatchPCCore.cast(patch)
    disc = J if isinstance(J, fem.Discretisation) else fem.discretisation_from_form(J, bcs)
    mesh = disc.mesh

    def op(pc):
        _, P = pc.getOperators()
        return _csr_of(J, bcs, P)

    patch.setDM(mesh)                                                         # :214
    patch.setPatchCellNumbering((disc.cell_numbering_dof, disc.cell_numbering_off))   # :215
    patch.setPatchDiscretisationInfo([(W.sec_dof, W.sec_off) for W in disc.spaces],   # :217-223
                                     disc.bs,
                                     [W.cell_node_list for W in disc.spaces],
                                     disc.offsets,
                                     disc.ghost_bc_nodes,
                                     disc.global_bc_nodes)
    patch.setPatchComputeOperator(op)                                         # :224
    patch.setPatchUserConstructionOperator(user_construction_op)              # :225
    patch.discretisation = disc
    return patch


def _form_and_bcs(pc):
    """ssc/patch.py:232-253."""
    A, P = pc.getOperators()
    if P.getType() == "python":
        ctx = P.getPythonContext()
        if ctx is None:
            raise ValueError("No context found on matrix")
        if not hasattr(ctx, "a") or not hasattr(ctx, "row_bcs"):
            raise ValueError("Don't know how to get form from %r" % (ctx,))
        J = ctx.a
        bcs = ctx.row_bcs
        if bcs != ctx.col_bcs:
            raise NotImplementedError("Row and column bcs must match")
    else:
        ctx = petsc_lite.get_appctx(pc.getDM()) if not hasattr(pc.getDM(), "getAppCtx") else pc.getDM().getAppCtx()
        if ctx is None:
            raise ValueError("No context found on form")
        if not hasattr(ctx, "J") or not hasattr(ctx, "_problem"):
            raise ValueError("Don't know how to get form from %r" % (ctx,))
        J = ctx.Jp or ctx.J
        bcs = ctx._problem.bcs
    return A, P, J, bcs, ctx


class PatchPC(PCBase):

    device = 0          # CUDA device of this rank (one process per GPU)

    def initialize(self, pc):
        A, P, J, bcs, ctx = _form_and_bcs(pc)
        mesh = J.ufl_domain()
        if getattr(getattr(mesh, "cell_set", None), "_extruded", False):
            raise NotImplementedError("Not implemented on extruded meshes")          # ssc/patch.py:256-257

        patch = PatchPCCore(device=self.device, comm=getattr(pc, "comm", None))
        patch.setOptionsPrefix((pc.getOptionsPrefix() or "") + "patch_")             # :260
        patch.setOperators(A, P)
        patch.setType("patch")
        patch = setup_patch_pc(patch, J, bcs)
        patch.setAttr("ctx", ctx)
        patch.incrementTabLevel(1, parent=pc)
        patch.setFromOptions()                                                        # :266
        self.patch = patch

    def update(self, pc):
        self.patch.setUp()

    def apply(self, pc, x, y):
        self.patch.apply(x, y)

    def applyTranspose(self, pc, x, y):
        self.patch.applyTranspose(x, y)

    def view(self, pc, viewer=None):
        self.patch.view(viewer=viewer)
