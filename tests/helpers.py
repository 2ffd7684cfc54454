"""Shared set-up of an oracle PC and a GPU PC on the same synthetic problem."""
import numpy as np

from oracle.oracle import OraclePatchPC

OPT_NAMES = {"dim": "pc_patch_construction_dim", "codim": "pc_patch_construction_codim",
             "construct_type": "pc_patch_construction_type", "exclude_subspace": "pc_patch_exclude_subspace",
             "vankadim": "pc_patch_vanka_dim", "partition_of_unity": "pc_patch_partition_of_unity",
             "symmetrise_sweep": "pc_patch_symmetrise_sweep"}


def make_oracle(disc, A=None, **opts):
    o = OraclePatchPC()
    o.setFromDiscretisation(disc)
    o.setOptions(**opts)
    o.setUpPatches()
    if A is not None:
        o.setOperatorCSR(A)
        o.setUpOperators(keep_mat=True)
    return o


def make_gpu(form, bcs, device=0, **opts):
    from ssc_b200 import PC
    from ssc_b200.patch import setup_patch_pc
    pc = PC(device=device)
    setup_patch_pc(pc, form, bcs)
    for k, v in opts.items():
        if k == "construct_type":
            v = ["star", "vanka", "user"][v]
        pc.setOption(OPT_NAMES.get(k, k), v)
    return pc


def relerr(a, b):
    """Relative error of a against b: the worse of the 2-norm ratio and the max-norm ratio, so that no single entry may
    be off by more than the tolerance times the largest entry (a 2-norm alone would let a few entries hide)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    e2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    emax = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    return max(e2, emax)


def entry_err(a, b, scale):
    """Per-entry error: max_i |a_i - b_i| / scale_i, where scale_i = sum over the patches p containing dof i of
    max|B_p^-1 R_p x| (OraclePatchPC.applyScale) -- the magnitude the patch solves that write into entry i work at.
    A patch solve is accurate to a few ulps of ITS largest entry, so this is the sharpest bound that holds entry by entry:
    entries near Dirichlet faces, whose patches have small solutions, are checked at their own small scale.  Entries
    with scale 0 (Dirichlet dofs: passed through) must agree exactly."""
    a, b, scale = (np.asarray(v, dtype=np.float64) for v in (a, b, scale))
    d = np.abs(a - b)
    zero = scale == 0
    if np.any(d[zero] != 0):
        return np.inf
    return float((d[~zero] / scale[~zero]).max()) if np.any(~zero) else 0.0


class OraclePythonPC(object):
    """The oracle behind the same python-PC interface as ssc.PatchPC, for Krylov iteration-count parity."""

    def __init__(self, disc, A, **opts):
        self.o = make_oracle(disc, A, **opts)

    def setUp(self, pc):
        pass

    def apply(self, pc, x, y):
        y.array[:] = self.o.apply(x.array)

    def view(self, pc, viewer=None):
        pass


def solve(A, b, ksp_type, pc_setup, prefix="", history=True, dm=None, P=None):
    """One KSP solve with the stand-in driver; pc_setup(pc) configures the PC.  Returns (x, its, history)."""
    from ssc_b200 import petsc_lite as pl
    ksp = pl.KSP()
    ksp.setOptionsPrefix(prefix)
    if dm is not None:
        ksp.setDM(dm)
    ksp.setOperators(A, P)
    ksp.type = ksp_type
    pc_setup(ksp.getPC())
    if history:
        ksp.setConvergenceHistory()
    x = pl.Vec(np.zeros(b.size))
    ksp.solve(pl.Vec(b), x)
    return x.array, ksp.getIterationNumber(), ksp.getConvergenceHistory()
