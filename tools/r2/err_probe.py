"""Round-2 probe: actual GPU-vs-oracle errors (2-norm and per-entry) of the p sweep and the pivoting blocks."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_gpu, make_oracle, relerr
from ssc_b200 import fem, plex

def entry_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()

for p in range(1, 9):
    mesh = plex.UnitSquareMesh(6, 6)
    form, bcs = fem.poisson_problem(mesh, p, "scalar")
    disc = fem.discretisation_from_form(form, bcs)
    A = form.assemble(bcs)
    o = make_oracle(disc, A)
    x = np.random.default_rng(20260101).uniform(-1, 1, disc.total_dofs)
    ref = o.apply(x)
    for sub in ("lu", "cholesky"):
        pc = make_gpu(form, bcs)
        pc.setOption("sub_pc_type", sub)
        pc.setUp()
        y = np.zeros_like(x)
        pc.apply(x, y)
        rel_pt = np.abs(y - ref) / np.maximum(np.abs(ref), 1e-300)
        print("p=%d %s n_max=%d relerr2=%.2e maxabs/maxref=%.2e worst pointwise rel=%.2e (|ref| there %.2e)" % (
            p, sub, np.diff(pc.getPatches()[0]).max(), relerr(y, ref), entry_err(y, ref), rel_pt.max(), np.abs(ref)[rel_pt.argmax()]), flush=True)
