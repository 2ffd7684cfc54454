"""User patch constructors of ssc/relaxation.py through -pc_patch_construction_type python: the product's host builder
and (on the GPU) PCApply against the oracle fed with the same point sets and iteration set."""
import numpy as np
import pytest

from conftest import rhs_vector
from helpers import relerr
from oracle.oracle import OraclePatchPC
from ssc_b200 import PC, fem, plex
from ssc_b200 import petsc_lite as pl
from ssc_b200.patch import setup_patch_pc

CASES = [
    ("ssc.PlaneSmoother", {"pc_patch_construction_ps_sweeps": "0+3:1-2"}),
    ("ssc.OrderedStar", {"pc_patch_construction_ostar_nclosures": 1, "pc_patch_construction_ostar_sort_order": "0+:1-|1+"}),
    ("ssc.OrderedStar", {"pc_patch_construction_ostar_dim": 1}),
    ("ssc.OrderedVanka", {"pc_patch_construction_ovanka_nclosures": 1, "pc_patch_construction_ovanka_exclude_dim": 0}),
    ("ssc.OrderedVanka", {"pc_patch_construction_ovanka_codim": 0, "pc_patch_construction_ovanka_sort_order": "1-"}),
]


def _setup(device, cls, opts):
    mesh = plex.RectangleMesh(5, 4, 2.0, 1.0)
    form, bcs = fem.poisson_problem(mesh, 2, "vector")
    disc = fem.discretisation_from_form(form, bcs)
    A = form.assemble(bcs)
    pl.Options.clear()
    o_ = pl.Options("r_")
    o_.setValue("pc_patch_construction_type", "python")
    o_.setValue("pc_patch_construction_python_type", cls)
    for k, v in opts.items():
        o_.setValue(k, v)
    pc = PC(device=device)
    pc.setOptionsPrefix("r_")
    setup_patch_pc(pc, form, bcs)
    pc.setFromOptions()
    # the same constructor output for the oracle
    import ssc
    fun = getattr(ssc, cls.split(".")[1])()
    patches, order = fun(pc)
    pl.Options.clear()
    o = OraclePatchPC()
    o.setFromDiscretisation(disc)
    o.setUserPatches(patches, order)
    o.setUpPatches()
    return pc, o, disc, A


@pytest.mark.parametrize("cls,opts", CASES)
def test_builder_matches_oracle(cls, opts):
    pc, o, disc, A = _setup(-1, cls, opts)
    for a, b in zip(o.patches(), pc.getPatches()):
        assert np.array_equal(a, b)
    assert pc.getPatches()[0][-1] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("cls,opts", CASES)
def test_gpu_apply_matches_oracle(cls, opts):
    pc, o, disc, A = _setup(0, cls, opts)
    o.setOperatorCSR(A)
    o.setUpOperators()
    pc.setUp()
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < 1e-11        # plane patches are large (n up to ~150) and worse conditioned than stars
