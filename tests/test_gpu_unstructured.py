"""GPU parity on unstructured (Delaunay) simplicial meshes: vertex valences vary, so the vertex-star patches come in
10-20 different sizes (ragged size classes, SURVEY.md 8(d) C4 "unstructured variant for variable n")."""
import numpy as np
import pytest

from conftest import delaunay_mesh, rhs_vector
from helpers import make_gpu, make_oracle, relerr
from ssc_b200 import fem

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _problem(mesh, p, kind):
    form, bcs = fem.poisson_problem(mesh, p, kind)
    disc = fem.discretisation_from_form(form, bcs)
    return form, bcs, disc, form.assemble(bcs)


@pytest.mark.parametrize("dim,npts,p,kind", [(2, 400, 3, "scalar"), (2, 150, 2, "mixed"), (3, 250, 2, "vector"), (3, 120, 3, "scalar")])
@pytest.mark.parametrize("opts", [{}, {"partition_of_unity": 1}, {"sub_pc_type": "cholesky", "ssc_symmetric_storage": True}, {"ssc_deterministic": True}],
                         ids=["default", "pou", "packed", "deterministic"])
def test_apply_matches_oracle_on_delaunay_meshes(dim, npts, p, kind, opts):
    form, bcs, disc, A = _problem(delaunay_mesh(dim, npts), p, kind)
    oopts = {k: v for k, v in opts.items() if k == "partition_of_unity"}
    o = make_oracle(disc, A, **oopts)
    pc = make_gpu(form, bcs, **opts)
    pc.setUp()
    assert all(np.array_equal(a, b) for a, b in zip(o.patches(), pc.getPatches()))
    sizes = np.diff(pc.getPatches()[0])
    assert np.unique(sizes[sizes > 0]).size >= 6
    for seed in (20260101, 20260102):
        x = rhs_vector(disc.total_dofs, seed)
        y = np.zeros_like(x)
        pc.apply(x, y)
        assert relerr(y, o.apply(x)) < TOL


def test_element_route_bit_identical_on_delaunay_mesh():
    mesh = delaunay_mesh(3, 150, seed=9)
    form, bcs, disc, A = _problem(mesh, 2, "scalar")
    Ke = fem.poisson_element_matrices(form.spaces)
    o = make_oracle(disc, A)
    pc = make_gpu(form, bcs, ssc_keep_patch_operators=True)
    pc._compute_op = None
    pc.setElementMatrices(Ke)
    pc.setUp()
    off, _ = pc.getPatches()
    for i in range(0, off.size - 1, 7):
        if off[i + 1] > off[i]:
            assert np.array_equal(pc.getPatchMatrix(i), o.patchMatrixFromElements(i, Ke))
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.apply(x)) < 1e-11


def test_multiplicative_sweep_on_delaunay_mesh():
    """Coloured Gauss-Seidel sweep over ragged patches against the oracle's sweep in the same colour order."""
    form, bcs, disc, A = _problem(delaunay_mesh(2, 200, seed=4), 2, "scalar")
    pc = make_gpu(form, bcs, pc_patch_multiplicative=True)
    pc.setUp()
    col = pc.getColours()
    ncol = pc.size("num_colours")
    coff, cells = pc.getCells()
    for c in range(ncol):                                  # no two patches of a colour share a cell
        seen = np.concatenate([cells[coff[q]:coff[q + 1]] for q in np.nonzero(col == c)[0]])
        assert seen.size == np.unique(seen).size
    order = np.concatenate([np.nonzero(col == c)[0] for c in range(ncol)])
    o = make_oracle(disc, A)
    x = rhs_vector(disc.total_dofs)
    y = np.zeros_like(x)
    pc.apply(x, y)
    assert relerr(y, o.applyMultiplicative(x, order)) < TOL
