"""The product's host patch builder (csrc/patch_builder.cpp, through the C ABI on a host-only handle) against the
oracle's restatement of ssc/libssc.c:452-900: cells, gtol and the per-cell dofs table, bit-exact."""
import numpy as np
import pytest

from conftest import delaunay_mesh, meshes
from helpers import OPT_NAMES, make_oracle
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc


def both(mesh, p, kind, threads=0, **opts):
    form, bcs = fem.poisson_problem(mesh, p, kind)
    disc = fem.discretisation_from_form(form, bcs)
    o = make_oracle(disc, **opts)
    pc = PC(device=-1)
    setup_patch_pc(pc, form, bcs)
    pc.setOption("ssc_keep_dofs_table", True)
    if threads:
        pc.setOption("ssc_builder_threads", threads)
    for k, v in opts.items():
        pc.setOption(OPT_NAMES[k], ["star", "vanka", "user"][v] if k == "construct_type" else v)
    return o, pc


def assert_same(o, pc):
    for a, b in zip(o.patches(), pc.getPatches()):
        assert np.array_equal(a, b)
    for a, b in zip(o.cells(), pc.getCells()):
        assert np.array_equal(a, b)
    assert np.array_equal(o.dofs(), pc.getDofs())


@pytest.mark.parametrize("p", [1, 2, 3])
def test_star_patches_reference_meshes(mesh, problem_type, p):
    if p == 3 and mesh.dim == 3 and problem_type in ("tensor", "mixed"):
        pytest.skip("slow on the oracle side")
    assert_same(*both(mesh, p, problem_type))


@pytest.mark.parametrize("opts", [dict(construct_type=1), dict(construct_type=1, vankadim=0), dict(dim=1), dict(codim=0),
                                  dict(codim=1), dict(construct_type=1, exclude_subspace=1)])
@pytest.mark.parametrize("name", ["Rectangle", "Box"])
def test_other_constructions(name, opts):
    mesh = meshes()[name]
    assert_same(*both(mesh, 2, "mixed", **opts))


@pytest.mark.parametrize("shape,p", [((5, 4), 3), ((3, 4, 2), 2), ((3, 3, 3), 4), ((7,), 5)])
def test_tensor_meshes(shape, p):
    assert_same(*both(plex.TensorPlex(shape), p, "vector"))


def test_threaded_builder_is_deterministic():
    """> 4096 patches switches the builder to worker threads; chunks must concatenate to the serial answer."""
    mesh = plex.TensorPlex((20, 20, 12))
    o, pc = both(mesh, 2, "scalar", threads=5)
    assert pc.size("npatch") if pc.getPatches() else True
    assert_same(o, pc)
    _, pc1 = both(mesh, 2, "scalar", threads=1)
    assert all(np.array_equal(a, b) for a, b in zip(pc.getPatches(), pc1.getPatches()))


def test_ghost_vertices_get_no_patch():
    """Only owned vertices get patches (ssc/libssc.c:500-521)."""
    mesh = plex.TensorPlex((3, 3, 4))
    mesh.mark_ghost_vertices(2, 1, 3)
    V = fem.FunctionSpace(mesh, 2)
    disc = fem.Discretisation([V], owned_nodes=[V.owned_node_count])
    o = make_oracle(disc)
    pc = PC(device=-1)
    setup_patch_pc(pc, disc, None)
    off, gtol = pc.getPatches()
    assert all(np.array_equal(a, b) for a, b in zip(o.patches(), (off, gtol)))
    vs, ve = mesh.depth_stratum(0)
    sizes = np.diff(off)
    assert np.all(sizes[mesh.ghost[vs:ve] == 1] == 0) and np.any(sizes > 0)


def test_user_patches_through_the_python_hook():
    """-pc_patch_construction_type python + -pc_patch_construction_python_type (ssc/patch.py:169-181, ssc/libssc.c:1594-1598)."""
    from ssc_b200 import petsc_lite as pl
    mesh = plex.RectangleMesh(4, 3, 1, 1)
    form, bcs = fem.poisson_problem(mesh, 2, "scalar")
    disc = fem.discretisation_from_form(form, bcs)
    star = make_oracle(disc)
    pl.Options.clear()
    pl.Options("t_").setValue("pc_patch_construction_type", "python")
    pl.Options("t_").setValue("pc_patch_construction_python_type", "user_patch_example.vertex_stars")
    pc = PC(device=-1)
    pc.setOptionsPrefix("t_")
    setup_patch_pc(pc, form, bcs)
    pc.setFromOptions()
    pl.Options.clear()
    assert all(np.array_equal(a, b) for a, b in zip(star.patches(), pc.getPatches()))


def test_print_patches(capfd):
    """-pc_patch_print_patches (ssc/libssc.c:674-730): owned = patch dofs + their Dirichlet dofs, artificial = seen \\ owned."""
    mesh = plex.RectangleMesh(3, 3, 1, 1)
    form, bcs = fem.poisson_problem(mesh, 2, "scalar")
    disc = fem.discretisation_from_form(form, bcs)
    pc = PC(device=-1)
    setup_patch_pc(pc, form, bcs)
    pc.setOption("pc_patch_print_patches", True)
    off, gtol = pc.getPatches()
    text = pc.getPrintout()
    assert capfd.readouterr().out == text
    vs, ve = mesh.depth_stratum(0)
    blocks = [b for b in text.split("\n\n") if b.strip()]
    assert len(blocks) == ve - vs
    for i, b in enumerate(blocks):
        lines = b.strip().split("\n")
        assert lines[0] == "Patch %d: owned dofs:" % (vs + i) and lines[2].endswith("seen dofs:")
        owned = set(map(int, lines[1].split()))
        seen = set(map(int, lines[3].split()))
        gbc = set(map(int, lines[5].split())) if lines[5].strip() else set()
        art = set(map(int, lines[7].split())) if len(lines) > 7 and lines[7].strip() else set()
        assert art == seen - owned and gbc == seen & set(disc.ghost_bc_nodes.tolist())
        assert set(gtol[off[i]:off[i + 1]].tolist()) == owned - gbc


@pytest.mark.parametrize("dim,npts,p,kind", [(2, 150, 3, "scalar"), (2, 90, 2, "mixed"), (3, 120, 2, "vector"), (3, 60, 3, "scalar")])
def test_unstructured_delaunay_meshes(dim, npts, p, kind):
    """Ragged patches: on a Delaunay mesh the vertex valence varies, so do the patch sizes (10-20 distinct ones)."""
    o, pc = both(delaunay_mesh(dim, npts), p, kind)
    assert_same(o, pc)
    sizes = np.diff(pc.getPatches()[0])
    assert np.unique(sizes[sizes > 0]).size >= 6


def test_unstructured_vanka_and_edge_patches():
    mesh = delaunay_mesh(2, 80, seed=5)
    for opts in (dict(construct_type=1), dict(dim=1), dict(codim=0)):
        assert_same(*both(mesh, 2, "mixed", **opts))
