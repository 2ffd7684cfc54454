"""The petsc4py-facing adapters (Plex.from_dmplex, _section_arrays, PC.setDM / setPatchCellNumbering /
setPatchDiscretisationInfo with live-PETSc-shaped arguments) cannot meet a real petsc4py in this image, so they are
driven by duck-typed stand-ins that expose exactly the petsc4py methods the adapters call (DMPlex: getChart, getCone,
getDimension, getDepthStratum, hasLabel, getLabelValue; Section: getChart, getDof, getOffset; DM: getDefaultSection).
The patches built through that route must equal the ones built from the arrays directly."""
import numpy as np
import pytest

from conftest import delaunay_mesh
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc


class FakeDMPlex(object):
    def __init__(self, p):
        self.p = p

    def getChart(self):
        return 0, int(self.p.pEnd)

    def getCone(self, q):
        return [int(c) for c in self.p.cone(q)]

    def getDimension(self):
        return int(self.p.dim)

    def getDepthStratum(self, d):
        s, e = self.p.depth_stratum(d)
        return int(s), int(e)

    def hasLabel(self, name):
        return name == "pyop2_ghost" and self.p.ghost is not None and bool(np.any(self.p.ghost))

    def getLabelValue(self, name, q):
        return 1 if self.p.ghost[q] else -1


class FakeSection(object):
    def __init__(self, dof, off):
        self.dof, self.off = np.asarray(dof), np.asarray(off)

    def getChart(self):
        return 0, int(self.dof.size)

    def getDof(self, q):
        return int(self.dof[q])

    def getOffset(self, q):
        return int(self.off[q])


class FakeDM(object):
    def __init__(self, sec):
        self.sec = sec

    def getDefaultSection(self):
        return self.sec


@pytest.mark.parametrize("mesh", [plex.RectangleMesh(5, 4, 1, 1), plex.TensorPlex((3, 3, 2)), delaunay_mesh(3, 60)], ids=["tri", "hex", "delaunay-tet"])
def test_live_petsc_shaped_inputs_give_the_same_patches(mesh):
    form, bcs = fem.poisson_problem(mesh, 2, "mixed" if mesh.cell_type == "simplex" else "vector")
    disc = fem.discretisation_from_form(form, bcs)
    ref = PC(device=-1)
    setup_patch_pc(ref, form, bcs)
    ref.setOption("ssc_keep_dofs_table", True)
    pc = PC(device=-1)
    pc.setOption("ssc_keep_dofs_table", True)
    flat = plex.Plex.from_dmplex(FakeDMPlex(mesh))
    for name in ("cone_off", "cones"):
        assert np.array_equal(getattr(flat, name), getattr(mesh, name))
    assert np.array_equal(flat.ghost.astype(bool), np.asarray(mesh.ghost).astype(bool))
    pc.setDM(FakeDMPlex(mesh))                                                   # goes through Plex.from_dmplex
    pc.setPatchCellNumbering(FakeSection(disc.cell_numbering_dof, disc.cell_numbering_off))                        # goes through _section_arrays
    dms = [FakeDM(FakeSection(V.sec_dof, V.sec_off)) for V in form.spaces]
    pc.setPatchDiscretisationInfo(dms, disc.bs, [V.cell_node_list for V in form.spaces], disc.offsets, disc.ghost_bc_nodes, disc.global_bc_nodes)
    for a, b in zip(ref.getPatches(), pc.getPatches()):
        assert np.array_equal(a, b)
    for a, b in zip(ref.getCells(), pc.getCells()):
        assert np.array_equal(a, b)
    assert np.array_equal(ref.getDofs(), pc.getDofs())
