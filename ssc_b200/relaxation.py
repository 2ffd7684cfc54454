"""User patch constructors, SURVEY.md 8(f) N3: mirror of ssc/relaxation.py.

Each class is a callable ``fun(pc) -> (patches, iteration_set)`` selected with
``-pc_patch_construction_type python -pc_patch_construction_python_type ssc.PlaneSmoother`` (ssc/patch.py:169-181),
where ``patches`` is a list of mesh-point arrays and ``iteration_set`` the order in which the patch PC visits them
(ssc/libssc.c:1328-1351).  In the additive PC the order does not change the sum; a patch listed k times is added k
times.  The DMPlex queries of the reference become array queries on :class:`ssc_b200.plex.Plex`.
"""
import numpy as np

__all__ = ("PlaneSmoother", "OrderedVanka", "OrderedStar")


def _options(pc):
    from . import petsc_lite
    db = getattr(pc, "options_db", None) or petsc_lite.Options
    return db(pc.getOptionsPrefix())


def _owned(dm, points):
    """select_entity(..., exclude="pyop2_ghost"), ssc/relaxation.py:8-19."""
    return [p for p in points if not dm.ghost[p]]


def _coords(dm, p):
    """Mean of the vertex coordinates in the closure of p (ssc/relaxation.py:23-29)."""
    vs, ve = dm.depth_stratum(0)
    verts = [q - vs for q in dm.closure(p) if vs <= q < ve]
    return dm.coords[verts].mean(axis=0)


class PlaneSmoother(object):
    """Slabs of mesh points along an axis; ``-pc_patch_construction_ps_sweeps "0+4:1-3"`` (ssc/relaxation.py:22-79)."""

    def sort_entities(self, dm, axis, dir, ndiv):
        entities = [(p, _coords(dm, p)) for p in _owned(dm, range(dm.pEnd))]
        minx = min(e[1][axis] for e in entities)
        maxx = max(e[1][axis] for e in entities)

        def keyfunc(z):
            c = tuple(z[1])
            return (c[axis],) + tuple(c[:axis] + c[axis + 1:])
        s = sorted(entities, key=keyfunc, reverse=(dir == -1))
        divisions = np.linspace(minx, maxx, ndiv + 1)
        pts = [e[0] for e in s]
        coords = [e[1][axis] for e in s]
        indices = np.searchsorted(coords[::dir], divisions)
        out = [pts[indices[k]:indices[k + 1]] for k in range(ndiv)]
        out.append(pts[indices[-1]:])
        return out

    def __call__(self, pc):
        dm = pc.getDM()
        sweeps = _options(pc).getString("pc_patch_construction_ps_sweeps", None)
        if sweeps is None:
            raise ValueError("Must set %spc_patch_construction_ps_sweeps" % pc.getOptionsPrefix())
        patches = []
        for sweep in sweeps.split(":"):
            axis = int(sweep[0])
            dir = {"+": +1, "-": -1}[sweep[1]]
            ndiv = int(sweep[2:])
            for patch in self.sort_entities(dm, axis, dir, ndiv):
                patches.append(np.asarray(patch, dtype=np.int64))
        return patches, np.arange(len(patches), dtype=np.int64)


class OrderedRelaxation(object):
    """ssc/relaxation.py:82-159."""
    name = None

    def callback(self, dm, entity, nclosures):
        raise NotImplementedError

    def set_options(self, dm, opts, name):
        pass

    @staticmethod
    def get_entities(opts, name, dm):
        codim = opts.getInt("pc_patch_construction_%s_codim" % name, None)
        if codim is None:
            dim = opts.getInt("pc_patch_construction_%s_dim" % name, 0)
            return range(*dm.depth_stratum(dim))
        return range(*dm.height_stratum(codim))

    def __call__(self, pc):
        dm = pc.getDM()
        opts = _options(pc)
        name = self.name
        self.set_options(dm, opts, name)
        entities = _owned(dm, self.get_entities(opts, name, dm))
        nclosure = opts.getInt("pc_patch_construction_%s_nclosures" % name, 1)
        patches = [np.asarray(self.callback(dm, e, nclosure), dtype=np.int64) for e in entities]
        sortorders = opts.getString("pc_patch_construction_%s_sort_order" % name, "None")
        if sortorders == "None":
            return patches, np.arange(len(patches), dtype=np.int64)
        coords = list(enumerate(_coords(dm, p) for p in entities))
        iterset = []
        for sortorder in sortorders.split("|"):
            sortdata = []
            for axis in sortorder.split(":"):
                sgn = {"+": 1, "-": -1}[axis[1]] if len(axis) > 1 else 1
                sortdata.append((int(axis[0]), sgn))
            iterset += [x[0] for x in sorted(coords, key=lambda z: tuple(s * z[1][a] for a, s in sortdata))]
        return patches, np.asarray(iterset, dtype=np.int64)


class OrderedVanka(OrderedRelaxation):
    """n-fold closure-of-star patches, optionally dropping one entity dimension (ssc/relaxation.py:162-199)."""
    name = "ovanka"

    def set_options(self, dm, opts, name):
        self.exclude_dim = opts.getInt("pc_patch_construction_%s_exclude_dim" % name, None)
        self.exclude_range = dm.depth_stratum(self.exclude_dim) if self.exclude_dim is not None else None

    def callback(self, dm, entity, nclosures):
        out = {entity}
        for _ in range(nclosures):
            new = set()
            for e in out:
                for p in dm.star(e):
                    for x in dm.closure(p):
                        if self.exclude_range is None or not (self.exclude_range[0] <= x < self.exclude_range[1]):
                            new.add(x)
            out |= new
        return sorted(out)


class OrderedStar(OrderedRelaxation):
    """Stars grown ``nclosures`` times (ssc/relaxation.py:202-219)."""
    name = "ostar"

    def callback(self, dm, entity, nclosures):
        out = {entity}
        candidates = {entity}
        while candidates:
            e = candidates.pop()
            star = dm.star(e)
            nclosures -= 1
            if nclosures > 0:
                candidates.update(star)
            out.update(star)
        return sorted(out)
