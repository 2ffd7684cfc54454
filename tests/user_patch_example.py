"""A user patch constructor in the shape ssc/relaxation.py uses: callable(pc) -> (list of point sets, iteration order)."""
import numpy as np


def vertex_stars(pc, *args, **kwargs):
    mesh = pc.plex
    vs, ve = mesh.depth_stratum(0)
    patches = [np.array(mesh.star(v), dtype=np.int64) for v in range(vs, ve)]
    return patches, np.arange(len(patches), dtype=np.int64)
