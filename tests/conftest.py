import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _ensure_built():
    # Build the product library and the oracle if they are missing or stale (no-op when up to date).  At conftest import
    # time, i.e. before the test modules are collected: they import ssc_b200, which refuses to load without its library;
    # build.py is therefore loaded by path, not through the package.
    import importlib.util
    spec = importlib.util.spec_from_file_location("ssc_b200_build", os.path.join(ROOT, "ssc_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    from oracle import oracle
    oracle.build()


_ensure_built()


def meshes():
    """The three meshes of the reference's tests (tests/test_jacobi_equivalent.py:11-15)."""
    from ssc_b200 import plex
    return {"Interval": plex.IntervalMesh(10, 5.0), "Rectangle": plex.RectangleMesh(10, 20, 2, 3),
            "Box": plex.BoxMesh(5, 3, 5, 1, 2, 3)}


@pytest.fixture(params=["Interval", "Rectangle", "Box"])
def mesh(request):
    return meshes()[request.param]


@pytest.fixture(params=["scalar", "vector", "tensor", "mixed"])
def problem_type(request):
    return request.param


def rhs_vector(n, seed=20260101):
    """i.i.d. uniform(-1, 1), numpy default_rng(20260101)."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def delaunay_mesh(dim, npts, seed=3):
    """Unstructured simplicial mesh (scipy Delaunay of random points): vertex valences vary, so the vertex-star patches
    come in many sizes (the ragged case of SURVEY.md 8(d) C4)."""
    from scipy.spatial import Delaunay
    from ssc_b200 import plex
    pts = np.random.default_rng(seed).uniform(0, 1, (npts, dim))
    tri = Delaunay(pts)
    return plex.simplex_plex(dim, tri.simplices.astype(np.int64), pts)
