"""Host side of the multigrid coarse solver (ssc_b200/csrc/coarse_amg.inc: strength graph, aggregation, smoothed prolongator,
Galerkin products), checked without a GPU: tools/micro/amg_host_check.cpp includes that file, builds the Q1 operator of an
m^3 mesh with Dirichlet rows, and runs CG preconditioned by the same V(1,1) cycle the device runs."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("amg") / "amg_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tools", "micro", "amg_host_check.cpp")], check=True)
    return exe


@pytest.mark.parametrize("m,levels", [(6, 1), (8, 2), (16, 2), (24, 2), (32, 3)])
def test_hierarchy_coarsens_and_the_cycle_is_h_independent(checker, m, levels):
    out = subprocess.run([checker, str(m)], check=True, capture_output=True, text=True, timeout=300).stdout
    sizes = [int(x) for x in re.findall(r"level \d+: n (\d+)", out)]
    its = int(re.search(r"PCG iterations to 1e-10: (\d+)", out).group(1))
    assert sizes[0] == (m + 1) ** 3 and len(sizes) == levels
    assert all(b * 8 < a for a, b in zip(sizes, sizes[1:]))          # aggregates of a 27-point stencil: ~27 rows each
    assert sizes[-1] <= 512                                          # the last level is inverted densely
    assert its <= (2 if levels == 1 else 16)                         # 12-13 from 16^3 to 96^3 cells


def test_block_aggregates_for_a_vector_valued_operator(checker):
    """Q1 elasticity on a clamped cube, dof = 3 * node + component: node aggregates with the three translations in the tentative
    prolongator (-lo_amg_block_size 3) keep the CG count flat; aggregates of the scalar matrix graph do not."""
    its = {}
    for m in (12, 16):
        for bs in (3, 1):
            out = subprocess.run([checker, str(m), "elasticity", str(bs)], check=True, capture_output=True, text=True, timeout=600).stdout
            its[m, bs] = int(re.search(r"PCG iterations to 1e-10: (\d+)", out).group(1))
    assert its[12, 3] <= 26 and its[16, 3] <= 28
    assert its[12, 3] < its[12, 1] and its[16, 3] < its[16, 1]


def test_rigid_body_modes_as_near_nullspace(checker):
    """The general form of the tentative prolongator: the given near-nullspace vectors (here the six rigid-body modes of
    elasticity) orthonormalised aggregate by aggregate, six coarse dofs per aggregate -- fewer iterations than with the
    three translations alone, and still flat in h."""
    its = {}
    for m in (12, 16):
        for extra in ((), ("rbm",)):
            out = subprocess.run([checker, str(m), "elasticity", "3", *extra], check=True, capture_output=True, text=True, timeout=600).stdout
            its[m, bool(extra)] = int(re.search(r"PCG iterations to 1e-10: (\d+)", out).group(1))
    assert its[12, True] < its[12, False] and its[16, True] < its[16, False]
    assert its[16, True] <= its[12, True] + 3
