"""The C-ABI library loads, exports every symbol include/ssc_b200.h declares, and refuses to compute without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ssc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(SSC[A-Za-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from ssc_b200 import _lib
    names = declared_functions()
    assert len(names) >= 40
    for name in names:
        assert hasattr(_lib.lib, name), "libssc_b200.so does not export %s" % name
    # and the Python binding knows every one of them
    assert set(names) <= set(_lib.EXPORTS)


def test_no_torch_types_and_no_oracle_in_product():
    """The boundary is plain C, and nothing under ssc_b200/ reaches into oracle/."""
    hdr = open(os.path.join(ROOT, "include", "ssc_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "cuda" not in code.lower()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ssc_b200")):
        for f in files:
            if f.endswith((".py", ".pyx", ".cu", ".cuh", ".inc", ".cpp", ".h")):      # .inc: the engine's functions by theme; .pyx: the Cython shim
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "ssc_oracle" not in src and "oracle/" not in src, f
                assert "import torch" not in src, f
    scanned = [f for _, _, files in os.walk(os.path.join(ROOT, "ssc_b200")) for f in files]
    assert any(f.endswith(".inc") for f in scanned) and any(f.endswith(".pyx") for f in scanned)


def test_host_only_handle_builds_patches_but_never_computes():
    from ssc_b200 import PC, SSCError, fem, plex
    from ssc_b200.patch import setup_patch_pc
    form, bcs = fem.poisson_problem(plex.RectangleMesh(4, 4, 1, 1), 2, "scalar")
    pc = PC(device=-1)
    setup_patch_pc(pc, form, bcs)
    off, gtol = pc.getPatches()
    assert off[-1] == gtol.size and pc.size("npatch") == 25
    with pytest.raises(SSCError) as e:
        pc.setUp()
    assert e.value.ierr == 97 and "no CPU fallback" in str(e.value)
    x = np.zeros(pc.size("nowned"))
    with pytest.raises(SSCError):
        pc.apply(x, x.copy())


def test_device_count_and_create_fail_loudly_without_gpu():
    import ssc_b200
    if ssc_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ssc_b200.SSCError) as e:
        ssc_b200.PC(device=0)
    assert "no CPU fallback" in str(e.value)


def test_option_errors_match_reference_messages():
    from ssc_b200 import PC, SSCError
    pc = PC(device=-1)
    pc.setOption("pc_patch_multiplicative", "false")
    with pytest.raises(SSCError) as e:
        pc.setOption("pc_patch_vanka_space", 1)
    assert "renamed to -pc_patch_exclude_subspace" in e.value.message                 # ssc/libssc.c:1591
    with pytest.raises(SSCError) as e:
        pc.setOption("pc_patch_construction_type", "blob")
    assert "Unknown patch construction type" in e.value.message                       # ssc/libssc.c:1600
    pc.setOption("pc_patch_construction_dim", 0)
    with pytest.raises(SSCError) as e:
        pc.setOption("pc_patch_construction_codim", 1)
    assert "Can only set one of dimension or co-dimension" in e.value.message         # ssc/libssc.c:1577
    with pytest.raises(SSCError):
        pc.setOption("no_such_option", 1)
    with pytest.raises(SSCError):
        pc.setOption("sub_ksp_type", "gmres")
    for ok in ("dense", "seqdense", "aij"):
        pc.setOption("pc_patch_sub_mat_type", ok)
    with pytest.raises(SSCError):
        pc.applyTranspose(None, None)                                                 # ssc/libssc.c:1715
    text = pc.view(viewer=__import__("ssc_b200.petsc_lite", fromlist=["Viewer"]).Viewer())
    assert "Schwarz type: additive" in text and "Patch construction operator: star" in text and "KSP not yet set" in text


def test_extension_options_are_range_checked_and_listed():
    """Every ssc_* extension option the library parses is in PATCH_OPTIONS (so setFromOptions forwards it), and the
    numeric ones reject values outside their range."""
    import re
    from ssc_b200 import PC, SSCError
    from ssc_b200.pcpatch import PATCH_OPTIONS
    csrc = os.path.join(ROOT, "ssc_b200", "csrc")
    src = "".join(open(os.path.join(csrc, f)).read() for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".inc")))
    parsed = set(re.findall(r'name == "(ssc_[a-z_0-9]+)"', src))
    assert parsed and parsed <= set(PATCH_OPTIONS), sorted(parsed - set(PATCH_OPTIONS))
    pc = PC(device=-1)
    for name, bad in (("ssc_factor_variant", 9), ("ssc_apply_variant", 2), ("ssc_refine_steps", 0), ("ssc_sub_solve", "gmres")):
        with pytest.raises(SSCError):
            pc.setOption(name, bad)
    for name, ok in (("ssc_factor_variant", 2), ("ssc_refine_steps", 2), ("ssc_host_pipeline_taper", "false"), ("ssc_sub_solve", "auto")):
        pc.setOption(name, ok)


def test_setup_without_inputs_is_an_error():
    from ssc_b200 import PC, SSCError
    pc = PC(device=-1)
    with pytest.raises(SSCError) as e:
        pc.buildPatches()
    assert "DM not yet set on patch PC" in e.value.message                             # ssc/libssc.c:475


def test_set_patches_validates_input():
    from ssc_b200 import PC, SSCError
    pc = PC(device=-1)
    pc.setPatches([0, 2, 3], [0, 1, 1], 4)                       # fine: dof 1 in two different patches
    for off, g in (([0, 2, 1], [0, 1]), ([1, 2], [0, 1]), ([0, 2], [3, 3]), ([0, 1], [-1]), ([0, 1], [9])):
        with pytest.raises(SSCError) as e:
            pc.setPatches(off, g, 4)
        assert e.value.ierr == 63
