"""The C ABI from plain C (the language of ssc/libssc.c): examples/c_api_demo.c compiles as C99 against include/ssc_b200.h,
links against the shared library, and verifies PCApply against a closed form on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import ssc_b200  # noqa: F401  (makes sure the library is built)
    exe = str(tmp_path / "c_api_demo")
    libdir = os.path.join(ROOT, "ssc_b200")
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_api_demo.c"),
           "-L" + libdir, "-lssc_b200", "-lm", "-Wl,-rpath," + libdir, "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_example_compiles_as_c99_and_fails_loudly_without_gpu(tmp_path):
    import ssc_b200
    if ssc_b200.device_count() > 0:
        pytest.skip("a GPU is visible: covered by the gpu test")
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 3 and "no CUDA device" in out.stderr


@pytest.mark.gpu
def test_c_example_matches_closed_form_on_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "relative error" in out.stdout
