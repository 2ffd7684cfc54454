"""Builds libssc_b200.so (CUDA engine + host patch builder) and libssc_synth.so (synthetic operators) in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssc_b200.so")
SYNTH = os.path.join(HERE, "libssc_synth.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "patch_builder.cpp")]
    # every header / include file of the translation unit (listed by directory, so a new .inc cannot be forgotten)
    deps = srcs + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".inc", ".cuh", ".h"))) + \
        [os.path.join(HERE, "..", "include", "ssc_b200.h")]
    if force or _stale(LIB, deps):
        cmd = [NVCC] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-O3,-pthread",
                               "-ccbin", "/usr/bin/g++", "-o", LIB] + srcs + ["-ldl", "-lpthread"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    ssrc = os.path.join(CSRC, "synth.cpp")
    if os.path.exists(ssrc) and (force or _stale(SYNTH, [ssrc])):
        subprocess.check_call(["/usr/bin/g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-shared", "-fPIC",
                               "-o", SYNTH, ssrc])
    build_cython(force)
    return LIB


def build_cython(force=False):
    """ssc_b200/PatchPC.pyx -> ssc_b200/PatchPC.<abi>.so, linked against libssc_b200.so through $ORIGIN."""
    import sysconfig
    import numpy
    pyx = os.path.join(HERE, "PatchPC.pyx")
    ext = os.path.join(HERE, "PatchPC" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not (force or _stale(ext, [pyx, os.path.join(HERE, "..", "include", "ssc_b200.h"), LIB])):
        return ext
    csrc = os.path.join(HERE, "csrc", "PatchPC_cython.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", csrc])
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-shared", "-fPIC", "-fno-strict-aliasing", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                           "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(), "-I", os.path.join(HERE, "..", "include"),
                           csrc, "-o", ext, "-L", HERE, "-l:libssc_b200.so", "-Wl,-rpath,$ORIGIN"])
    return ext


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
