"""ssc_b200: B200-native patch subspace-correction preconditioner (the hot path of wence-/ssc).

The product names (PC, PatchPC, SSC, ...) load the CUDA shared library on first use; there is no CPU fallback: without
the library they raise ImportError.  The synthetic problem generator (``fem``, ``plex``, ``synth``, ``partition``) is
plain numpy scaffolding and imports without it -- bench.py's reference arm and the oracle tests use it alone.
"""
import importlib

_LAZY = {"SSCError": "._lib", "device_count": "._lib", "LIB_PATH": "._lib",
         "PC": ".pcpatch", "DeviceVec": ".pcpatch", "PinnedArray": ".pcpatch",
         "PatchPC": ".patch", "setup_patch_pc": ".patch", "P1PC": ".lo", "SSC": ".ssc_pc"}


def __getattr__(name):
    if name in _LAZY:
        value = getattr(importlib.import_module(_LAZY[name], __name__), name)
        globals()[name] = value
        return value
    try:
        return importlib.import_module("." + name, __name__)
    except ImportError:
        raise AttributeError("module %r has no attribute %r" % (__name__, name))


def __dir__():
    return sorted(list(globals()) + list(_LAZY))
