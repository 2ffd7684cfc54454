"""ssc_b200: B200-native patch subspace-correction preconditioner (the hot path of wence-/ssc).

Importing the package loads the CUDA shared library; there is no CPU fallback.
"""
from ._lib import SSCError, device_count, LIB_PATH          # noqa: F401
from .pcpatch import PC, DeviceVec, PinnedArray            # noqa: F401
from .patch import PatchPC, setup_patch_pc                 # noqa: F401
from .lo import P1PC                                        # noqa: F401
from .ssc_pc import SSC                                    # noqa: F401
