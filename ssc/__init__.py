"""Drop-in names: ``-pc_python_type ssc.PatchPC`` / ``ssc.SSC`` resolve here (cf. ssc/__init__.py:1-2 of the reference)."""
from ssc_b200.ssc_pc import SSC      # noqa: F401
from ssc_b200.patch import PatchPC   # noqa: F401
from ssc_b200.relaxation import *   # noqa: F401,F403
