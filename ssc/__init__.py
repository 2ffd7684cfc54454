"""Drop-in name: ``-pc_python_type ssc.PatchPC`` resolves here (cf. ssc/__init__.py:2 of the reference)."""
from ssc_b200.patch import PatchPC   # noqa: F401
