"""Import-name shim: put ``aind-exaspim-image-compression_b200/`` ahead of
site-packages on ``sys.path`` and the reference's ``from bm4d import bm4d``
(machine_learning/data_handling.py:12, evaluate.py:11) binds to the B200 path."""
from b4d.api import BM4DProfile, BM4DStages, bm4d  # noqa: F401

__version__ = "4.2.5+b4d.b200"
__all__ = ["bm4d", "BM4DProfile", "BM4DStages"]
