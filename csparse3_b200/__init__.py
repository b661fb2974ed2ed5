"""csparse3_b200 -- B200-native (sm_100a) drop-in for the numeric hot path of SanPen/CSparse3.

Public surface mirrors `import CSparse3` (src/CSparse3/__init__.py:1-4) for the CSC path:
CscMat, scipy_to_mat, Diag, Diags, pack_4_by_4 and the flat csc_* kernels, plus the batched LU object
(csparse3_b200.lu.LuSymbolic) and the device SpMV plan (csparse3_b200.spmv.SpmvPlan).
"""
from .csc import CscMat, Diag, Diags, pack_4_by_4, scipy_to_mat  # noqa: F401
from .csc_b200 import *  # noqa: F401,F403
from .lu import LuSymbolic  # noqa: F401
from .spmv import SpmvPlan  # noqa: F401

__version__ = "0.1.0"
