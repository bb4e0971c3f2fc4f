"""mauve-py_b200: B200 (sm_100a) implementation of progressiveMauve's anchoring hot path.

Scope: sorted-mer-list construction, seed-match enumeration, match
extension, the inter-anchor gapped DP and the pairwise homology HMM -- behind the libMems /
libMUSCLE interfaces of that path (`mauve_py_b200.libmems`) and a C ABI (include/mauve_cuda.h,
libmauve_cuda.so).  There is no CPU fallback: without the built library the import of a compute
entry fails, without an sm_100 GPU every compute call raises McuError.
"""
from . import libmems  # noqa: F401
from .libmems import *  # noqa: F401,F403
from ._capi import LIB_PATH, SYMBOLS, McuError, lib  # noqa: F401
from .buildindex import buildIndex  # noqa: F401

__version__ = "0.1.0"
