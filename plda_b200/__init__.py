"""plda_b200 -- B200-native PLDA / LDA hot path behind the reference's ``liblda`` API.

    from plda_b200 import PLDA, LDA        # or: from liblda import PLDA, LDA

Everything numeric runs in hand-written sm_100a CUDA kernels reached through the
C ABI of ``include/plda_b200.h`` (ctypes, ``plda_b200/_ffi.py``).  Importing the
package does not load the shared library; constructing ``PLDA()`` / ``LDA()`` does,
and fails loudly if it is missing or no B200 is present (no CPU fallback).
"""
from .lda import LDA
from .plda import PLDA

__all__ = ["PLDA", "LDA"]
__version__ = "0.1.0"
