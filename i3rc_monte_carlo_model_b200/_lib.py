"""Loader of the product's C-ABI shared library (csrc/ -> libi3rc_b200.so, built in-tree).

There is no fallback: a missing library is an ImportError-grade failure, and every compute entry point
of the library itself fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

from ._abi import Backend

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libi3rc_b200.so")
_backend = None


def backend() -> Backend:
    global _backend
    if _backend is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The product has no CPU fallback.")
        _backend = Backend(C.CDLL(LIB_PATH), "i3rc_", "cuda")
    return _backend
