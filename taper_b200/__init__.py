"""taper_b200 — B200-native (sm_100a) backend for taper's tape-evaluation hot path.

The product is ``libtaper_b200.so`` (hand-written CUDA kernels + C ABI + C++ host layer mirroring
taper's Tensor/Tape/nn::Module API).  This Python package is only the ctypes binding used by the
tests, ``bench.py`` and ``__graft_entry__.py``.  There is no CPU fallback: importing the package
without the built library raises ImportError.
"""
from . import capi  # noqa: F401  (loads the shared library; raises if it is missing)
from .capi import Ctx, Buf, ConvDesc, PoolDesc, TaperError, lib  # noqa: F401

__all__ = ["capi", "Ctx", "Buf", "ConvDesc", "PoolDesc", "TaperError", "lib"]
