"""fwb200 — ctypes binding over libfwb200.so (C ABI declared in include/fwb200.h).

PyTorch is used only for device memory and streams; every op here launches our own sm_100a kernels.
There is NO fallback: if the shared library is missing or the device is not sm_100, the ops raise.
"""
from . import _lib  # noqa: F401
from ._lib import lib, FwbError, library_path, abi_symbols  # noqa: F401
from .ops import *  # noqa: F401,F403
