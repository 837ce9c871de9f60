"""rayforce_b200 — B200 (sm_100a) implementation of RayforceDB's vectorised columnar execution hot path.

Layers:
  include/rfb200.h + rayforce_b200/csrc/*.cu   C ABI + hand-written CUDA kernels (the product)
  rayforce_b200/capi.py                        ctypes declarations of that ABI
  rayforce_b200/device.py                      thin Python host helper: contexts, device columns, calls
"""
from . import capi  # noqa: F401
from .device import ColumnFile, Context, MultiGpu, RfbError  # noqa: F401

__all__ = ["capi", "ColumnFile", "Context", "MultiGpu", "RfbError"]
