"""tf_mnf_b200: B200-native (sm_100a) implementation of the tf-mnf MNF-layer hot path.

`tf_mnf_b200.layers`, `.flows`, `.models` mirror the reference's `tf_mnf.*` classes; all
arithmetic runs in libmnf_b200.so (hand-written CUDA, C ABI in include/mnf_b200.h, bound
with ctypes).  There is no CPU fallback: importing works anywhere, but creating a device
context (and therefore any layer call) raises without a CUDA device or a built library."""
from __future__ import annotations

from . import _lib, flows, layers, models  # noqa: F401
from .init import set_seed  # noqa: F401
from .runtime import Context, DeviceArray, as_device, default_context  # noqa: F401

__version__ = "0.1.0"
