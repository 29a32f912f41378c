"""Device context and the DeviceArray handle the host classes pass around.

DeviceArray is deliberately thin: a device pointer, a shape and an owner.  It speaks DLPack
(`__dlpack__`) and `__cuda_array_interface__`, so TF (`tf.experimental.dlpack.from_dlpack`),
PyTorch (`torch.from_dlpack` / `torch.as_tensor`) or CuPy consume it zero-copy, and
`as_device()` accepts any of those producers zero-copy in the other direction.
NumPy inputs are copied host->device explicitly (what Keras does for the reference when
it is handed a NumPy batch, notebooks/lenet_mnist.py:84); compute never runs on the host.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Any, Sequence

import numpy as np

from . import _lib as L

_DT = {"f4": np.float32, "i1": np.int8, "u4": np.uint32}
_ITEMSIZE = {k: np.dtype(v).itemsize for k, v in _DT.items()}


class Context:
    """One (device, stream, workspace).  Not thread-safe: use one context per thread."""

    def __init__(self, device: int = 0, stream: int | None = None, _fork_of: "Context | None" = None, _rank: int = 0) -> None:
        self.lib = L.load()
        h = C.c_void_p()
        if _fork_of is not None:  # a lower-priority sibling (mnf_ctx_fork)
            L.check(self.lib.mnf_ctx_fork(_fork_of.handle, int(_rank), C.byref(h)))
        else:
            L.check(self.lib.mnf_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = device
        self.math = "tf32x3"  # the library's default (mnf_b200.h): tcgen05 with hi/lo operand splitting
        # Host-side free list of stream-ordered allocations, by size.  A block released by Python
        # and handed out again is only ever touched by LATER work of this context's stream, which
        # is exactly the guarantee of cudaFreeAsync + cudaMallocAsync on one stream -- without
        # the two library calls per temporary (a forward + KL step makes ~40 of them).
        self._capture_keep: list | None = None  # graph capture: owners of every buffer handed out
        self._pool: dict[int, list[int]] = {}
        self._pooled_bytes = 0
        self.pool_limit_bytes = 2 << 30

    # -- memory
    def empty(self, shape: Sequence[int], dtype: str = "f4") -> "DeviceArray":
        shape = tuple(int(s) for s in shape)
        nbytes = (math.prod(shape) * _ITEMSIZE[dtype] + 255) & ~255
        free = self._pool.get(nbytes)
        if free:
            ptr = free.pop()
            self._pooled_bytes -= nbytes
        else:
            if getattr(self, "_capturing", False):
                raise RuntimeError(f"buffer pool miss ({nbytes} bytes) while capturing a CUDA graph: the step "
                                   "allocated differently from its warm-up pass")
            p = C.c_void_p()
            L.check(self.lib.mnf_malloc(self.handle, nbytes, C.byref(p)))
            ptr = p.value
        owner = _Owned(self, ptr, nbytes)
        if self._capture_keep is not None:
            self._capture_keep.append(owner)
        return DeviceArray(self, ptr, shape, dtype, owner=owner)

    def trim(self) -> None:
        """Return the pooled blocks to the driver's stream-ordered pool."""
        for free in self._pool.values():
            for ptr in free:
                self.lib.mnf_free(self.handle, C.c_void_p(ptr))
        self._pool.clear()
        self._pooled_bytes = 0

    def zeros(self, shape: Sequence[int], dtype: str = "f4") -> "DeviceArray":
        a = self.empty(shape, dtype)
        L.check(self.lib.mnf_memset(self.handle, a.ptr, 0, a.nbytes))
        return a

    def to_device(self, arr: Any, dtype: str = "f4") -> "DeviceArray":
        h = np.ascontiguousarray(arr, dtype=_DT[dtype])
        a = self.empty(h.shape, dtype)
        L.check(self.lib.mnf_memcpy_h2d(self.handle, a.ptr, h.ctypes.data, h.nbytes))
        self.sync()  # h may be a temporary
        return a

    def sync(self) -> None:
        L.check(self.lib.mnf_ctx_sync(self.handle))

    def fork(self, rank: int = 0) -> "Context":
        """A sibling context on the same device with its own stream and workspace (e.g. to run
        kl_div(), which does not depend on x, concurrently with a forward pass).  Its stream runs below the
        main streams' priority, `rank` levels above the device's least (mnf_ctx_fork): side work fills in
        behind the main stream, and the side stream whose result the main stream needs first gets the
        highest rank."""
        c = Context(self.device, _fork_of=self, _rank=rank)
        if getattr(self, "math", None):
            c.set_math(self.math)
        if getattr(self, "frozen", False):
            c.freeze_weights(True)
        return c

    def freeze_weights(self, frozen: bool = True) -> None:
        """Inference with fixed parameters: keep the packed weight blocks of the tensor-core
        kernels across calls (mnf_ctx_freeze_weights).  Unfreeze before changing any parameter."""
        L.check(self.lib.mnf_ctx_freeze_weights(self.handle, int(bool(frozen))))
        self.frozen = bool(frozen)

    def wait(self, other: "Context") -> None:
        """Order this context's future work after everything enqueued so far on `other`."""
        L.check(self.lib.mnf_ctx_wait(self.handle, other.handle))

    def wait_stream(self, stream: int) -> None:
        """Order this context's future work after everything enqueued so far on a foreign stream
        (cudaStream_t handle; 1 = legacy default, 2 = per-thread default)."""
        L.check(self.lib.mnf_ctx_wait_stream(self.handle, C.c_void_p(int(stream))))

    def set_math(self, mode: str) -> None:
        """"fp32": FFMA kernels; "tf32x3": tcgen05 tensor cores with hi/lo operand splitting
        (fp32-level accuracy); "bf16": as tf32x3, but MNFDense contractions use one bf16 MMA per
        product (tolerance in include/mnf_b200.h).  Applies to the dense / conv forward contractions."""
        codes = {"fp32": 0, "tf32x3": 1, "bf16": 2}
        if mode not in codes:
            raise ValueError(f"math mode must be one of {sorted(codes)}, got {mode!r}")
        L.check(self.lib.mnf_ctx_set_math(self.handle, codes[mode]))
        self.math = mode

    def set_stream(self, stream: int) -> None:
        L.check(self.lib.mnf_ctx_set_stream(self.handle, C.c_void_p(stream)))

    @property
    def stream(self) -> int:
        p = C.c_void_p()
        L.check(self.lib.mnf_ctx_get_stream(self.handle, C.byref(p)))
        return p.value or 0

    @property
    def sm_count(self) -> int:
        n = C.c_int()
        L.check(self.lib.mnf_ctx_sm_count(self.handle, C.byref(n)))
        return n.value

    @property
    def launches(self) -> int:
        n = C.c_int64()
        L.check(self.lib.mnf_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.trim()
            self.lib.mnf_ctx_destroy(self.handle)
            self.handle = None


class _Owned:
    """Frees a stream-ordered allocation when the last view dies."""

    __slots__ = ("ctx", "ptr", "nbytes")

    def __init__(self, ctx: Context, ptr: int, nbytes: int = 0) -> None:
        self.ctx, self.ptr, self.nbytes = ctx, ptr, nbytes

    def __del__(self) -> None:
        try:
            ctx = self.ctx
            if ctx.handle:
                n = self.nbytes
                if n and ctx._pooled_bytes + n <= ctx.pool_limit_bytes:
                    ctx._pool.setdefault(n, []).append(self.ptr)
                    ctx._pooled_bytes += n
                else:
                    ctx.lib.mnf_free(ctx.handle, C.c_void_p(self.ptr))
        except Exception:  # interpreter shutdown
            pass


_tls = threading.local()

# Bumped whenever a parameter buffer may have changed (DeviceArray.assign / copy_from_host, the
# optimiser step): anything derived from parameters AHEAD of its use -- the z a layer draws one call
# early (MNFLayerBase.presample) -- records the version it was computed under and is discarded when
# the version has moved on.
_param_version = 0


def param_version() -> int:
    return _param_version


def bump_param_version() -> None:
    global _param_version
    _param_version += 1


def default_context(device: int | None = None) -> Context:
    """Per-thread, per-device default context (created on first use; fails without a GPU)."""
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    if device is None:
        device = getattr(_tls, "device", 0)
    if device not in ctxs:
        ctxs[device] = Context(device)
    return ctxs[device]


def set_default_device(device: int) -> None:
    _tls.device = device


class DeviceArray:
    __slots__ = ("ctx", "ptr", "shape", "dtype", "owner", "__weakref__")

    def __init__(self, ctx: Context, ptr: int, shape, dtype: str = "f4", owner: Any = None) -> None:
        self.ctx, self.ptr, self.shape, self.dtype, self.owner = ctx, int(ptr or 0), tuple(shape), dtype, owner

    # -- basic properties
    @property
    def size(self) -> int:
        return math.prod(self.shape)

    @property
    def itemsize(self) -> int:
        return _ITEMSIZE[self.dtype]

    @property
    def nbytes(self) -> int:
        return self.size * self.itemsize

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def __len__(self) -> int:
        return self.shape[0]

    def __repr__(self) -> str:
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, ptr=0x{self.ptr:x}, device={self.ctx.device})"

    # -- views
    def reshape(self, *shape) -> "DeviceArray":
        if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            i = shape.index(-1)
            rest = math.prod(int(s) for s in shape if s != -1)
            shape[i] = self.size // max(rest, 1)
        if math.prod(int(s) for s in shape) != self.size:
            raise ValueError(f"cannot reshape {self.shape} to {tuple(shape)}")
        return DeviceArray(self.ctx, self.ptr, shape, self.dtype, self.owner)

    def __getitem__(self, idx) -> "DeviceArray":
        """Leading-axis integer index or contiguous slice (a view)."""
        inner = math.prod(self.shape[1:]) * self.itemsize
        if isinstance(idx, (int, np.integer)):
            i = int(idx) % self.shape[0] if self.shape[0] else 0
            return DeviceArray(self.ctx, self.ptr + i * inner, self.shape[1:], self.dtype, self.owner)
        if isinstance(idx, slice):
            start, stop, step = idx.indices(self.shape[0])
            if step != 1:
                raise ValueError("only contiguous slices are views")
            n = max(stop - start, 0)
            return DeviceArray(self.ctx, self.ptr + start * inner, (n, *self.shape[1:]), self.dtype, self.owner)
        raise TypeError("DeviceArray supports leading-axis int / slice indexing only")

    # -- host transfer
    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, _DT[self.dtype])
        if self.nbytes:
            L.check(self.ctx.lib.mnf_memcpy_d2h(self.ctx.handle, out.ctypes.data, C.c_void_p(self.ptr), self.nbytes))
            self.ctx.sync()
        return out

    def copy_from_host(self, arr) -> "DeviceArray":
        h = np.ascontiguousarray(arr, dtype=_DT[self.dtype])
        if h.shape != self.shape:
            raise ValueError(f"shape mismatch: {h.shape} vs {self.shape}")
        L.check(self.ctx.lib.mnf_memcpy_h2d(self.ctx.handle, C.c_void_p(self.ptr), h.ctypes.data, h.nbytes))
        self.ctx.sync()
        bump_param_version()
        return self

    def copy(self) -> "DeviceArray":
        out = self.ctx.empty(self.shape, self.dtype)
        L.check(self.ctx.lib.mnf_memcpy_d2d(self.ctx.handle, C.c_void_p(out.ptr), C.c_void_p(self.ptr), self.nbytes))
        return out

    def assign(self, value) -> "DeviceArray":
        """tf.Variable.assign look-alike (host array or device array)."""
        if isinstance(value, DeviceArray):
            if value.shape != self.shape:
                raise ValueError(f"shape mismatch: {value.shape} vs {self.shape}")
            L.check(self.ctx.lib.mnf_memcpy_d2d(self.ctx.handle, C.c_void_p(self.ptr), C.c_void_p(value.ptr), self.nbytes))
            bump_param_version()
            return self
        return self.copy_from_host(value)

    # -- zero-copy export
    @property
    def __cuda_array_interface__(self) -> dict:
        typestr = {"f4": "<f4", "i1": "|i1", "u4": "<u4"}[self.dtype]
        return {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False), "version": 3,
                "strides": None, "stream": self.ctx.stream or None}

    def __dlpack_device__(self):
        return (2, self.ctx.device)  # kDLCUDA

    def __dlpack__(self, stream=None, **_):
        if self.dtype != "f4":
            raise TypeError("DLPack export is float32 only")
        self.ctx.sync()  # old-style capsules carry no stream: hand over finished data
        keep = _KEEP.pin(self)
        shape = (C.c_int64 * max(self.ndim, 1))(*self.shape)
        out = C.c_void_p()
        L.check(self.ctx.lib.mnf_dlpack_wrap(C.c_void_p(self.ptr), self.ndim, shape, self.ctx.device,
                                             _RELEASE, C.c_void_p(keep), C.byref(out)))
        return _capsule_new(out.value)


class _KeepAlive:
    """Keeps exported arrays alive until the DLPack consumer calls the deleter."""

    def __init__(self) -> None:
        self.live: dict[int, Any] = {}
        self.n = 0
        self.lock = threading.Lock()

    def pin(self, obj) -> int:
        with self.lock:
            self.n += 1
            self.live[self.n] = obj
            return self.n

    def release(self, key) -> None:
        with self.lock:
            self.live.pop(int(key or 0), None)


_KEEP = _KeepAlive()
_RELEASE = L.RELEASE_FN(lambda key: _KEEP.release(key))

_PyCapsule_New = C.pythonapi.PyCapsule_New
_PyCapsule_New.restype = C.py_object
_PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = C.c_void_p
_PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_PyCapsule_IsValid = C.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = C.c_int
_PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]


_IsValid_raw = C.PYFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p)(("PyCapsule_IsValid", C.pythonapi))
_GetPointer_raw = C.PYFUNCTYPE(C.c_void_p, C.c_void_p, C.c_char_p)(("PyCapsule_GetPointer", C.pythonapi))


@C.PYFUNCTYPE(None, C.c_void_p)
def _capsule_destructor(cap_ptr):
    # a consumer renames the capsule to "used_dltensor"; if it is still "dltensor" nobody took it
    if _IsValid_raw(cap_ptr, b"dltensor"):
        L.load().mnf_dlpack_delete(C.c_void_p(_GetPointer_raw(cap_ptr, b"dltensor")))


def _capsule_new(ptr: int):
    return _PyCapsule_New(ptr, b"dltensor", C.cast(_capsule_destructor, C.c_void_p))


def dlpack_view(capsule) -> L.TensorView:
    """Borrowed view of a "dltensor" capsule (the capsule keeps ownership)."""
    if not _PyCapsule_IsValid(capsule, b"dltensor"):
        raise ValueError("expected an unconsumed 'dltensor' PyCapsule")
    p = _PyCapsule_GetPointer(capsule, b"dltensor")
    tv = L.TensorView()
    L.check(L.load().mnf_dlpack_view(C.c_void_p(p), C.byref(tv)))
    return tv


def as_device(x: Any, ctx: Context | None = None) -> DeviceArray:
    """Zero-copy view of a device tensor from any framework, or an explicit H2D copy of a
    NumPy array.  float32 and C-contiguous only; anything else is an error (no silent casts
    on device tensors)."""
    if isinstance(x, DeviceArray):
        return x
    ctx = ctx or default_context()
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        if cai["typestr"] not in ("<f4", "=f4"):
            raise ValueError(f"expected float32 device tensor, got {cai['typestr']}")
        if cai.get("strides") is not None:
            exp, ok = 4, True
            for s, st in zip(reversed(cai["shape"]), reversed(cai["strides"])):
                ok &= s == 1 or st == exp
                exp *= s
            if not ok:
                raise ValueError("device tensor must be C-contiguous")
        # the producer's stream (interface v3): our non-blocking stream has no implicit ordering with
        # it, so wait for what it has enqueued; without the field, order against the legacy default stream
        ctx.wait_stream(cai.get("stream") or 1)
        return DeviceArray(ctx, cai["data"][0], cai["shape"], "f4", owner=x)
    if isinstance(x, np.ndarray) or isinstance(x, (list, tuple, float, int)):
        return ctx.to_device(np.asarray(x, dtype=np.float32))
    if hasattr(x, "__dlpack__"):
        dev = x.__dlpack_device__() if hasattr(x, "__dlpack_device__") else (2, ctx.device)
        if int(dev[0]) not in (2, 13):
            raise ValueError(f"tensor is on DLPack device type {int(dev[0])}, not CUDA: tf_mnf_b200 has no CPU path")
        try:  # DLPack protocol: the producer makes the data safe to use on the consumer's stream
            cap = x.__dlpack__(stream=ctx.stream)
        except TypeError:  # pre-protocol producer: no stream argument
            cap = x.__dlpack__()
            ctx.wait_stream(1)
        tv = dlpack_view(cap)
        if (tv.dtype_code, tv.dtype_bits) != (2, 32) or not tv.contiguous:
            raise ValueError("expected a C-contiguous float32 tensor")
        shape = tuple(tv.shape[i] for i in range(tv.ndim))
        return DeviceArray(ctx, tv.data, shape, "f4", owner=(cap, x))
    raise TypeError(f"cannot interpret {type(x).__name__} as a device tensor")


def from_dlpack_capsule(capsule, ctx: Context | None = None, producer_stream: int | None = 1) -> DeviceArray:
    """For TF 2.1-style `tf.experimental.dlpack.to_dlpack(t)` capsules.  They carry no stream: the
    context's stream is ordered behind `producer_stream` (default 1 = the legacy default stream; pass
    the framework's compute stream, or None when the caller has synchronised / adopted that stream
    with Context.set_stream)."""
    ctx = ctx or default_context()
    if producer_stream is not None:
        ctx.wait_stream(producer_stream)
    tv = dlpack_view(capsule)
    if (tv.dtype_code, tv.dtype_bits) != (2, 32) or not tv.contiguous:
        raise ValueError("expected a C-contiguous float32 tensor")
    return DeviceArray(ctx, tv.data, tuple(tv.shape[i] for i in range(tv.ndim)), "f4", owner=capsule)
