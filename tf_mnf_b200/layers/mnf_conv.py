"""MNFConv2D on libmnf_b200 (reference: tf_mnf/layers/mnf_conv.py:11-197)."""
from __future__ import annotations

import ctypes as C
from typing import Any

from .. import _lib as L
from .. import noise as nz
from ..runtime import DeviceArray, as_device, default_context
from .base import MNFLayerBase, _p


def conv_out_hw(H: int, W: int, kh: int, kw: int, stride: int, padding: str) -> tuple[int, int]:
    """TF output-size rule for VALID / SAME (host-side twin of mnf_conv_out_hw)."""
    if padding == "SAME":
        return -(-H // stride), -(-W // stride)
    if H < kh or W < kw:
        raise ValueError(f"VALID convolution: input {H}x{W} is smaller than the kernel {kh}x{kw}")
    return (H - kh) // stride + 1, (W - kw) // stride + 1


class McActivation:
    """The output of a first MNFConv2D layer under Monte-Carlo prediction (call_mc(..., lazy=True)), NOT yet
    expanded: per-image pre-noise maps + the per-row noise specification.  The next MNFConv2D consumes it
    directly (mnf_conv2d_forward_mc: its tensor-core kernel generates the rows on the fly, so the [B, H, W, F]
    activation never exists in memory); anything else calls materialize()."""

    def __init__(self, ctx, B: int, n: int, Ho: int, Wo: int, F: int, mean, sd, z, eps, pooled: bool) -> None:
        self.ctx, self.B, self.n, self.Ho, self.Wo, self.F = ctx, B, n, Ho, Wo, F
        self.mean, self.sd, self.z, self.eps, self.pooled = mean, sd, z, eps, pooled
        self.shape = (B, Ho // 2, Wo // 2, F) if pooled else (B, Ho, Wo, F)
        self.ndim = 4

    def source(self) -> L.McSource:
        s = L.McSource()
        s.n_rep, s.Ho, s.Wo, s.F = self.n, self.Ho, self.Wo, self.F
        s.mean, s.sd, s.z, s.eps = self.mean.ptr, self.sd.ptr, (self.z.ptr if self.z is not None else None), self.eps
        return s

    def materialize(self) -> DeviceArray:
        ctx = self.ctx
        y = ctx.empty(self.shape)
        L.check(ctx.lib.mnf_conv2d_mc_expand(ctx.handle, self.B, self.n, self.Ho, self.Wo, self.F, _p(self.mean), _p(self.sd),
                                             _p(self.z), C.byref(self.eps), int(self.pooled), _p(y)))
        return y


class MNFConv2D(MNFLayerBase):
    """Bayesian conv2d layer, multiplicative noise on the OUTPUT filters (z scales conv+bias,
    mnf_conv.py:193).  `strides` / `padding` are the constructor arguments the reference's
    callers already pass (mnf_lenet.py:22,25) but its `call` forgets to use (App. B-1)."""

    _conv = True

    def __init__(self, n_filters: int, kernel_size, strides: int = 1, padding: str = "VALID", **kwargs: Any) -> None:
        self.n_filters = int(n_filters)
        self.kernel_size = [kernel_size, kernel_size] if isinstance(kernel_size, int) else list(kernel_size)
        if len(self.kernel_size) != 2:
            raise ValueError("kernel_size must be an int or a pair")
        if isinstance(strides, (tuple, list)):
            if len(set(strides)) != 1:
                raise NotImplementedError("different strides per spatial dimension")
            strides = strides[0]
        self.strides = int(strides)
        self.padding = str(padding).upper()
        if self.padding not in ("VALID", "SAME"):
            raise ValueError(f"padding must be 'VALID' or 'SAME', got {padding!r}")
        self.input_dim: int | None = None
        self.stack_size: int | None = None
        super().__init__(**kwargs)

    def _flow_dim(self) -> int:
        return self.n_filters

    def _kl_shape(self):
        return self.input_dim, self.n_filters

    def build(self, input_shape) -> None:
        self.stack_size = int(input_shape[-1])
        n_rows, n_cols = self.kernel_size
        self.input_dim = n_cols * self.stack_size * n_rows
        self._build_common((n_rows, n_cols, self.stack_size, self.n_filters), self.input_dim)

    def _shape(self, x: DeviceArray) -> L.ConvShape:
        s = L.ConvShape()
        s.B, s.H, s.W, s.C = x.shape
        s.kh, s.kw = self.kernel_size
        s.F, s.stride, s.same = self.n_filters, self.strides, int(self.padding == "SAME")
        return s

    def call(self, x: Any, noise: dict | None = None, fuse_relu_pool: bool = False) -> DeviceArray:  # noqa: A002
        """mnf_conv.py:182-197.  noise: {eps_z [B,F], bern_q [...], eps_out [B,Ho,Wo,F]}.
        fuse_relu_pool: ask the kernel to also apply the ReLU + MaxPool2D(2x2) that follows the
        layer in MNFLeNet (mnf_lenet.py:23-24); `self.last_fused` tells whether it did (it does
        not when training, or for shapes the fused epilogue does not cover)."""
        self.ctx = self.ctx or default_context()
        lazy = x if isinstance(x, McActivation) else None
        if lazy is not None and (not lazy.pooled or self.training or noise or not fuse_relu_pool or lazy.ctx is not self.ctx):
            x, lazy = lazy.materialize(), None
        if lazy is None:
            x = as_device(x, self.ctx)
        if x.ndim != 4:
            raise ValueError(f"MNFConv2D expects NHWC input, got shape {x.shape}")
        if not self.built:
            self.build(x.shape)
        if x.shape[3] != self.stack_size:
            raise ValueError(f"MNFConv2D built for {self.stack_size} input channels, got {x.shape[3]}")
        ctx, B = self.ctx, x.shape[0]
        Ho, Wo = conv_out_hw(x.shape[1], x.shape[2], *self.kernel_size, self.strides, self.padding)
        inj = noise or {}
        z, _, run = self._sample_z(B, nz.CALL_Z, nz.BERN_Q_CALL, inj, self.training, self.row_offset)
        if lazy is not None:
            # the first layer's rows are generated inside this layer's kernel (mnf_conv2d_forward_mc)
            eps = nz.make(self.seed, self.call_index, self.uid, nz.CALL_OUT, self.row_offset, None)
            s = self._shape(x)
            s.fuse_relu_pool = 1
            src, used = lazy.source(), C.c_int32(0)
            y = ctx.empty((B, Ho // 2, Wo // 2, self.n_filters))
            L.check(ctx.lib.mnf_conv2d_forward_mc(ctx.handle, C.byref(s), C.byref(src), _p(z), _p(self.mean_W), _p(self.log_std_W),
                                                  _p(self.mean_b), _p(self.log_var_b), self.max_std, C.byref(eps), _p(y),
                                                  C.byref(used)))
            if used.value:
                self.last_fused = True
                self.call_index += 1
                return y
            x = lazy.materialize()  # outside the fused kernel's domain: expand, then the ordinary path
        oshape = (B, Ho, Wo, self.n_filters)
        eo = as_device(inj["eps_out"], ctx) if inj.get("eps_out") is not None else None
        if eo is not None and eo.shape != oshape:
            raise ValueError(f"injected eps_out must have shape {oshape}, got {eo.shape}")
        eps = nz.make(self.seed, self.call_index, self.uid, nz.CALL_OUT, self.row_offset, eo)
        s = self._shape(x)
        self.last_fused = False
        if fuse_relu_pool and not self.training:
            ok = C.c_int32(0)
            L.check(ctx.lib.mnf_conv2d_can_fuse_pool(ctx.handle, C.byref(s), C.byref(ok)))
            if ok.value:
                s.fuse_relu_pool = 1
                self.last_fused = True
        y = ctx.empty((B, Ho // 2, Wo // 2, self.n_filters) if self.last_fused else oshape)
        mean = ctx.empty(oshape) if self.training else None
        sd = ctx.empty(oshape) if self.training else None
        L.check(ctx.lib.mnf_conv2d_forward(ctx.handle, C.byref(s), _p(x), _p(z), _p(self.mean_W), _p(self.log_std_W),
                                           _p(self.mean_b), _p(self.log_var_b), self.max_std, C.byref(eps),
                                           _p(y), _p(mean), _p(sd)))
        if self.training:
            self._tape = dict(x=x, z=z, run=run, mean=mean, sd=sd, eps=eps, eps_keep=eo, shape=s)
        self.call_index += 1
        return y

    __call__ = call

    def mc_supported(self, x_shape, fuse_relu_pool: bool = False) -> bool:
        """Whether call_mc() covers this input shape (else tile the rows and use call())."""
        if self.training:
            return False
        ctx = self.ctx or default_context()
        Ho, Wo = conv_out_hw(x_shape[1], x_shape[2], *self.kernel_size, self.strides, self.padding)
        ok = C.c_int32(0)
        L.check(ctx.lib.mnf_conv2d_mc_can_expand(ctx.handle, Ho, Wo, self.n_filters, int(fuse_relu_pool), C.byref(ok)))
        return bool(ok.value)

    def call_mc(self, x: Any, n_samples: int, noise: dict | None = None, fuse_relu_pool: bool = False,  # noqa: A002
                lazy: bool = False) -> DeviceArray | McActivation:
        """call(np.repeat(x, n_samples, axis=0)) for a FIRST layer: z and eps enter
        after the contractions (mnf_conv.py:190-197), so the two convolutions are evaluated once per
        distinct input row and only the noise is drawn per (row, sample).  Same Philox counters as
        call() on the repeated batch; inference only.  noise: {eps_z [B,F], bern_q, eps_out [B,Ho,Wo,F]}
        with B = len(x) * n_samples.  lazy=True returns an McActivation (nothing expanded yet) for the next layer."""
        self.ctx = self.ctx or default_context()
        x = as_device(x, self.ctx)
        if x.ndim != 4:
            raise ValueError(f"MNFConv2D expects NHWC input, got shape {x.shape}")
        if not self.built:
            self.build(x.shape)
        if x.shape[3] != self.stack_size:
            raise ValueError(f"MNFConv2D built for {self.stack_size} input channels, got {x.shape[3]}")
        if self.training:
            raise RuntimeError("call_mc() is an inference path: backward() needs the per-row tape of call()")
        ctx, n, rows = self.ctx, int(n_samples), x.shape[0]
        B = rows * n
        Ho, Wo = conv_out_hw(x.shape[1], x.shape[2], *self.kernel_size, self.strides, self.padding)
        F = self.n_filters
        inj = noise or {}
        z, _, _ = self._sample_z(B, nz.CALL_Z, nz.BERN_Q_CALL, inj, False, self.row_offset)
        eo = as_device(inj["eps_out"], ctx) if inj.get("eps_out") is not None else None
        if eo is not None and eo.shape != (B, Ho, Wo, F):
            raise ValueError(f"injected eps_out must have shape {(B, Ho, Wo, F)}, got {eo.shape}")
        # pre-noise mean / standard-deviation maps of the distinct rows
        s = self._shape(x)
        mean, sd = ctx.empty((rows, Ho, Wo, F)), ctx.empty((rows, Ho, Wo, F))
        L.check(ctx.lib.mnf_conv2d_mc_maps(ctx.handle, C.byref(s), _p(x), _p(self.mean_W), _p(self.log_std_W),
                                           _p(self.mean_b), _p(self.log_var_b), self.max_std, _p(mean), _p(sd)))
        eps = nz.make(self.seed, self.call_index, self.uid, nz.CALL_OUT, self.row_offset, eo)
        self.last_fused = bool(fuse_relu_pool)
        self.call_index += 1
        act = McActivation(ctx, B, n, Ho, Wo, F, mean, sd, z, eps, bool(fuse_relu_pool))
        act._keep = eo
        # lazy: leave the per-row expansion to the consumer (the next MNFConv2D's kernel draws this layer's noise)
        return act if (lazy and eo is None) else act.materialize()

    def backward(self, dy: Any, need_dx: bool = True, fresh_grads: bool = False) -> DeviceArray | None:
        """Gradient of the most recent call() (made with `training=True`) given dL/dy:
        returns dL/dx (None with need_dx=False: the first layer of a model has no use for it)
        and ACCUMULATES parameter gradients into self.grads (fresh_grads=True: the caller has just
        zeroed them, so the kernels write the four weight / bias gradients in place instead of into
        temporaries that are then added) (App. A-1c, A-2)."""
        t = self._tape
        if not t:
            raise RuntimeError(f"{self.name}: call the layer with layer.training = True before backward()")
        if not self.grads:
            self.zero_grads()
        ctx = self.ctx
        dy = as_device(dy, ctx)
        x, z = t["x"], t["z"]
        if dy.shape != t["sd"].shape:
            raise ValueError(f"dy must have shape {t['sd'].shape}, got {dy.shape}")
        dx = ctx.empty(x.shape) if need_dx else None
        dz = ctx.empty((x.shape[0], self.n_filters)) if self.use_z else None
        names = ("mean_W", "log_std_W", "mean_b", "log_var_b")
        tmp = {n: self.grads[n] for n in names} if fresh_grads else {n: ctx.empty(getattr(self, n).shape) for n in names}
        L.check(ctx.lib.mnf_conv2d_backward(ctx.handle, C.byref(t["shape"]), _p(x), _p(z), _p(self.mean_W),
                                            _p(self.log_std_W), _p(self.log_var_b), self.max_std, C.byref(t["eps"]),
                                            _p(t["mean"]), _p(t["sd"]), _p(dy), _p(dx), _p(dz), _p(tmp["mean_W"]),
                                            _p(tmp["log_std_W"]), _p(tmp["mean_b"]), _p(tmp["log_var_b"])))
        if not fresh_grads:
            for n, buf in tmp.items():
                self._add_grad(n, buf)
        if self.use_z:
            self._sample_z_backward(t["run"], dz, None)
        return dx
