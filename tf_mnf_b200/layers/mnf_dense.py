"""MNFDense on libmnf_b200 (reference: tf_mnf/layers/mnf_dense.py:11-170)."""
from __future__ import annotations

import ctypes as C
from typing import Any

from .. import _lib as L
from .. import noise as nz
from ..runtime import DeviceArray, as_device, default_context
from .base import MNFLayerBase, _p


class MNFDense(MNFLayerBase):
    """Bayesian dense layer with multiplicative normalizing-flow noise on the INPUT units.
    Same constructor and methods as the reference class; tensors are DeviceArrays (DLPack /
    __cuda_array_interface__), or NumPy arrays which are copied to the GPU first."""

    _conv = False

    def __init__(self, n_out: int, **kwargs: Any) -> None:
        self.n_out = int(n_out)
        self.n_in: int | None = None
        super().__init__(**kwargs)

    def _flow_dim(self) -> int:
        return self.n_in

    def _kl_shape(self):
        return self.n_in, self.n_out

    def build(self, input_shape) -> None:
        self.n_in = int(input_shape[-1])
        self._build_common((self.n_in, self.n_out), self.n_in)

    def call(self, x: Any, noise: dict | None = None, activation: str | None = None) -> DeviceArray:  # noqa: A002
        """mnf_dense.py:159-170.  noise: optional injected draws {eps_z [B,n_in], bern_q [...], eps_out [B,n_out]}.
        activation: "relu" / "softmax" asks the kernel to also apply the stock layer that follows in
        MNFLeNet / MNFFeedForward (mnf_lenet.py:29-32); `self.last_fused` tells whether it did (not
        when training, nor for shapes / math modes whose epilogue does not cover it)."""
        self.ctx = self.ctx or default_context()
        x = as_device(x, self.ctx)
        if x.ndim != 2:
            raise ValueError(f"MNFDense expects [batch, features], got shape {x.shape}")
        if not self.built:
            self.build(x.shape)
        if x.shape[1] != self.n_in:
            raise ValueError(f"MNFDense built for {self.n_in} input features, got {x.shape[1]}")
        ctx, B = self.ctx, x.shape[0]
        inj = noise or {}
        z, _, run = self._sample_z(B, nz.CALL_Z, nz.BERN_Q_CALL, inj, self.training, self.row_offset)
        eo = as_device(inj["eps_out"], ctx) if inj.get("eps_out") is not None else None
        if eo is not None and eo.shape != (B, self.n_out):
            raise ValueError(f"injected eps_out must have shape {(B, self.n_out)}, got {eo.shape}")
        eps = nz.make(self.seed, self.call_index, self.uid, nz.CALL_OUT, self.row_offset, eo)
        y = ctx.empty((B, self.n_out))
        sd = ctx.empty((B, self.n_out)) if self.training else None
        act = L.MNF_ACT_NONE
        self.last_fused = False
        if activation is not None and not self.training:
            want = {"relu": L.MNF_ACT_RELU, "softmax": L.MNF_ACT_SOFTMAX}[activation]
            ok = C.c_int32(0)
            L.check(ctx.lib.mnf_dense_can_fuse_act(ctx.handle, self.n_in, self.n_out, want, C.byref(ok)))
            if ok.value:
                act, self.last_fused = want, True
        L.check(ctx.lib.mnf_dense_forward_act(ctx.handle, B, self.n_in, self.n_out, _p(x), _p(z), _p(self.mean_W),
                                              _p(self.log_std_W), _p(self.mean_b), _p(self.log_var_b),
                                              self.max_std, C.byref(eps), act, _p(y), _p(sd)))
        if self.training:
            self._tape = dict(x=x, z=z, run=run, sd=sd, eps=eps, eps_keep=eo)
        self.call_index += 1
        return y

    __call__ = call

    def backward(self, dy: Any, need_dx: bool = True, fresh_grads: bool = False) -> DeviceArray | None:
        """Gradient of the most recent call() (made with `training=True`) given dL/dy:
        returns dL/dx (None with need_dx=False: the first layer of a model has no use for it)
        and ACCUMULATES parameter gradients into self.grads (fresh_grads=True: the caller has just
        zeroed them, so the kernels write the four weight / bias gradients in place instead of into
        temporaries that are then added) (App. A-1, A-2)."""
        t = self._tape
        if not t:
            raise RuntimeError(f"{self.name}: call the layer with layer.training = True before backward()")
        if not self.grads:
            self.zero_grads()
        ctx = self.ctx
        dy = as_device(dy, ctx)
        x, z = t["x"], t["z"]
        B = x.shape[0]
        if dy.shape != (B, self.n_out):
            raise ValueError(f"dy must have shape {(B, self.n_out)}, got {dy.shape}")
        dx = ctx.empty(x.shape) if need_dx else None
        dz = ctx.empty(x.shape) if self.use_z else None
        names = ("mean_W", "log_std_W", "mean_b", "log_var_b")
        tmp = {n: self.grads[n] for n in names} if fresh_grads else {n: ctx.empty(getattr(self, n).shape) for n in names}
        L.check(ctx.lib.mnf_dense_backward(ctx.handle, B, self.n_in, self.n_out, _p(x), _p(z), _p(self.mean_W),
                                           _p(self.log_std_W), _p(self.log_var_b), self.max_std, C.byref(t["eps"]),
                                           _p(t["sd"]), _p(dy), _p(dx), _p(dz), _p(tmp["mean_W"]),
                                           _p(tmp["log_std_W"]), _p(tmp["mean_b"]), _p(tmp["log_var_b"])))
        if not fresh_grads:
            for n, buf in tmp.items():
                self._add_grad(n, buf)
        if self.use_z:
            self._sample_z_backward(t["run"], dz, None)
        return dx
