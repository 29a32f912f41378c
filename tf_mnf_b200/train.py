"""Training-step closure around the MNF layers: the loss the reference
notebooks minimise (lenet_mnist.py:53-68 mean categorical cross-entropy + kl/len(X_train);
bandgap_regression.py:57-69 MSE + kl/(2*batch)), backward through the model with the
hand-written kernels, an optional data-parallel gradient all-reduce, and Adam
(lenet_mnist.py:47,115) -- what `tape.gradient` + `adam.apply_gradients` do in the reference."""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from . import _lib as L
from .layers.base import _p
from .models.stock import Softmax
from .runtime import DeviceArray, as_device, default_context


class Trainer:
    def __init__(self, model: Any, loss: str = "xent", kl_weight: float = 1.0, lr: float = 1e-3, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-7, world: int = 1, group: Any = None) -> None:
        if loss not in ("xent", "mse"):
            raise ValueError("loss must be 'xent' (softmax cross-entropy) or 'mse'")
        self.model, self.loss, self.kl_weight = model, loss, float(kl_weight)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.world, self.group = int(world), group
        self.t = 0
        self.state: dict[int, tuple[DeviceArray, DeviceArray]] = {}

    def _forward_layers(self):
        layers = list(self.model.layers)
        if self.loss == "xent" and layers and isinstance(layers[-1], Softmax):
            layers = layers[:-1]  # cross-entropy is fused with the softmax
        return layers

    def loss_and_grads(self, x: Any, target: Any, noise: list | None = None):
        """Forward + backward: returns (loss [1], kl [1]); gradients land in layer.grads."""
        ctx = default_context()
        model = self.model
        model.set_training(True)
        x = as_device(x, ctx)
        target = as_device(target, ctx)
        layers = self._forward_layers()
        it = iter(noise or [])
        h = x
        for lyr in layers:
            h = lyr(h, noise=next(it, None)) if hasattr(lyr, "kl_div") else lyr(h)
        for lyr in model.mnf_layers:
            lyr.zero_grads()
        B = h.shape[0]
        gb = B * self.world  # global batch: every rank holds B rows
        loss = ctx.empty((1,))
        dh = ctx.empty(h.shape)
        if self.loss == "xent":
            L.check(ctx.lib.mnf_softmax_xent(ctx.handle, B, h.shape[1], _p(h), _p(target), 1.0 / gb, _p(loss), _p(dh)))
        else:
            L.check(ctx.lib.mnf_mse_loss(ctx.handle, h.size, _p(h), _p(target), 1.0 / (gb * h.size // B), _p(loss), _p(dh)))
        for lyr in reversed(layers):
            dh = lyr.backward(dh)
        kl = model.kl_div()
        for lyr in model.mnf_layers:
            lyr.kl_backward(self.kl_weight / self.world)  # summed over ranks = contributed once
        model.set_training(False)
        return loss, kl

    def _named_params(self):
        for li, lyr in enumerate(self.model.mnf_layers):
            for name, v in lyr.trainable_variables.items():
                yield (li, name), v, lyr.grads[name]

    def allreduce(self) -> None:
        if self.world <= 1:
            return
        import torch

        from .parallel import allreduce_gradients

        default_context().sync()
        grads = [torch.as_tensor(g, device="cuda") for _, _, g in self._named_params()]
        allreduce_gradients(grads, None, self.group)
        torch.cuda.synchronize()

    def apply(self) -> None:
        ctx = default_context()
        self.t += 1
        for key, p, g in self._named_params():
            st = self.state.get(key)
            if st is None:
                st = self.state[key] = (ctx.zeros(p.shape), ctx.zeros(p.shape))
            L.check(ctx.lib.mnf_adam_step(ctx.handle, p.size, _p(p), _p(g), _p(st[0]), _p(st[1]), self.lr, self.beta1,
                                          self.beta2, self.eps, self.t))

    def step(self, x: Any, target: Any):
        loss, kl = self.loss_and_grads(x, target)
        self.allreduce()
        self.apply()
        return loss, kl
