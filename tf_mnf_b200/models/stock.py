"""The stock Keras layers the reference models put between MNF layers (mnf_lenet.py:23-32,
mnf_feed_forward.py:20), as thin host objects over libmnf_b200's element-wise kernels, and
a minimal Sequential container with the reference's `.layers` / `__call__` protocol."""
from __future__ import annotations

from typing import Any

import numpy as np

from .. import _lib as L
from ..layers.base import _p
from ..runtime import DeviceArray, as_device, default_context


class StockLayer:
    training = False

    def backward(self, dy: DeviceArray) -> DeviceArray:  # pragma: no cover - abstract
        raise NotImplementedError


class ReLU(StockLayer):
    def __call__(self, x: DeviceArray) -> DeviceArray:
        y = x.ctx.empty(x.shape)
        L.check(x.ctx.lib.mnf_relu(x.ctx.handle, x.size, _p(x), _p(y)))
        if self.training:
            self._x = x
        return y

    def backward(self, dy):
        dx = dy.ctx.empty(self._x.shape)
        L.check(dy.ctx.lib.mnf_relu_backward(dy.ctx.handle, dx.size, _p(self._x), _p(dy), _p(dx)))
        return dx


class MaxPool2D(StockLayer):
    """2x2 / stride 2 (Keras defaults).  padding="SAME" as in mnf_lenet.py:24,27."""

    def __init__(self, pool_size=2, strides=None, padding: str = "VALID") -> None:
        ps = pool_size if isinstance(pool_size, int) else pool_size[0]
        st = ps if strides is None else (strides if isinstance(strides, int) else strides[0])
        if ps != 2 or st != 2:
            raise NotImplementedError("only the 2x2 / stride-2 pool of MNFLeNet is implemented")
        self.padding = padding.upper()
        self.fuse_relu = False

    def _pool(self, x: DeviceArray, argmax) -> DeviceArray:
        B, H, W, Cc = x.shape
        if self.padding == "VALID" and (H % 2 or W % 2):
            raise NotImplementedError("VALID max-pool on odd sizes")
        y = x.ctx.empty((B, -(-H // 2), -(-W // 2), Cc))
        L.check(x.ctx.lib.mnf_relu_maxpool2(x.ctx.handle, B, H, W, Cc, _p(x), _p(y), _p(argmax)))
        return y


class ReLUMaxPool2D(MaxPool2D):
    """ReLU + MaxPool2D fused into one pass (relu(max) == max(relu))."""

    def __call__(self, x: DeviceArray) -> DeviceArray:
        am = None
        if self.training:
            B, H, W, Cc = x.shape
            am = x.ctx.empty((B, -(-H // 2), -(-W // 2), Cc), "i1")
            self._x, self._am = x, am
        return self._pool(x, am)

    def backward(self, dy):
        x = self._x
        dx = x.ctx.empty(x.shape)
        L.check(x.ctx.lib.mnf_relu_maxpool2_backward(x.ctx.handle, *x.shape, _p(x), _p(self._am), _p(dy), _p(dx)))
        return dx


class Flatten(StockLayer):
    def __call__(self, x: DeviceArray) -> DeviceArray:
        self._shape = x.shape
        return x.reshape(x.shape[0], -1)

    def backward(self, dy):
        return dy.reshape(self._shape)


class Softmax(StockLayer):
    def __call__(self, x: DeviceArray) -> DeviceArray:
        y = x.ctx.empty(x.shape)
        L.check(x.ctx.lib.mnf_softmax(x.ctx.handle, x.shape[0], x.shape[1], _p(x), _p(y)))
        return y


class BatchNormalization(StockLayer):
    """tf.keras.layers.BatchNormalization() as MNFFeedForward uses it (mnf_feed_forward.py:20): last
    axis, momentum 0.99, epsilon 1e-3, trainable gamma / beta, moving mean / variance.

    Keras picks the mode per call: `model.fit` (bandgap_regression.py:76) normalises with the batch
    statistics and updates the moving averages; a plain `model(x)` -- the reference's GradientTape
    loop (:106-107) and its predictions (:83) -- normalises with the moving statistics while gamma
    and beta still receive gradients.  Here `batch_stats` selects the mode of a TRAINING call
    (`layer.training`): True = fit semantics (default, what Trainer runs), False = the tape loop's.
    Inference calls always use the moving statistics."""

    def __init__(self, momentum: float = 0.99, epsilon: float = 1e-3, batch_stats: bool = True) -> None:
        self.momentum, self.epsilon, self.batch_stats = float(momentum), float(epsilon), bool(batch_stats)
        self.built = False
        self.grads: dict = {}
        self.name = "batch_normalization"

    def build(self, input_shape, ctx=None) -> None:
        ctx = ctx or default_context()
        n = int(input_shape[-1])
        self.gamma, self.beta = ctx.to_device(np.ones(n, np.float32)), ctx.zeros((n,))
        self.moving_mean, self.moving_variance = ctx.zeros((n,)), ctx.to_device(np.ones(n, np.float32))
        self.built = True

    @property
    def trainable_variables(self) -> dict:
        return {"gamma": self.gamma, "beta": self.beta} if self.built else {}

    def variables(self) -> dict:
        return {**self.trainable_variables, "moving_mean": self.moving_mean, "moving_variance": self.moving_variance}

    def zero_grads(self) -> dict:
        for name, v in self.trainable_variables.items():
            g = self.grads.get(name)
            if g is None or g.shape != v.shape:
                self.grads[name] = v.ctx.zeros(v.shape)
            else:
                L.check(v.ctx.lib.mnf_memset(v.ctx.handle, _p(g), 0, g.nbytes))
        return self.grads

    def __call__(self, x: DeviceArray) -> DeviceArray:
        if x.ndim != 2:
            raise NotImplementedError("BatchNormalization: [batch, features] inputs (MNFFeedForward) only")
        if not self.built:
            self.build(x.shape, x.ctx)
        ctx, (B, N) = x.ctx, x.shape
        y = ctx.empty(x.shape)
        mode = int(self.training and self.batch_stats)
        sm = si = None
        if self.training:
            sm, si = ctx.empty((N,)), ctx.empty((N,))
            self._tape = dict(x=x, mean=sm, invstd=si, mode=mode)
        L.check(ctx.lib.mnf_batchnorm_forward(ctx.handle, B, N, _p(x), _p(self.gamma), _p(self.beta), _p(self.moving_mean),
                                              _p(self.moving_variance), self.epsilon, self.momentum, mode, _p(y), _p(sm), _p(si)))
        return y

    def backward(self, dy: DeviceArray, fresh_grads: bool = False) -> DeviceArray:
        """dL/dx; ACCUMULATES d gamma / d beta into self.grads (fresh_grads: written in place)."""
        t = self._tape
        ctx, x = dy.ctx, t["x"]
        if not self.grads:
            self.zero_grads()
        dx = ctx.empty(x.shape)
        tmp = self.grads if fresh_grads else {n: ctx.empty(v.shape) for n, v in self.trainable_variables.items()}
        L.check(ctx.lib.mnf_batchnorm_backward(ctx.handle, x.shape[0], x.shape[1], _p(x), _p(self.gamma), _p(t["mean"]),
                                               _p(t["invstd"]), t["mode"], _p(dy), _p(dx), _p(tmp["gamma"]), _p(tmp["beta"])))
        if not fresh_grads:
            for n, buf in tmp.items():
                L.check(ctx.lib.mnf_axpy(ctx.handle, buf.size, 1.0, _p(buf), _p(self.grads[n])))
        return dx


class Sequential:
    def __init__(self, layers=None) -> None:
        self.layers = list(layers or [])
        self.presample_z = True   # draw the later layers' z on side streams beside the first layers
        self.presample_first = True  # ... and the first layer's z one call ahead (off inside captured graphs)
        self.mc_share_first = True   # mc_predict(): a first MNFConv2D convolves each distinct row once
        # ... and a second MNFConv2D can generate the first one's rows inside its own kernel (mnf_conv2d_forward_mc).
        # Off by default: bit-identical, but measured SLOWER on B200 (conv1 + conv2 470 us against 150 + 218 us) --
        # the conv kernel's 10 producer warps per SM cannot hide the latency of the Philox chains and of the per-image
        # map reads that 64 resident warps per SM hide in conv_mc_expand.
        self.mc_fuse_second = False
        self._sides: list = []    # one forked context per later MNF layer: their flow chains
        #                           (80-CTA kernels at 10k rows) pack the SMs side by side

    def __call__(self, x: Any, noise: list | None = None, mc_samples: int = 0) -> DeviceArray:  # noqa: A002
        """noise: optional list with one injected-draw dict per MNF layer (program order).
        mc_samples = n > 0: x holds DISTINCT rows and the result is that of the batch
        np.repeat(x, n, axis=0); the first layer (an MNFConv2D that supports it) evaluates its
        contractions once per distinct row (MNFConv2D.call_mc) -- see mc_predict()."""
        x = as_device(x, default_context())
        it = iter(noise or [])
        skip = False
        n_mc = int(mc_samples)
        rows = x.shape[0] * (n_mc or 1)
        self._first_rows = rows
        first = None
        if self.presample_z and not noise:
            mnf = [l for l in self.mnf_layers if l.built and l.use_z and l.ctx is x.ctx]
            if mnf:
                if len(self._sides) < len(mnf) or self._sides[0].device != x.ctx.device:
                    # layer 0's side stream draws the z of the NEXT call (least priority); the later layers' chains are
                    # needed in program order by THIS call: the earlier the consumer, the higher the rank
                    self._sides = [x.ctx.fork(rank=0 if i == 0 else max(1, 5 - i)) for i in range(len(mnf))]
                for lyr, side in zip(mnf, self._sides):
                    if lyr is self.mnf_layers[0]:
                        if self.presample_first:
                            first = (lyr, side)   # its z for THIS call was drawn during the previous call
                    else:
                        lyr.presample(rows, side)
        for i, lyr in enumerate(self.layers):
            if skip:  # this ReLU+MaxPool was folded into the preceding conv's epilogue
                skip = False
                continue
            if hasattr(lyr, "kl_div"):
                nxt = self.layers[i + 1] if i + 1 < len(self.layers) else None
                if n_mc and i == 0:
                    fuse = isinstance(nxt, ReLUMaxPool2D) and lyr.mc_supported(x.shape, True)
                    # conv -> ReLU/pool -> conv -> ReLU/pool (MNF-LeNet): the second conv's kernel generates the first
                    # layer's rows itself (MNFConv2D.call on an McActivation)
                    nn = self.layers[i + 2] if i + 2 < len(self.layers) else None
                    nnn = self.layers[i + 3] if i + 3 < len(self.layers) else None
                    lazy = (fuse and self.mc_fuse_second and not noise and hasattr(nn, "n_filters") and not nn.training
                            and isinstance(nnn, ReLUMaxPool2D))
                    x = lyr.call_mc(x, n_mc, noise=next(it, None), fuse_relu_pool=fuse, lazy=lazy)
                    skip = lyr.last_fused
                elif isinstance(nxt, ReLUMaxPool2D) and hasattr(lyr, "n_filters") and not lyr.training:
                    x = lyr.call(x, noise=next(it, None), fuse_relu_pool=True)
                    skip = lyr.last_fused
                elif type(nxt) in (ReLU, Softmax) and hasattr(lyr, "n_out") and not lyr.training:
                    x = lyr.call(x, noise=next(it, None), activation="relu" if type(nxt) is ReLU else "softmax")
                    skip = lyr.last_fused
                else:
                    x = lyr(x, noise=next(it, None))
                if first is not None and lyr is first[0]:
                    # the first layer's z has nothing in front of it to hide behind: draw the one
                    # for the NEXT call now, beside the rest of this call (same counters as in-line)
                    lyr.presample(self._first_rows, first[1])
            else:
                if hasattr(x, "materialize"):
                    x = x.materialize()
                x = lyr(x)
        if hasattr(x, "materialize"):
            x = x.materialize()
        if first is not None:
            x.ctx.wait(first[1])  # a call is complete when its stream is: include the draw made for the next one
        return x

    def freeze_weights(self, frozen: bool = True) -> None:
        """Promise that no parameter changes until unfrozen: the contexts this model runs on keep
        their packed weight blocks across calls (Context.freeze_weights).  Outputs are unchanged."""
        ctxs = {id(l.ctx): l.ctx for l in self.mnf_layers if l.ctx is not None}
        for c in list(ctxs.values()) + list(self._sides):
            c.freeze_weights(frozen)

    def mc_predict(self, x: Any, n_samples: int) -> DeviceArray:
        """Monte-Carlo predictive: every input row under `n_samples` independent posterior draws
        (what the reference's notebooks do by calling the model n_samples times on one batch).
        The repeated batch np.repeat(x, n_samples, axis=0) is formed on the device, so only the
        distinct rows cross PCIe; returns [len(x) * n_samples, ...] in that row order.  When the
        first layer is an MNFConv2D its two convolutions are evaluated once per distinct row and
        only the noise is drawn per sample (mc_share_first; same Philox counters either way)."""
        x = as_device(x, default_context())
        n = int(n_samples)
        if n < 1:
            raise ValueError("n_samples must be >= 1")
        first = self.layers[0] if self.layers else None
        if (n > 1 and self.mc_share_first and hasattr(first, "call_mc") and (first.built or x.ndim == 4)
                and x.ndim == 4 and first.mc_supported(x.shape)):
            # the samples of an image share the first layer's contractions
            return self(x, mc_samples=n)
        rows = x.shape[0]
        tiled = x.ctx.empty((rows * n, *x.shape[1:]))
        L.check(x.ctx.lib.mnf_tile_rows(x.ctx.handle, rows, x.size // max(rows, 1), n, _p(x), _p(tiled)))
        return self(tiled)

    def kl_div(self, noise: list | None = None, ctx: Any = None) -> DeviceArray:  # noqa: A002
        """Sum of the layers' KL terms (mnf_lenet.py:35-39, mnf_feed_forward.py:23-28).
        `ctx`: optional sibling context (Context.fork()), or a list of them, so the KL runs beside
        a forward pass: layer i uses ctx[i % len(ctx)], the sum is formed on ctx[0] (join the
        caller's context with `main.wait(ctx[0])`)."""
        it = iter(noise or [])
        lanes = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx]
        total, parts = None, []
        for i, lyr in enumerate(self.mnf_layers):
            lane = lanes[i % len(lanes)]
            kl = lyr.kl_div(noise=next(it, None), ctx=lane)
            if total is None:
                total = kl
            else:
                parts.append(kl)
        for kl in parts:
            if kl.ctx is not total.ctx:
                total.ctx.wait(kl.ctx)
            L.check(total.ctx.lib.mnf_axpy(total.ctx.handle, 1, 1.0, _p(kl), _p(total)))
        for c in {id(k.ctx): k.ctx for k in parts if k.ctx is not total.ctx}.values():
            c.wait(total.ctx)  # the partial terms are released (stream-ordered) only behind the sum
        return total

    @property
    def mnf_layers(self) -> list:
        return [l for l in self.layers if hasattr(l, "kl_div")]

    @property
    def param_layers(self) -> list:
        """Every layer that owns trainable variables: the MNF layers and BatchNormalization."""
        return [l for l in self.layers if hasattr(type(l), "trainable_variables")]

    def set_training(self, flag: bool) -> None:
        for lyr in self.layers:
            lyr.training = flag

    def set_row_offset(self, off: int) -> None:
        for lyr in self.mnf_layers:
            lyr.row_offset = int(off)
