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
from .runtime import DeviceArray, as_device, bump_param_version, default_context


class Trainer:
    def __init__(self, model: Any, loss: str = "xent", kl_weight: float = 1.0, lr: float = 1e-3, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-7, world: int = 1, comm: Any = None) -> None:
        if loss not in ("xent", "mse"):
            raise ValueError("loss must be 'xent' (softmax cross-entropy) or 'mse'")
        self.model, self.loss, self.kl_weight = model, loss, float(kl_weight)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.world, self.comm = int(world), comm
        self.t = 0
        # flat training state (built on the first step): parameters, gradients and the two Adam
        # moments each live in ONE buffer; the layers' variables / grads are views into them
        self.flat_p: DeviceArray | None = None
        self.flat_g: DeviceArray | None = None
        self.flat_m: DeviceArray | None = None
        self.flat_v: DeviceArray | None = None
        # the KL branch (kl_div + its backward: single-row flow chains, latency-bound) runs on a forked
        # context beside the likelihood branch and accumulates into its own flat buffer
        self.kl_ctx = None
        self._z_sides: list = []  # one forked context per MNF layer: the z chains of a step run side by side
        self.flat_gk: DeviceArray | None = None
        self.kl_grads: list[dict] = []
        self._dev_state: DeviceArray | None = None  # {int32 step, float lr_t} once a step has been captured

    def _forward_layers(self):
        layers = list(self.model.layers)
        if self.loss == "xent" and layers and isinstance(layers[-1], Softmax):
            layers = layers[:-1]  # cross-entropy is fused with the softmax
        return layers

    def _ensure_flat(self, x: DeviceArray) -> None:
        """Move every trainable variable into one flat parameter buffer (the live variable objects
        are re-pointed, so layers and flows keep working unchanged) and make the gradients views of
        one flat gradient buffer: zeroing, the data-parallel all-reduce and Adam then are ONE
        memset / collective / launch per step instead of one per tensor (LeNet: 132 tensors)."""
        if self.flat_p is not None:
            return
        ctx = default_context()
        model = self.model
        if not all(l.built for l in model.param_layers):
            idx = [(l.call_index, l.kl_index) for l in model.mnf_layers]
            model.set_training(False)
            model(x)  # Keras build protocol: variables are created by the first call
            for l, (ci, ki) in zip(model.mnf_layers, idx):
                l.call_index, l.kl_index = ci, ki
        for c in [ctx, *getattr(model, "_sides", [])]:
            c.sync()
        for l in model.mnf_layers:
            l._pre = None  # a z drawn ahead carries pointers to the old parameter buffers
        for l in model.param_layers:
            l.zero_grads()
        # MNF layers and BatchNormalization (gamma, beta); moving statistics are not trained
        items = [(lyr, name, v) for lyr in model.param_layers for name, v in lyr.trainable_variables.items()]
        offs, total = [], 0
        for _, _, v in items:
            offs.append(total)
            total += (v.size + 63) // 64 * 64  # 256-byte aligned slots
        self.flat_p, self.flat_g = ctx.zeros((total,)), ctx.zeros((total,))
        self.flat_m, self.flat_v = ctx.zeros((total,)), ctx.zeros((total,))
        self.flat_gk = ctx.zeros((total,))
        self.kl_ctx = ctx.fork()
        per_layer = {id(l): {} for l in model.param_layers}
        for (lyr, name, v), off in zip(items, offs):
            slot = self.flat_p[off:off + v.size].reshape(v.shape)
            slot.assign(v)
            ctx.sync()
            v.ptr, v.owner = slot.ptr, slot.owner
            lyr.grads[name] = self.flat_g[off:off + v.size].reshape(v.shape)
            per_layer[id(lyr)][name] = self.flat_gk[off:off + v.size].reshape(v.shape)
        self.kl_grads = [per_layer[id(l)] for l in model.mnf_layers]
        ctx.sync()

    def loss_and_grads(self, x: Any, target: Any, noise: list | None = None):
        """Forward + backward: returns (loss [1], kl [1]); gradients land in layer.grads."""
        ctx = default_context()
        model = self.model
        x = as_device(x, ctx)
        target = as_device(target, ctx)
        self._ensure_flat(x)
        model.set_training(True)
        layers = self._forward_layers()
        it = iter(noise or [])
        if not noise:
            # z does not depend on x: every layer's flow chain (one CTA per 128 rows, latency-bound at
            # minibatch sizes) is enqueued up front on its own stream instead of in line, one after another
            # (only the wide ones: for a 50-wide chain the extra stream fork / join costs the host more than it saves)
            zl = [l for l in model.mnf_layers if l.use_z and l._flow_dim() >= 256]
            while len(self._z_sides) < len(zl):
                self._z_sides.append(ctx.fork())
            for lyr, side in zip(zl, self._z_sides):
                side.wait(ctx)  # behind the previous step's parameter update
                lyr.presample(x.shape[0], side)
        h = x
        for lyr in layers:
            h = lyr(h, noise=next(it, None)) if hasattr(lyr, "kl_div") else lyr(h)
        L.check(ctx.lib.mnf_memset(ctx.handle, _p(self.flat_g), 0, self.flat_g.nbytes))
        B = h.shape[0]
        gb = B * self.world  # global batch: every rank holds B rows
        loss = ctx.empty((1,))
        dh = ctx.empty(h.shape)
        if self.loss == "xent":
            L.check(ctx.lib.mnf_softmax_xent(ctx.handle, B, h.shape[1], _p(h), _p(target), 1.0 / gb, _p(loss), _p(dh)))
        else:
            L.check(ctx.lib.mnf_mse_loss(ctx.handle, h.size, _p(h), _p(target), 1.0 / (gb * h.size // B), _p(loss), _p(dh)))
        # KL branch on the forked context (ordered after everything enqueued so far, i.e. after the
        # previous step's Adam update); it does not depend on x, so it runs beside the backward pass
        kc = self.kl_ctx
        kc.wait(ctx)
        L.check(kc.lib.mnf_memset(kc.handle, _p(self.flat_gk), 0, self.flat_gk.nbytes))
        kl = model.kl_div(ctx=kc)
        for lyr, kg in zip(model.mnf_layers, self.kl_grads):
            # summed over ranks = contributed once; the KL buffer was zeroed above: outputs land in place
            lyr.kl_backward(self.kl_weight / self.world, ctx=kc, grads=kg, fresh_grads=True)
        for i in range(len(layers) - 1, -1, -1):
            mnf = hasattr(layers[i], "kl_div")  # MNF layers: the flat gradient buffer was zeroed above
            if i == 0 and mnf:
                layers[0].backward(dh, need_dx=False, fresh_grads=True)  # nobody consumes dL/d(input)
            elif mnf or hasattr(type(layers[i]), "trainable_variables"):
                dh = layers[i].backward(dh, fresh_grads=True)
            else:
                dh = layers[i].backward(dh)
        ctx.wait(kc)
        L.check(ctx.lib.mnf_axpy(ctx.handle, self.flat_g.size, 1.0, _p(self.flat_gk), _p(self.flat_g)))
        kc.wait(ctx)  # the KL buffers are re-used next step only after this sum
        model.set_training(False)
        return loss, kl

    def _named_params(self):
        for li, lyr in enumerate(self.model.param_layers):
            for name, v in lyr.trainable_variables.items():
                yield (li, name), v, lyr.grads[name]

    def allreduce(self) -> None:
        """One in-place sum of the flat gradient buffer over the ranks (mnf_allreduce_grads: NCCL,
        enqueued behind the backward pass on the context's own stream -- no host synchronisation).
        `comm`: a tf_mnf_b200.comm.Communicator, created from the torchrun environment on first use;
        any object with `.allreduce(flat: DeviceArray)` works (the CPU tests plug in a gloo one)."""
        if self.world <= 1:
            return
        if self.comm is None:
            from .comm import Communicator

            self.comm = Communicator(default_context(), world=self.world)
        self.comm.allreduce(self.flat_g)

    def apply(self) -> None:
        ctx = default_context()
        self.t += 1
        if self._dev_state is not None:  # inside / after Trainer.capture: the step count lives on the device
            L.check(ctx.lib.mnf_adam_step_dev(ctx.handle, self.flat_p.size, _p(self.flat_p), _p(self.flat_g), _p(self.flat_m),
                                              _p(self.flat_v), self.lr, self.beta1, self.beta2, self.eps, _p(self._dev_state)))
        else:
            L.check(ctx.lib.mnf_adam_step(ctx.handle, self.flat_p.size, _p(self.flat_p), _p(self.flat_g), _p(self.flat_m),
                                          _p(self.flat_v), self.lr, self.beta1, self.beta2, self.eps, self.t))
        bump_param_version()  # a z drawn ahead of the next call is stale now

    def step(self, x: Any, target: Any):
        loss, kl = self.loss_and_grads(x, target)
        self.allreduce()
        self.apply()
        return loss, kl

    def capture(self, x: Any, target: Any) -> "CapturedTrainStep":
        """The whole step -- forward, loss, backward, the KL branch on its forked stream, Adam -- captured ONCE
        into a CUDA graph (single process; the data-parallel all-reduce stays eager).  A LeNet step is ~170
        small launches whose HOST enqueue (ctypes) costs more than the GPU work; a replay is one launch.
        Noise stays fresh (device-side replay counter added to every Philox call index, graph.py), Adam's
        bias correction follows a device-side step count (mnf_adam_step_dev), and unfrozen weights are
        re-packed by kernels inside the graph (shape-only tables are uploaded once per context).  The warm-up
        passes the capture needs run on a snapshot: parameters, moments, step and noise counters are restored,
        so `cap.step(x, y)` k times equals `trainer.step(x, y)` k times."""
        if self.world > 1:
            raise NotImplementedError("Trainer.capture: single-process steps (the NCCL all-reduce is enqueued eagerly)")
        return CapturedTrainStep(self, x, target)


class CapturedTrainStep:
    def __init__(self, tr: Trainer, x: Any, target: Any) -> None:
        from .graph import CapturedStep

        ctx = default_context()
        self.tr = tr
        self.x = ctx.to_device(np.asarray(x.numpy() if hasattr(x, "numpy") else x, np.float32))
        self.y = ctx.to_device(np.asarray(target.numpy() if hasattr(target, "numpy") else target, np.float32))
        tr._ensure_flat(self.x)
        model = tr.model
        # snapshot of everything the warm-up steps change
        ctx.sync()
        snap = [ctx.empty(b.shape) for b in (tr.flat_p, tr.flat_m, tr.flat_v)]
        for dst, src in zip(snap, (tr.flat_p, tr.flat_m, tr.flat_v)):
            L.check(ctx.lib.mnf_memcpy_d2d(ctx.handle, _p(dst), _p(src), src.nbytes))
        idx = [(l.call_index, l.kl_index) for l in model.mnf_layers]
        t0 = tr.t
        state = np.zeros(2, np.uint32)
        state[0] = t0
        tr._dev_state = ctx.to_device(state, "u4")
        ctx.sync()

        def fn():
            loss, kl = tr.loss_and_grads(self.x, self.y)
            tr.apply()
            return loss, kl

        def restore():
            for dst, src in zip((tr.flat_p, tr.flat_m, tr.flat_v), snap):
                L.check(ctx.lib.mnf_memcpy_d2d(ctx.handle, _p(dst), _p(src), src.nbytes))
            tr._dev_state.assign(state)
            tr.t = t0
            ctx.sync()

        fn()  # the side contexts (z chains of wide layers) are created by the first step
        ctx.sync()
        forks = [c for c in [tr.kl_ctx, *tr._z_sides] if c is not None]
        # CapturedStep runs fn eagerly three times (warm-up + stocking passes) and once under capture (nothing
        # executes then); its launch() advances the layers' counters by one call / one kl_div per replay
        self.cap = CapturedStep(ctx, fn, forks=forks, layers=model.mnf_layers, calls_per_step=1, kl_per_step=1)
        ctx.sync()
        restore()
        # replay 0 must draw with the call indices of the step that was snapshotted: the device-side replay counter
        # is added to the indices the CAPTURE pass used, so rewind the Python counters to those of the capture pass
        # minus the three eager passes, and offset the counter accordingly
        self._idx_capture = [(l.call_index - 1, l.kl_index - 1) for l in model.mnf_layers]  # indices baked into the graph
        back = self._idx_capture[0][0] - idx[0][0] if idx else 0   # eager passes that ran before the capture
        off = np.array([(0x100000000 - back) & 0xFFFFFFFF], np.uint32)  # counter = -back (mod 2^32): index + counter = snapshot index
        self.cap.counter.assign(off)
        for l, (ci, ki) in zip(model.mnf_layers, idx):
            l.call_index, l.kl_index = ci, ki
        self.cap.replays = 1  # launch() then advances the Python-side counters on every replay, the first included
        ctx.sync()

    def step(self, x: Any = None, target: Any = None):
        """Copy the batch into the graph's input buffers (skipped for None) and replay the step; returns the
        (re-used) loss and KL device arrays."""
        ctx = default_context()
        if x is not None:
            self.x.assign(x)
        if target is not None:
            self.y.assign(target)
        out = self.cap.launch()
        self.tr.t += 1
        bump_param_version()
        return out
