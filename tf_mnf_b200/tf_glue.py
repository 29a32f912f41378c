"""TensorFlow glue: the MNF layers of libmnf_b200 as differentiable TF ops, so that the reference's own
training loop (notebooks/lenet_mnist.py:108-115: `with tf.GradientTape() as tape: loss = loss_fn(
labels, model(x)); grads = tape.gradient(loss, model.trainable_variables); adam.apply_gradients(...)`)
runs unchanged on top of the CUDA kernels.

TensorFlow is import-guarded: nothing here runs at `import tf_mnf_b200` time, and the rest of the
package never imports this module.  Tensors cross the boundary zero-copy through DLPack
(`tf.experimental.dlpack`, available since TF 2.1 -- the reference's pin); the library never sees a
TF type.  TF cannot be installed in the build container, so tests/test_gpu_tf_glue.py drives these
wrappers through a ~60-line stand-in for the five TF entry points used (`custom_gradient`,
`experimental.dlpack.to_dlpack / from_dlpack`, `Variable`, `keras.layers.Layer`).

Design
  * Every MNF layer keeps its parameters in library-owned device buffers (DeviceArray).  A
    `TFMNFDense` / `TFMNFConv2D` Keras layer mirrors them as `tf.Variable`s (same names as the
    reference: mean_W, log_std_W, ..., mnf_dense.py:49-86); before each call the variables' current
    values are copied device-to-device into the library buffers (optimisers assign to tf.Variables).
  * `call` = `tf.custom_gradient` around `layer.call`: forward returns y; the gradient function runs
    `layer.backward(dy)` (mnf_dense_backward / mnf_conv2d_backward + the flow chain's backward) and
    returns dL/dx plus one gradient per mirrored variable.
  * `kl_div` = `tf.custom_gradient` around `layer.kl_div`; its gradient function runs
    `layer.kl_backward(upstream)` -- mnf_kl_backward + both flow chains.
  * Streams: the context adopts nothing; DLPack capsules of TF 2.1 carry no stream, so inputs are
    ordered behind the legacy default stream (`from_dlpack_capsule(..., producer_stream=1)`; pass TF's
    compute stream when known) and outputs are handed over complete (`ctx.sync()` in `__dlpack__`).
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .layers import MNFConv2D, MNFDense
from .runtime import DeviceArray, default_context, from_dlpack_capsule


def _tf():
    try:
        import tensorflow as tf  # noqa: PLC0415
    except ImportError as exc:  # pragma: no cover - TF is absent in the build container
        raise ImportError("tf_mnf_b200.tf_glue needs TensorFlow >= 2.1 (tf.experimental.dlpack); the rest of "
                          "tf_mnf_b200 does not") from exc
    return tf


def to_device(t: Any, ctx=None, producer_stream: int | None = 1) -> DeviceArray:
    """tf.Tensor on the GPU -> borrowed DeviceArray view (zero-copy; the capsule keeps the TF buffer alive)."""
    tf = _tf()
    return from_dlpack_capsule(tf.experimental.dlpack.to_dlpack(t), ctx or default_context(), producer_stream)


def to_tf(a: DeviceArray) -> Any:
    """DeviceArray -> tf.Tensor (zero-copy; the deleter TF calls releases our reference)."""
    tf = _tf()
    return tf.experimental.dlpack.from_dlpack(a.__dlpack__())


class _Mirror:
    """tf.Variable mirrors of one host layer's trainable variables, in a fixed order."""

    def __init__(self, layer: Any, make_variable) -> None:
        self.layer = layer
        self.names = list(layer.trainable_variables)
        self.vars = [make_variable(layer.trainable_variables[n], n) for n in self.names]

    def push(self, values) -> None:
        """Current values of the TF variables -> the library's parameter buffers (device to device)."""
        tv = self.layer.trainable_variables
        for name, t in zip(self.names, values):
            tv[name].assign(to_device(t, tv[name].ctx))

    def grads(self) -> list:
        """layer.grads (filled by backward / kl_backward) as tf.Tensors, in variable order; each is
        copied out of the accumulation buffer so that the next zero_grads() cannot touch it."""
        return [to_tf(self.layer.grads[n].copy()) for n in self.names]


def _make_tf_variable(arr: DeviceArray, name: str):
    tf = _tf()
    return tf.Variable(to_tf(arr.copy()), name=name.replace("/", "_"), trainable=True)


def mnf_call_op(layer: Any, mirror: _Mirror):
    """x, *variables -> y with the gradient of MNFDense.call (mnf_dense.py:159-170) / MNFConv2D.call
    (mnf_conv.py:182-197) as tape.gradient would compute it: App. A-1 / A-1c / A-2 of SURVEY.md."""
    tf = _tf()

    @tf.custom_gradient
    def op(x, *variables):
        mirror.push(variables)
        layer.training = True  # keep what backward needs (z, flow states, sqrt(V), the Philox spec of eps)
        y = layer.call(to_device(x, layer.ctx or default_context()))

        def grad(dy, variables=None):  # noqa: ARG001 - TF passes the captured variables by keyword
            layer.zero_grads()
            dx = layer.backward(to_device(dy, layer.ctx))
            return (to_tf(dx), *mirror.grads())

        return to_tf(y), grad

    return op


def mnf_kl_op(layer: Any, mirror: _Mirror):
    """*variables -> kl (scalar) with the gradient of kl_div (mnf_dense.py:104-157, mnf_conv.py:116-180;
    SURVEY.md App. A-4 / A-5), scaled by the upstream gradient (e.g. 1 / len(X_train),
    lenet_mnist.py:61)."""
    tf = _tf()

    @tf.custom_gradient
    def op(*variables):
        mirror.push(variables)
        layer.training = True
        kl = layer.kl_div()

        def grad(dkl, variables=None):  # noqa: ARG001
            layer.zero_grads()
            scale = float(np.asarray(dkl).reshape(-1)[0])  # upstream dLoss/dKL: a host scalar for mnf_kl_backward
            layer.kl_backward(scale)
            return tuple(mirror.grads())

        return tf.reshape(to_tf(kl), []), grad

    return op


def _keras_layer_base():
    return _tf().keras.layers.Layer


def make_keras_layers():
    """Build the Keras classes lazily (they subclass tf.keras.layers.Layer, which needs TF):
    returns (TFMNFDense, TFMNFConv2D) with the reference's constructor signatures
    (mnf_dense.py:22-35, mnf_conv.py:22-39) and methods build / call / kl_div / sample_z."""
    Base = _keras_layer_base()

    class _TFMNFBase(Base):
        _host_cls: Any = None

        def __init__(self, *args: Any, **kwargs: Any) -> None:
            keras_kw = {k: kwargs.pop(k) for k in ("name", "dtype", "trainable") if k in kwargs}
            super().__init__(**keras_kw)
            self.host = self._host_cls(*args, **kwargs)
            self._mirror: _Mirror | None = None

        def build(self, input_shape) -> None:
            self.host.build(tuple(int(d) if d is not None else 1 for d in input_shape))
            self._mirror = _Mirror(self.host, _make_tf_variable)
            for name, v in zip(self._mirror.names, self._mirror.vars):
                setattr(self, name.replace("/", "_"), v)  # mean_W, log_std_W, ... as public attributes
            self._call_op = mnf_call_op(self.host, self._mirror)
            self._kl_op = mnf_kl_op(self.host, self._mirror)
            self._mnf_variables = list(self._mirror.vars)
            super().build(input_shape)

        @property
        def trainable_variables(self):  # what model.trainable_variables collects (lenet_mnist.py:114)
            return list(getattr(self, "_mnf_variables", []))

        def call(self, x):
            return self._call_op(x, *self._mirror.vars)

        def kl_div(self):
            return self._kl_op(*self._mirror.vars)

        def sample_z(self, batch_size: int):
            z, ld = self.host.sample_z(batch_size)
            return to_tf(z), to_tf(ld)

    class TFMNFDense(_TFMNFBase):
        _host_cls = MNFDense

    class TFMNFConv2D(_TFMNFBase):
        _host_cls = MNFConv2D

    return TFMNFDense, TFMNFConv2D
