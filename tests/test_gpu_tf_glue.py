"""tf_mnf_b200/tf_glue.py (the tf.custom_gradient wrappers of INTEGRATION.md) driven through a
stand-in for the few TensorFlow entry points it uses -- TensorFlow cannot be installed here.  The
stand-in exchanges tensors through REAL DLPack capsules (torch is the producer / consumer), so the
zero-copy path, the parameter mirroring, the forward op and both gradient functions are exercised;
the arithmetic itself is pinned against the oracle elsewhere (test_gpu_backward.py, test_gpu_timed_mode.py)."""
from __future__ import annotations

import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fake_tensorflow():
    import torch
    from torch.utils import dlpack as tdl

    tf = types.ModuleType("tensorflow")

    def custom_gradient(f):
        def wrapper(*args):
            y, grad = f(*args)
            wrapper.tape.append(grad)  # what a GradientTape would record
            return y

        wrapper.tape = []
        return wrapper

    class Layer:
        def __init__(self, name=None, **_):
            self.name, self.built = name, False

        def build(self, input_shape):
            self.built = True

        def __call__(self, x):
            if not self.built:
                self.build(tuple(x.shape))
            return self.call(x)

    tf.custom_gradient = custom_gradient
    tf.Variable = lambda value, name=None, trainable=True: value.clone()
    tf.reshape = lambda t, shape: t.reshape(tuple(shape))
    tf.experimental = types.SimpleNamespace(dlpack=types.SimpleNamespace(
        to_dlpack=lambda t: tdl.to_dlpack(t.contiguous()), from_dlpack=lambda cap: tdl.from_dlpack(cap)))
    tf.keras = types.SimpleNamespace(layers=types.SimpleNamespace(Layer=Layer))
    tf._torch = torch
    return tf


@pytest.fixture()
def fake_tf(monkeypatch):
    tf = _fake_tensorflow()
    monkeypatch.setitem(sys.modules, "tensorflow", tf)
    return tf


@pytest.mark.parametrize("kind", ["dense", "conv"])
def test_custom_gradient_wrappers(fake_tf, kind):
    import tf_mnf_b200 as M
    from tf_mnf_b200 import tf_glue

    torch = fake_tf._torch
    M.set_seed(21)
    TFMNFDense, TFMNFConv2D = tf_glue.make_keras_layers()
    rng = np.random.RandomState(0)
    if kind == "dense":
        layer = TFMNFDense(12, flow_h_sizes=[6], std_init=1.0)
        x = torch.tensor(rng.standard_normal((9, 20)).astype(np.float32), device="cuda")
    else:
        layer = TFMNFConv2D(8, 3, flow_h_sizes=[6], std_init=1.0)
        x = torch.tensor(rng.uniform(size=(5, 10, 10, 4)).astype(np.float32), device="cuda")
    y = layer(x)  # Keras protocol: build on first call, then call
    host = layer.host
    names = list(host.trainable_variables)
    assert [v.shape for v in layer.trainable_variables] == [tuple(host.trainable_variables[n].shape) for n in names]
    assert hasattr(layer, "mean_W") and hasattr(layer, "q0_mean") and hasattr(layer, "flow_q_0_w1")
    # forward: the op's output is the host layer's call() under the same Philox counters
    host.call_index = 0
    y_host = host.call(tf_glue.to_device(x, host.ctx)).numpy()
    np.testing.assert_array_equal(y.cpu().numpy(), y_host)
    # gradient function of call(): (dx, one gradient per mirrored variable), equal to host.backward
    dy = torch.tensor(rng.standard_normal(tuple(y.shape)).astype(np.float32), device="cuda")
    host.zero_grads()
    dx_host = host.backward(tf_glue.to_device(dy, host.ctx)).numpy()
    want = {n: host.grads[n].numpy().copy() for n in names}
    host.call_index = 0
    y2 = layer(x)
    np.testing.assert_array_equal(y2.cpu().numpy(), y_host)
    out = layer._call_op.tape[-1](dy)
    assert len(out) == 1 + len(names)
    np.testing.assert_allclose(out[0].cpu().numpy(), dx_host, rtol=1e-6, atol=1e-7)
    for n, g in zip(names, out[1:]):
        np.testing.assert_allclose(g.cpu().numpy(), want[n], rtol=2e-5, atol=1e-7, err_msg=n)
    assert float(np.abs(want["mean_W"]).max()) > 0 and float(np.abs(want["q0_mean"]).max()) > 0
    # kl_div and its gradient function, scaled by the upstream gradient (1 / len(X_train))
    kl = layer.kl_div()
    assert kl.shape == () and np.isfinite(float(kl))
    gk = layer._kl_op.tape[-1](np.float32(1.0 / 600))
    host.kl_index -= 1
    host.kl_div()
    host.zero_grads()
    host.kl_backward(1.0 / 600)
    for n, g in zip(names, gk):
        np.testing.assert_allclose(g.cpu().numpy(), host.grads[n].numpy(), rtol=2e-5, atol=1e-9, err_msg=n)
    assert float(np.abs(host.grads["r0_mean"].numpy()).max()) > 0
    # an optimiser step on the mirrored variables reaches the kernels on the next call
    with torch.no_grad():
        layer.trainable_variables[names.index("mean_b")].add_(3.0)
    y3 = layer(x)
    assert float((y3 - y2).abs().mean()) > 1e-2  # conv: the bias is scaled by z (mnf_conv.py:193)
    z, ld = layer.sample_z(4)
    assert tuple(z.shape) == (4, host._flow_dim()) and tuple(ld.shape) == (4,)
