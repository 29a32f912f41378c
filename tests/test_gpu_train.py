"""Training-step closure on the GPU: model-level backward (MNF layers + stock layers), the
loss of lenet_mnist.py:53-62, and Adam.  Layer-level gradients are pinned against the oracle in
test_gpu_backward.py; here the chain through the whole model is checked by finite differences
of the loss under identical Philox noise, and a few Adam steps must reduce the loss."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model_and_batch(seed=0):
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFLeNet

    M.set_seed(seed)
    model = MNFLeNet(4, 6, 16, 10, max_std=1, flow_h_sizes=[8], std_init=1.0)
    rng = np.random.RandomState(1)
    x = rng.uniform(size=(16, 28, 28, 1)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=16)]
    return model, x, y


def _reset_calls(model):
    for lyr in model.mnf_layers:
        lyr.call_index = lyr.kl_index = 0


def test_model_backward_matches_finite_differences():
    from tf_mnf_b200.train import Trainer

    model, x, y = _model_and_batch()
    tr = Trainer(model, loss="xent", kl_weight=1.0 / 600)
    loss0, kl0 = tr.loss_and_grads(x, y)
    base = float(loss0.numpy()[0]) + float(kl0.numpy()[0]) / 600
    checked = 0
    for li, name, idx in [(0, "mean_W", 3), (1, "mean_W", 17), (2, "mean_W", 40), (3, "mean_b", 2), (2, "q0_mean", 5)]:
        lyr = model.mnf_layers[li]
        g = lyr.grads[name].numpy().ravel()[idx]
        p = getattr(lyr, name)
        host = p.numpy()
        eps = 2e-2
        vals = []
        for sgn in (+1, -1):
            h2 = host.copy().ravel()
            h2[idx] += sgn * eps
            p.assign(h2.reshape(host.shape))
            _reset_calls(model)
            l, k = tr.loss_and_grads(x, y)
            vals.append(float(l.numpy()[0]) + float(k.numpy()[0]) / 600)
        p.assign(host)
        num = (vals[0] - vals[1]) / (2 * eps)
        assert abs(num - g) <= 0.08 * max(abs(num), abs(g)) + 2e-4, (li, name, num, g)
        checked += 1
    assert checked == 5 and np.isfinite(base)


def test_adam_steps_reduce_the_loss():
    from tf_mnf_b200.train import Trainer

    model, x, y = _model_and_batch(seed=3)
    tr = Trainer(model, loss="xent", kl_weight=1e-6, lr=5e-3)
    losses = []
    for _ in range(25):
        loss, _ = tr.step(x, y)
        losses.append(float(loss.numpy()[0]))
    assert np.all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < 0.95 * np.mean(losses[:5]), losses


def test_feed_forward_mse_step():
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFFeedForward
    from tf_mnf_b200.train import Trainer

    M.set_seed(5)
    model = MNFFeedForward(layer_sizes=(32, 16, 4), flow_h_sizes=[8], std_init=1e-2)
    rng = np.random.RandomState(2)
    x = rng.standard_normal((64, 24)).astype(np.float32)
    t = (x[:, :4] * 0.5).astype(np.float32)
    tr = Trainer(model, loss="mse", kl_weight=1e-6, lr=1e-2)
    losses = [float(tr.step(x, t)[0].numpy()[0]) for _ in range(30)]
    assert np.all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])


def test_trainer_flat_state_keeps_the_model_usable():
    """Trainer moves the variables into one flat buffer (re-pointing the live objects) and makes the
    gradients views of another: layers, flows and inference calls must keep working on them, a
    second trainer may take the model over, and the KL branch's forked stream must leave the same
    gradients as a serial evaluation (checked through the parameter update being deterministic)."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.train import Trainer

    def run(seed):
        model, x, y = _model_and_batch(seed=seed)
        tr = Trainer(model, loss="xent", kl_weight=1e-3, lr=1e-3)
        for _ in range(3):
            tr.step(x, y)
        return model, tr, x, y

    model, tr, x, y = run(7)
    lo, hi = tr.flat_p.ptr, tr.flat_p.ptr + tr.flat_p.nbytes
    for lyr in model.mnf_layers:
        for name, v in lyr.trainable_variables.items():
            assert lo <= v.ptr < hi, (lyr.name, name)
            g = lyr.grads[name]
            assert tr.flat_g.ptr <= g.ptr < tr.flat_g.ptr + tr.flat_g.nbytes and g.shape == v.shape
    p = model.mc_predict(x[:2], 5).numpy()  # inference on the re-pointed variables (shared first layer)
    assert p.shape == (10, 10) and np.all(np.isfinite(p))
    np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-5)
    # the same three steps from the same seed give the same parameters, up to the order of the
    # fp32 atomic adds in the weight-gradient kernels
    model2, tr2, _, _ = run(7)
    a, b = tr.flat_p.numpy(), tr2.flat_p.numpy()
    assert np.max(np.abs(a - b)) <= 1e-5 * max(np.max(np.abs(a)), 1.0)
    # a second trainer takes the model over
    tr3 = Trainer(model, loss="xent", kl_weight=1e-3, lr=1e-3)
    loss, kl = tr3.step(x, y)
    assert np.isfinite(float(loss.numpy()[0])) and np.isfinite(float(kl.numpy()[0]))
    M.default_context().sync()
