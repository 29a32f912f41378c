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


@pytest.mark.parametrize("batch_stats", [True, False])
def test_batchnorm_layer_matches_oracle(batch_stats):
    """BatchNormalization (mnf_feed_forward.py:20) forward, moving-average update and backward against
    the oracle: fit-mode batch statistics and the tape loop's moving statistics; 1e-5 of the maximum."""
    import tf_mnf_b200 as M
    from gpu_util import close
    from tf_mnf_b200.models.stock import BatchNormalization

    from oracle import mnf_oracle as O

    ctx = M.default_context()
    rng = np.random.RandomState(8)
    B, N = 1024, 100
    x = (3 + 2 * rng.standard_normal((B, N))).astype(np.float32)
    dy = rng.standard_normal((B, N)).astype(np.float32)
    bn = BatchNormalization(batch_stats=batch_stats)
    bn.build((B, N), ctx)
    g, b = 1 + 0.3 * rng.standard_normal(N), 0.2 * rng.standard_normal(N)
    mm, mv = 0.1 * rng.standard_normal(N), 1 + 0.2 * rng.uniform(size=N)
    for arr, v in ((bn.gamma, g), (bn.beta, b), (bn.moving_mean, mm), (bn.moving_variance, mv)):
        arr.assign(v.astype(np.float32))
    g, b, mm, mv = (a.numpy().astype(np.float64) for a in (bn.gamma, bn.beta, bn.moving_mean, bn.moving_variance))
    x64 = x.astype(np.float64)
    # inference call: moving statistics, nothing updated
    y_inf = bn(ctx.to_device(x)).numpy()
    ref_inf, _, _ = O.batchnorm_forward(x64, g, b, mm, mv, training=False)
    close(y_inf, ref_inf, 1e-5, "inference")
    bn.training = True
    y = bn(ctx.to_device(x)).numpy()
    ref, c, (nm, nv) = O.batchnorm_forward(x64, g, b, mm, mv, training=batch_stats)
    close(y, ref, 1e-5, "training-call y")
    close(bn.moving_mean.numpy(), nm, 1e-6, "moving mean")
    close(bn.moving_variance.numpy(), nv, 1e-6, "moving variance")
    bn.zero_grads()
    dx = bn.backward(ctx.to_device(dy)).numpy()
    rdx, rdg, rdb = O.batchnorm_backward(dy.astype(np.float64), c)
    close(dx, rdx, 1e-5, "dx")
    close(bn.grads["gamma"].numpy(), rdg, 1e-5, "dgamma")
    close(bn.grads["beta"].numpy(), rdb, 1e-5, "dbeta")


def test_feed_forward_trains_batchnorm_parameters():
    """BASELINE configs[1] (MNF-MLP regressor): gamma / beta of the BatchNormalization layers are
    trainable variables of the step (they live in the Trainer's flat buffer and move), the moving
    statistics follow the batches, and a prediction afterwards uses them."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFFeedForward
    from tf_mnf_b200.train import Trainer

    M.set_seed(6)
    model = MNFFeedForward(layer_sizes=(32, 16, 4), flow_h_sizes=[8], std_init=1e-2)
    rng = np.random.RandomState(2)
    x = (1.5 + rng.standard_normal((64, 24))).astype(np.float32)
    t = (x[:, :4] * 0.5).astype(np.float32)
    tr = Trainer(model, loss="mse", kl_weight=1e-6, lr=1e-2)
    for _ in range(5):
        tr.step(x, t)
    bns = [l for l in model.layers if type(l).__name__ == "BatchNormalization"]
    assert len(bns) == 2
    lo, hi = tr.flat_p.ptr, tr.flat_p.ptr + tr.flat_p.nbytes
    for bn in bns:
        assert lo <= bn.gamma.ptr < hi and lo <= bn.beta.ptr < hi
        assert np.max(np.abs(bn.gamma.numpy() - 1)) > 1e-4 and np.max(np.abs(bn.beta.numpy())) > 1e-4
        assert np.max(np.abs(bn.moving_mean.numpy())) > 1e-4  # relu outputs have a positive mean
    p = model(x).numpy()
    assert p.shape == (64, 4) and np.all(np.isfinite(p))


@pytest.mark.parametrize("which", ["lenet", "mlp"])
def test_captured_train_step_equals_eager_steps(which):
    """Trainer.capture: the whole step (forward, loss, backward, KL branch, Adam) replayed as ONE CUDA graph gives
    the parameters, losses and KL values of the same number of eager Trainer.step calls -- fresh Philox noise per
    replay (device-side replay counter), Adam's bias correction from the device-side step count, unfrozen weights
    re-packed inside the graph -- although the capture itself ran warm-up steps (on a snapshot)."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFFeedForward
    from tf_mnf_b200.train import Trainer

    def make():
        if which == "lenet":
            model, x, y = _model_and_batch(seed=3)
            return model, x, y, dict(loss="xent", kl_weight=1.0 / 600, lr=1e-2)
        M.set_seed(3)
        rng = np.random.RandomState(2)
        model = MNFFeedForward(layer_sizes=(32, 16, 1), flow_h_sizes=[8], std_init=1.0)
        return model, rng.standard_normal((64, 24)).astype(np.float32), rng.standard_normal((64, 1)).astype(np.float32), \
            dict(loss="mse", kl_weight=1.0 / 128, lr=1e-2)

    n = 4
    model, x, y, kw = make()
    tr = Trainer(model, **kw)
    eager = []
    for _ in range(n):
        l, k = tr.step(x, y)
        eager.append((float(l.numpy()[0]), float(k.numpy()[0])))
    p_eager = tr.flat_p.numpy().copy()

    model2, x2, y2, kw2 = make()
    tr2 = Trainer(model2, **kw2)
    cap = tr2.capture(x2, y2)
    graph = []
    for _ in range(n):
        l, k = cap.step()
        graph.append((float(l.numpy()[0]), float(k.numpy()[0])))
    p_graph = tr2.flat_p.numpy().copy()
    assert tr2.t == n
    np.testing.assert_allclose(np.array(graph), np.array(eager), rtol=2e-5, atol=1e-6)
    scale = np.abs(p_eager).max()
    assert np.abs(p_graph - p_eager).max() <= 2e-5 * scale, np.abs(p_graph - p_eager).max() / scale
    # and an eager step afterwards continues the same sequence (counters, step count and noise indices in sync)
    l_e, _ = tr.step(x, y)
    l_g, _ = tr2.step(x2, y2)
    np.testing.assert_allclose(l_g.numpy(), l_e.numpy(), rtol=2e-5, atol=1e-6)
