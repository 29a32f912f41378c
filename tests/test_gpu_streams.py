"""Stream-level concurrency must not change results: z drawn ahead of call() on a side
context (Sequential.presample_z), and kl_div() on a forked context, are bit-identical to the
in-line path (same Philox counters, same kernels)."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lenet(seed=3, **kw):
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFLeNet

    M.set_seed(seed)
    return MNFLeNet(6, 12, 40, 10, max_std=1, flow_h_sizes=[16], std_init=1.0, **kw)


@pytest.mark.parametrize("flow_type", ["rnvp", "iaf"])
@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
def test_presampled_z_is_bit_identical(flow_type, math):
    import tf_mnf_b200 as M

    ctx = M.default_context()
    ctx.set_math(math)
    try:
        x = np.random.RandomState(0).uniform(size=(96, 28, 28, 1)).astype(np.float32)
        outs = []
        for pre in (False, True):
            model = _lenet(flow_type=flow_type)
            model.presample_z = pre
            ys = [model(x).numpy() for _ in range(3)]  # call 0 builds the layers; 1 and 2 use the side stream
            outs.append(ys)
            if pre:
                assert len(model._sides) == 4
        for a, b in zip(*outs):
            np.testing.assert_array_equal(a, b)
        assert not np.array_equal(outs[0][1], outs[0][2])  # fresh noise per call
    finally:
        ctx.set_math("tf32x3")


def test_presample_survives_changed_batch_and_training_flag():
    model = _lenet()
    rng = np.random.RandomState(1)
    xa = rng.uniform(size=(32, 28, 28, 1)).astype(np.float32)
    xb = rng.uniform(size=(48, 28, 28, 1)).astype(np.float32)
    model(xa)
    lyr = model.mnf_layers[2]
    side = lyr.ctx.fork()
    lyr.presample(32, side)          # stale: the next call has 48 rows
    y1 = model(xb).numpy()
    ref = _lenet()
    ref.presample_z = False
    ref(xa)
    y2 = ref(xb).numpy()
    np.testing.assert_array_equal(y1, y2)


def test_kl_on_forked_context_matches_inline():
    x = np.random.RandomState(2).uniform(size=(8, 28, 28, 1)).astype(np.float32)
    m1 = _lenet()
    m1(x)  # variables are created on the first call, from counters set_seed() resets
    m2 = _lenet()
    m2(x)
    side = m1.mnf_layers[0].ctx.fork()
    k1 = m1.kl_div(ctx=side)
    m1.mnf_layers[0].ctx.wait(side)
    side.sync()
    k2 = m2.kl_div()
    assert float(k1.numpy()[0]) == float(k2.numpy()[0])


def test_tile_rows_and_mc_predict_match_host_repeat():
    """mnf_tile_rows == np.repeat(x, n, axis=0); model.mc_predict(x, n) == model(np.repeat(x, n))
    bit for bit (same rows -> same Philox counters)."""
    import ctypes as C

    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L

    ctx = M.default_context()
    rng = np.random.RandomState(4)
    for shape, n in (((5, 7), 3), ((3, 4, 4, 1), 6), ((2, 8), 1)):
        x = rng.standard_normal(shape).astype(np.float32)
        xd = ctx.to_device(x)
        out = ctx.empty((shape[0] * n, *shape[1:]))
        L.check(ctx.lib.mnf_tile_rows(ctx.handle, shape[0], x.size // shape[0], n, C.c_void_p(xd.ptr), C.c_void_p(out.ptr)))
        np.testing.assert_array_equal(out.numpy(), np.repeat(x, n, axis=0))
    imgs = rng.uniform(size=(3, 28, 28, 1)).astype(np.float32)
    m1 = _lenet()
    m1(imgs)
    y1 = m1.mc_predict(imgs, 8).numpy()
    m2 = _lenet()
    m2(imgs)
    y2 = m2(np.repeat(imgs, 8, axis=0)).numpy()
    np.testing.assert_array_equal(y1, y2)
    assert y1.shape == (24, 10) and not np.array_equal(y1[0], y1[1])  # same image, different posterior draws


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
def test_frozen_weights_are_bit_identical_and_unfreeze_sees_updates(math):
    """Context.freeze_weights keeps the packed weight blocks across calls: same outputs bit for
    bit; after unfreezing, a changed parameter is picked up again."""
    import tf_mnf_b200 as M

    ctx = M.default_context()
    ctx.set_math(math)
    try:
        x = np.random.RandomState(7).uniform(size=(40, 28, 28, 1)).astype(np.float32)
        ref = _lenet()
        ys_ref = [ref(x).numpy() for _ in range(3)]
        mdl = _lenet()
        mdl(x)                       # call 0 builds the variables
        mdl.freeze_weights(True)
        ys = [mdl(x).numpy(), mdl(x).numpy()]
        np.testing.assert_array_equal(ys[0], ys_ref[1])
        np.testing.assert_array_equal(ys[1], ys_ref[2])
        mdl.freeze_weights(False)
        w = mdl.mnf_layers[2].mean_W
        w.assign(w.numpy() * 0.5)
        for lyr in mdl.mnf_layers:
            lyr.call_index = 1
        y_new = mdl(x).numpy()
        assert not np.array_equal(y_new, ys_ref[1])
    finally:
        ctx.freeze_weights(False)
        ctx.set_math("tf32x3")


def test_captured_step_replays_match_eager_calls():
    """A forward + kl_div step captured into a CUDA graph: replay i == eager call i bit for bit
    (the device-side replay counter reproduces the eager call indices), and eager calls after the
    graph keep drawing fresh noise."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.graph import CapturedStep

    ctx = M.default_context()
    ctx.set_math("tf32x3")
    try:
        x = np.random.RandomState(9).uniform(size=(64, 28, 28, 1)).astype(np.float32)
        ref = _lenet()
        ref(x)  # build
        ref.freeze_weights(True)
        ys_ref, kls_ref = [], []
        for _ in range(6):
            ys_ref.append(ref(x).numpy())
            kls_ref.append(float(ref.kl_div().numpy()[0]))

        mdl = _lenet()
        mdl(x)
        mdl.freeze_weights(True)
        xd = ctx.to_device(x)
        lanes = [ctx.fork() for _ in range(2)]

        def step():
            for c in lanes:
                c.wait(ctx)
            y = mdl(xd)
            kl = mdl.kl_div(ctx=lanes)
            ctx.wait(lanes[0])
            return y, kl

        # CapturedStep runs the step three times eagerly (warm-up, 2 x pool stocking), then captures call 4
        g = CapturedStep(ctx, step, forks=lanes, layers=mdl.mnf_layers, models=[mdl])
        for i in range(2):
            y, kl = g.launch()
            ctx.sync()
            np.testing.assert_array_equal(y.numpy(), ys_ref[3 + i])
            assert float(kl.numpy()[0]) == kls_ref[3 + i]
        y_after = mdl(xd).numpy()  # eager call after two replays = call index 6
        np.testing.assert_array_equal(y_after, ys_ref[5])
        g.close()
    finally:
        ctx.freeze_weights(False)
        ctx.set_math("tf32x3")


def test_presampled_z_is_discarded_when_parameters_change():
    """Sequential draws the first layer's z for the NEXT call ahead of time; a parameter assigned in
    between must not leave that call with a z drawn from the old q0_mean (it is redrawn in line, with
    the same Philox counters, from the current parameters)."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFLeNet

    x = np.random.RandomState(0).uniform(size=(6, 28, 28, 1)).astype(np.float32)
    outs = []
    for ahead in (True, False):
        M.set_seed(9)
        model = MNFLeNet(4, 6, 16, 10, max_std=1, flow_h_sizes=[8], std_init=1.0)
        model.presample_first = model.presample_z = ahead
        model(x)  # builds the layers (Keras protocol); nothing can be drawn ahead before that
        model(x)  # draws the first layer's z for the next call beside the rest of this one
        l0 = model.mnf_layers[0]
        assert (l0._pre is not None) == ahead
        l0.q0_mean.assign(l0.q0_mean.numpy() + 0.5)
        l0.flow_q.flows[0].net.layers[0].kernel.assign(l0.flow_q.flows[0].net.layers[0].kernel.numpy() * 0.5)
        outs.append(model(x).numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
