"""Monte-Carlo predictive with a shared first layer: MNFConv2D.call_mc /
Sequential.mc_predict evaluate the first layer's two convolutions once per distinct image and
draw only the noise per posterior sample.  mnf_conv.py:182-197 applies z and eps after the
contractions, so the result is that of call() on np.repeat(x, n, axis=0)."""
from __future__ import annotations

import numpy as np
import pytest
from gpu_util import close, make_layer

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _f32(a):
    return np.asarray(a, np.float32)


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
@pytest.mark.parametrize("flow", ["iaf", "rnvp"])
@pytest.mark.parametrize("shape,wshape,n,pool,padding", [
    ((3, 28, 28, 1), (5, 5, 1, 20), 7, True, "VALID"), ((2, 12, 12, 20), (5, 5, 20, 48), 5, True, "VALID"),
    ((4, 9, 11, 3), (3, 3, 3, 8), 3, False, "VALID"), ((1, 10, 10, 2), (3, 3, 2, 4), 33, True, "VALID"),
    ((2, 8, 6, 3), (3, 3, 3, 12), 4, True, "SAME"), ((3, 7, 9, 2), (2, 3, 2, 4), 2, False, "SAME")])
def test_call_mc_matches_oracle_on_repeated_batch(math, flow, shape, wshape, n, pool, padding):
    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L

    ctx = M.default_context()
    ctx.set_math(math)
    try:
        rng = np.random.RandomState(sum(shape) + n)
        P = O.make_layer_params(rng, wshape, flow=flow, std_init=10.0, dtype=np.float64)
        F = wshape[-1]
        P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(F)
        P["mean_b"][...] = 0.1 * rng.standard_normal(F)
        x = rng.uniform(size=shape)
        xr = np.repeat(x, n, axis=0)
        B = xr.shape[0]
        y0 = O.conv2d(xr, P["mean_W"], 1, padding)
        eps_z, eps = rng.standard_normal((B, F)), rng.standard_normal(y0.shape)
        bern = [(rng.uniform(size=(B, F)) <= 0.5).astype(np.float64) for _ in range(2)]
        ref, _ = O.conv_call(P, xr, eps_z, eps, 1.0, padding=padding, bern_q=bern)
        if pool:
            ref = O.maxpool2x2_same(O.relu(ref))
        lyr = make_layer(P, x.shape, padding=padding)
        inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern] if flow == "rnvp" else None)
        assert lyr.mc_supported(x.shape, pool)
        got = lyr.call_mc(_f32(x), n, noise=inj, fuse_relu_pool=pool)
        assert L.last_path().startswith("conv_mc_expand"), L.last_path()
        close(got.numpy(), ref, RTOL, "call_mc")
    finally:
        ctx.set_math("tf32x3")


def test_mc_predict_shared_equals_tiled_batch():
    """Same seeds, same Philox counters: the shared-first-layer path and the tiled-batch path of
    mc_predict agree to fp32 accuracy (the per-image maps come from a different kernel than the
    full-batch convolution, so not bit for bit), for any row offset (multi-GPU shards)."""
    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.models import MNFLeNet

    ctx = M.default_context()
    ctx.set_math("tf32x3")
    try:
        x = np.random.RandomState(0).uniform(size=(3, 28, 28, 1)).astype(np.float32)
        outs = {}
        for share in (True, False):
            M.set_seed(4)
            model = MNFLeNet(flow_type="rnvp", flow_h_sizes=[50], std_init=10.0, max_std=1)
            model.mc_share_first = share
            model.presample_first = False
            model.set_row_offset(1000)
            a = model.mc_predict(x, 40).numpy()
            b = model.mc_predict(x, 40).numpy()  # second call: fresh noise
            outs[share] = (a, b)
            assert not np.array_equal(a, b)
        for k in (0, 1):
            close(outs[True][k], outs[False][k], RTOL, f"mc_predict call {k}")
            np.testing.assert_allclose(outs[True][k].sum(1), 1.0, rtol=1e-5)
        # and it equals the model called on the repeated batch
        M.set_seed(4)
        model = MNFLeNet(flow_type="rnvp", flow_h_sizes=[50], std_init=10.0, max_std=1)
        model.presample_first = False
        model.set_row_offset(1000)
        close(model(np.repeat(x, 40, axis=0)).numpy(), outs[True][0], RTOL, "model(np.repeat)")
    finally:
        ctx.set_math("tf32x3")


def test_mc_predict_falls_back_when_unsupported():
    """A first layer the expansion kernel does not cover (F % 4 != 0) or a dense first layer takes
    the tiled-batch path with the same result shape."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.layers import MNFConv2D
    from tf_mnf_b200.models import MNFFeedForward, Sequential

    M.set_seed(1)
    x = np.random.RandomState(1).uniform(size=(2, 8, 8, 3)).astype(np.float32)
    m1 = Sequential([MNFConv2D(6, 3, flow_h_sizes=[8])])
    assert m1.mc_predict(x, 5).shape == (10, 6, 6, 6)
    m2 = MNFFeedForward(layer_sizes=(8, 4), flow_h_sizes=[8])
    assert m2.mc_predict(np.ones((3, 12), np.float32), 4).shape == (12, 4)


def test_mc_predict_second_conv_generates_first_layer_rows():
    """MNF-LeNet under mc_predict with enough rows for the CTA-pair conv kernel: the second MNFConv2D's plane
    producers draw the first layer's noise themselves (mnf_conv2d_forward_mc, conv_raw_f16x2<gen>) instead of
    reading the expanded, pooled activation.  Same Philox counters and the same arithmetic per element, so the
    class probabilities equal those of the two-kernel path bit for bit; fresh noise on the second call."""
    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.models import MNFLeNet

    ctx = M.default_context()
    ctx.set_math("tf32x3")
    x = np.random.RandomState(5).uniform(size=(4, 28, 28, 1)).astype(np.float32)
    n = 301  # 1204 rows: 602 two-image tiles >= 2 per SM, ragged last pass
    outs, paths = {}, {}
    for fuse in (True, False):
        M.set_seed(9)
        model = MNFLeNet(flow_type="rnvp", flow_h_sizes=[50], std_init=10.0, max_std=1)
        model.mc_fuse_second = fuse
        model.presample_first = False
        model.set_row_offset(77)
        seen = []
        conv2 = model.mnf_layers[1]
        orig = conv2.call

        def spy(*a, _o=orig, **k):
            y = _o(*a, **k)
            seen.append(L.last_path())
            return y

        conv2.call = spy
        a = model.mc_predict(x, n).numpy()
        b = model.mc_predict(x, n).numpy()
        outs[fuse], paths[fuse] = (a, b), seen
        assert not np.array_equal(a, b)
    assert paths[True] == ["conv_raw_f16x2<gen>"] * 2, paths
    assert paths[False] == ["conv_raw_f16x2"] * 2, paths
    for k in (0, 1):
        np.testing.assert_array_equal(outs[True][k], outs[False][k])
        np.testing.assert_allclose(outs[True][k].sum(1), 1.0, rtol=1e-5)
