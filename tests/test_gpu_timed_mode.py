"""Parity of the TIMED mode: bench.py runs with in-kernel Philox noise, so here nothing is injected.
Each kernel family draws its own eps / eps_z / Bernoulli masks; the test materialises exactly those
draws with mnf_philox_normal / mnf_philox_bernoulli for the same (seed, call index, stream, global
row) and feeds them to the float64 oracle (mnf_dense.py:159-170, mnf_conv.py:182-197,
mnf_dense.py:104-157, mnf_conv.py:116-180).  Also: BASELINE config C3 at its full 10240 rows (two
1280-row slices against the oracle, through model() and mc_predict()), the layer shapes of C5
forward + backward, C4's 4096 -> 4096 layer in the bf16 mode, and Trainer.loss_and_grads against the
oracle's whole train step.  Bar: 1e-5 of the tensor's maximum (north_star), stated per test."""
from __future__ import annotations

import numpy as np
import pytest
from gpu_util import call_noise, close, kl_noise, make_layer, params_from_layer

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _f32(a):
    return None if a is None else np.asarray(a, np.float32)


@pytest.fixture()
def math_ctx(request):
    import tf_mnf_b200 as M

    ctx = M.default_context()
    prev = ctx.math
    yield ctx
    ctx.set_math(prev)


def _layer(rng, wshape, flow, x_shape, **kw):
    P = O.make_layer_params(rng, wshape, flow=flow, std_init=10.0, dtype=np.float64)
    D = wshape[-1] if len(wshape) == 4 else wshape[0]
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(D)
    P["mean_b"][...] = 0.1 * rng.standard_normal(wshape[-1])
    return make_layer(P, x_shape, **kw)


# ---------------------------------------------------------------- (a) every epilogue family, Philox in the kernel
DENSE_CASES = [
    # K, N, B, act, kernel family under tf32x3
    (800, 500, 300, "relu", "paired_umma<128>"),
    (800, 500, 129, None, "paired_umma<128>"),
    (136, 100, 64, None, "paired_umma<64>"),
    (800, 500, 1100, "relu", "paired_f16x3<128>"),   # many rows: packed split-fp16 operands (dense_f16x3.cu)
    (136, 100, 1024, None, "paired_f16x3<128>"),
    (130, 70, 1031, None, "paired_f16x3<128>"),      # ragged: K % 32, N % 4, M % 128 all non-zero
    (500, 10, 257, "softmax", "dense_small"),
    (500, 10, 33, None, "dense_small"),
    (500, 10, 1301, "softmax", "dense_small_rows"),      # many rows: four rows per warp share the weight reads
    (500, 10, 1027, None, "dense_small_rows"),
    (100, 14, 1100, "relu", "dense_small_rows"),
    (1000, 10, 1040, "softmax", "dense_small_rows"),
]


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("flow", ["rnvp", "iaf"])
@pytest.mark.parametrize("K,N,B,act,path", DENSE_CASES)
def test_dense_philox_mode(math_ctx, math, flow, K, N, B, act, path):
    from tf_mnf_b200 import _lib as L

    math_ctx.set_math(math)
    rng = np.random.RandomState(K + N + B)
    lyr = _layer(rng, (K, N), flow, (B, K))
    lyr.row_offset = 4096 + 77  # a shard that does not start at global row 0
    x = rng.standard_normal((B, K))
    for call in range(2):  # the second call must draw fresh noise (call index in the Philox key)
        assert lyr.call_index == call
        got = lyr.call(_f32(x), activation=act).numpy()
        if math == "tf32x3":
            assert L.last_path() == path, L.last_path()
        eps_z, eps_out, bern = call_noise(lyr, B, (B, N), call, lyr.row_offset)
        ref, _ = O.dense_call(params_from_layer(lyr), x.astype(np.float32).astype(np.float64), eps_z, eps_out, 1.0,
                              bern_q=bern)
        if act == "relu" and lyr.last_fused:
            ref = O.relu(ref)
        if act == "softmax" and lyr.last_fused:
            ref = O.softmax(ref)
        close(got, ref, RTOL, f"dense {K}->{N} call {call} ({math}, {flow})")


CONV_CASES = [
    # x shape, filter shape, padding, fuse ReLU+pool, kernel family under tf32x3
    ((37, 12, 12, 20), (5, 5, 20, 50), "VALID", True, "conv_raw_f16"),
    ((130, 12, 12, 20), (5, 5, 20, 50), "VALID", False, "conv_raw_f16"),
    ((701, 12, 12, 20), (5, 5, 20, 50), "VALID", True, "conv_raw_f16x2"),   # enough tiles for every CTA pair (cta_group::2), odd tile count
    ((640, 12, 12, 20), (5, 5, 20, 36), "VALID", True, "conv_raw_f16x2"),
    ((9, 28, 28, 1), (5, 5, 1, 20), "VALID", True, "conv_win_umma"),
    ((5, 28, 28, 1), (5, 5, 1, 20), "VALID", False, "conv_win_umma"),
    ((3, 9, 7, 3), (3, 3, 3, 64), "SAME", False, None),
    ((4, 8, 8, 8), (3, 3, 8, 24), "VALID", True, None),
]


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("flow", ["rnvp", "iaf"])
@pytest.mark.parametrize("shape,wshape,padding,pool,path", CONV_CASES)
def test_conv_philox_mode(math_ctx, math, flow, shape, wshape, padding, pool, path):
    from tf_mnf_b200 import _lib as L

    math_ctx.set_math(math)
    rng = np.random.RandomState(sum(shape) + sum(wshape))
    lyr = _layer(rng, wshape, flow, shape, padding=padding)
    lyr.row_offset = 10240 * 3 + 5
    x = rng.uniform(size=shape)
    B, F = shape[0], wshape[-1]
    x64 = x.astype(np.float32).astype(np.float64)
    oshape = O.conv2d(x64, np.zeros(wshape), 1, padding).shape
    for call in range(2):
        got = lyr.call(_f32(x), fuse_relu_pool=pool).numpy()
        if math == "tf32x3" and path:
            assert L.last_path() == path, L.last_path()
        eps_z, eps_out, bern = call_noise(lyr, B, oshape, call, lyr.row_offset)
        ref, _ = O.conv_call(params_from_layer(lyr), x64, eps_z, eps_out, 1.0, padding=padding, bern_q=bern)
        if pool and lyr.last_fused:
            ref = O.maxpool2x2_same(O.relu(ref))
        close(got, ref, RTOL, f"conv {wshape} call {call} ({math}, {flow})")


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("flow", ["rnvp", "iaf"])
@pytest.mark.parametrize("pool", [True, False])
def test_conv_mc_expand_philox_mode(math_ctx, math, flow, pool):
    """MNFConv2D.call_mc (conv_mc_maps + conv_mc_expand): the rows of the repeated batch."""
    from tf_mnf_b200 import _lib as L

    math_ctx.set_math(math)
    rng = np.random.RandomState(11)
    shape, wshape, n = (3, 28, 28, 1), (5, 5, 1, 20), 19
    lyr = _layer(rng, wshape, flow, shape)
    lyr.row_offset = 999
    x = rng.uniform(size=shape)
    xr = np.repeat(x.astype(np.float32).astype(np.float64), n, axis=0)
    B = xr.shape[0]
    got = lyr.call_mc(_f32(x), n, fuse_relu_pool=pool).numpy()
    assert L.last_path().startswith("conv_mc_expand"), L.last_path()
    eps_z, eps_out, bern = call_noise(lyr, B, (B, 24, 24, 20), 0, lyr.row_offset)
    ref, _ = O.conv_call(params_from_layer(lyr), xr, eps_z, eps_out, 1.0, bern_q=bern)
    if pool:
        ref = O.maxpool2x2_same(O.relu(ref))
    close(got, ref, RTOL, "call_mc")


@pytest.mark.parametrize("flow", ["rnvp", "iaf", "planar"])
@pytest.mark.parametrize("wshape,B", [((800, 500), 200), ((5, 5, 20, 50), 70), ((500, 10), 1)])
def test_sample_z_philox_mode(math_ctx, flow, wshape, B):
    """sample_z (z0 kernel / first flow step + chain) with in-kernel eps_z and masks."""
    math_ctx.set_math("tf32x3")
    rng = np.random.RandomState(B)
    conv = len(wshape) == 4
    lyr = _layer(rng, wshape, flow, (B, 12, 12, wshape[2]) if conv else (B, wshape[0]))
    lyr.row_offset = 123456
    z, ld = lyr.sample_z(B)
    eps_z, _, bern = call_noise(lyr, B, (B, wshape[-1]), 0, lyr.row_offset)
    rz, rld, _ = O.sample_z(params_from_layer(lyr), B, eps_z, bern_q=bern)
    close(z.numpy(), rz, RTOL, "z")
    close(ld.numpy(), rld, RTOL, "log_dets")


@pytest.mark.parametrize("flow", ["rnvp", "iaf", "planar"])
@pytest.mark.parametrize("wshape", [(800, 500), (500, 10), (5, 5, 1, 20), (5, 5, 20, 50)])
def test_kl_philox_mode(math_ctx, flow, wshape):
    """kl_div() with its own draws (single-row chains flow_chain_row1(_cluster) + kl_fwd_kernel)."""
    math_ctx.set_math("tf32x3")
    rng = np.random.RandomState(wshape[0] + len(wshape))
    conv = len(wshape) == 4
    lyr = _layer(rng, wshape, flow, (1, 12, 12, wshape[2]) if conv else (1, wshape[0]))
    P = params_from_layer(lyr)
    for k in range(2):
        got = lyr.kl_div().numpy()
        eps_z, eps_a, eps_b, bq, br = kl_noise(lyr, k)
        if conv:
            ref, _ = O.conv_kl(P, eps_z, eps_a, eps_b, bern_q=bq, bern_r=br)
        else:
            ref, _ = O.dense_kl(P, eps_z, eps_a, bern_q=bq, bern_r=br)
        close(got, np.asarray(ref).reshape(1), RTOL, f"kl call {k}")


# ---------------------------------------------------------------- (b) C3 at full size against the oracle
def _lenet_noise(model, rows, lo, call):
    outs = [(rows, 24, 24, 20), (rows, 8, 8, 50), (rows, 500), (rows, 10)]
    noise, bern = [], []
    for lyr, o in zip(model.mnf_layers, outs):
        ez, eo, bq = call_noise(lyr, rows, o, call, lo)
        noise.append((ez, eo))
        bern.append(bq)
    return noise, bern


@pytest.mark.parametrize("flow", ["rnvp"])
def test_c3_full_size_slices_vs_oracle(math_ctx, flow):
    """BASELINE configs[2]: 1024 samples x 10 images = 10240 rows, the timed configuration of bench.py
    (tf32x3, in-kernel Philox, frozen weights).  Rows [0,1280) and [8960,10240) of the probabilities
    of model(x) (call 0) and of mc_predict(images, 1024) (call 1) against the float64 oracle, and the
    model's kl_div()."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFLeNet

    math_ctx.set_math("tf32x3")
    M.set_seed(0)
    model = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=10.0, flow_type=flow)
    rng = np.random.RandomState(0)
    imgs = rng.uniform(size=(10, 28, 28, 1)).astype(np.float32)
    x = np.repeat(imgs, 1024, axis=0)
    xd = math_ctx.to_device(x)
    full0 = model(xd).numpy()
    model.freeze_weights(True)
    try:
        full1 = model.mc_predict(imgs, 1024).numpy()
        kl = model.kl_div().numpy()
    finally:
        model.freeze_weights(False)
    assert full0.shape == full1.shape == (10240, 10)
    Ps = [params_from_layer(l) for l in model.mnf_layers]
    for call, full in ((0, full0), (1, full1)):
        for lo, hi in ((0, 1280), (8960, 10240)):
            noise, bern = _lenet_noise(model, hi - lo, lo, call)
            ref = O.lenet_forward(Ps, x[lo:hi].astype(np.float64), noise, max_std=1.0, bern=bern)
            close(full[lo:hi], ref, RTOL, f"C3 probabilities rows [{lo},{hi}) call {call}")
    kn, kb = [], []
    for lyr in model.mnf_layers:
        ez, ea, eb, bq, br = kl_noise(lyr, 0)
        kn.append((ez, ea, eb) if lyr._conv else (ez, ea))
        kb.append((bq, br))
    close(kl, np.asarray(O.lenet_kl(Ps, kn, bern=kb)).reshape(1), RTOL, "C3 kl")


# ---------------------------------------------------------------- (c) C5 layer shapes, forward + backward
C5_LAYERS = [((4, 32, 32, 64), (3, 3, 64, 64)), ((4, 16, 16, 128), (3, 3, 128, 128)), ((3, 8, 8, 256), (3, 3, 256, 256)),
             ((4, 32, 32, 3), (3, 3, 3, 64))]


def _grad_names():
    return ("mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var")


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("shape,wshape", C5_LAYERS)
def test_c5_conv_layer_shapes(math_ctx, math, shape, wshape):
    """The 3x3 SAME convolutions of the MNF VGG-style net: K = 576 / 1152 / 2304."""
    from test_gpu_backward import RTOL_G, flow_grad_map, perturb_biases

    math_ctx.set_math(math)
    rng = np.random.RandomState(sum(wshape))
    P = O.make_layer_params(rng, wshape, flow="iaf", std_init=2.0, dtype=np.float64)
    F = wshape[-1]
    P["q0_log_var"][...] = -2 + 0.5 * rng.standard_normal(F)
    P["mean_b"][...] = 0.1 * rng.standard_normal(F)
    perturb_biases(P["flow_q"], rng)
    x = rng.standard_normal(shape)
    B = shape[0]
    oshape = (B, shape[1], shape[2], F)
    eps_z, eps, dy = rng.standard_normal((B, F)), rng.standard_normal(oshape), rng.standard_normal(oshape)
    ref, c = O.conv_call(P, x, eps_z, eps, 1.0, stride=1, padding="SAME")
    ref_dx, ref_g = O.conv_call_backward(P, c, dy)
    lyr = make_layer(P, shape, padding="SAME")
    lyr.training = True
    y = lyr(_f32(x), noise=dict(eps_z=_f32(eps_z), eps_out=_f32(eps)))
    close(y.numpy(), ref, RTOL, "y")
    lyr.zero_grads()
    dx = lyr.backward(_f32(dy))
    close(dx.numpy(), ref_dx, RTOL, "dx")
    # Written bound for the one case above 1e-5: with K = kh kw C > 1024 terms per output the tensor core's
    # fp32 accumulation (no round-to-nearest between MMAs) leaves the forward's saved mean
    # at ~6e-6 of its maximum (K = 1152; the library keeps K <= 1280 on tcgen05).  dz = sum_hw g * mean
    # inherits that relative error, and the q-flow gradients are linear in dz: measured 1.2e-5, bound 2e-5.
    K = int(np.prod(wshape[:3]))
    flow_tol = 2e-5 if (math == "tf32x3" and K > 1024) else RTOL_G
    for name in _grad_names():
        close(lyr.grads[name].numpy(), ref_g[name], flow_tol if name.startswith("q0") else RTOL_G, name)
    for k, g in enumerate(ref_g["flow_q"]):
        for name, r in flow_grad_map("iaf", g).items():
            close(lyr.grads[f"flow_q/{k}/{name}"].numpy().reshape(np.shape(r)), r, flow_tol, f"flow_q/{k}/{name}")


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("K,N,B", [(4096, 512, 96), (512, 10, 96)])
def test_c5_dense_layer_shapes(math_ctx, math, K, N, B):
    from test_gpu_backward import RTOL_G, perturb_biases

    math_ctx.set_math(math)
    rng = np.random.RandomState(K + N)
    P = O.make_layer_params(rng, (K, N), flow="iaf", std_init=2.0, dtype=np.float64)
    P["q0_log_var"][...] = -2 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    perturb_biases(P["flow_q"], rng)
    for st in P["flow_q"]:
        st["W"][1] *= 0.1  # two un-masked 4096-wide IAF steps: keep exp(s) inside the fp32 range
    x = rng.standard_normal((B, K))
    eps_z, eps, dy = rng.standard_normal((B, K)), rng.standard_normal((B, N)), rng.standard_normal((B, N))
    ref, c = O.dense_call(P, x, eps_z, eps, 1.0)
    ref_dx, ref_g = O.dense_call_backward(P, c, dy)
    lyr = make_layer(P, x.shape)
    lyr.training = True
    y = lyr(_f32(x), noise=dict(eps_z=_f32(eps_z), eps_out=_f32(eps)))
    close(y.numpy(), ref, RTOL, "y")
    lyr.zero_grads()
    dx = lyr.backward(_f32(dy))
    close(dx.numpy(), ref_dx, RTOL, "dx")
    for name in _grad_names():
        close(lyr.grads[name].numpy(), ref_g[name], RTOL_G, name)


# ---------------------------------------------------------------- (d) C4: 4096 -> 4096, bf16 mode
def test_c4_wide_dense_bf16_full_width(math_ctx):
    """BASELINE configs[3]: MNFDense 4096 -> 4096, planar flows, bf16 operands / fp32 accumulate.
    Tolerance of the bf16 mode: rel-L2 <= 1e-2 against the float64 oracle and per
    element |err| <= 2^-7 (|x.z| . |W| + sd |eps|) -- term-by-term rounding of both operands to 8
    significant bits (2^-9 each, two operands, worst case all terms aligned: 2^-8; bound used 2^-7)."""
    from tf_mnf_b200 import _lib as L

    math_ctx.set_math("bf16")
    rng = np.random.RandomState(4)
    K = N = 4096
    B = 512
    lyr = _layer(rng, (K, N), "planar", (B, K))
    x = rng.standard_normal((B, K)).astype(np.float32)
    got = lyr.call(x).numpy().astype(np.float64)
    assert L.last_path().startswith("paired_bf16"), L.last_path()
    P = params_from_layer(lyr)
    eps_z, eps_out, _ = call_noise(lyr, B, (B, N), 0, 0)
    x64 = x.astype(np.float64)
    ref, c = O.dense_call(P, x64, eps_z, eps_out, 1.0)
    rel_l2 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel_l2 <= 1e-2, rel_l2
    z, _, _ = O.sample_z(P, B, eps_z)
    sd = np.sqrt((x64 * x64) @ np.minimum(np.exp(P["log_std_W"]), 1.0) ** 2 + np.minimum(np.exp(P["log_var_b"]), 1.0))
    bound = 2.0 ** -7 * (np.abs(x64 * z) @ np.abs(P["mean_W"]) + sd * np.abs(eps_out)) + 1e-6
    assert np.all(np.abs(got - ref) <= bound), float(np.max(np.abs(got - ref) / bound))


# ---------------------------------------------------------------- (e) the whole train step against the oracle
def _pool_fwd(h):
    B, H, W, Cc = h.shape
    r = np.maximum(h, 0).reshape(B, H // 2, 2, W // 2, 2, Cc)
    p = r.max(axis=(2, 4))
    return p, (r, p, h.shape)


def _pool_bwd(g, c):
    r, p, shp = c
    mask = (r == p[:, :, None, :, None, :]) & (r > 0)
    first = np.cumsum(np.cumsum(mask, axis=2), axis=4) == 1  # not used: ties have measure zero with continuous noise
    del first
    return (mask * g[:, :, None, :, None, :]).reshape(shp)


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("flow", ["rnvp", "iaf"])
def test_trainer_loss_and_grads_vs_oracle_train_step(math_ctx, math, flow):
    """BASELINE configs[0] (MNF-LeNet5 train step; lenet_mnist.py:53-62,108-114): loss, KL and EVERY
    gradient tensor Trainer.loss_and_grads leaves in layer.grads against the oracle's forward +
    analytic backward on the same (Philox-materialised) noise."""
    import tf_mnf_b200 as M
    from test_gpu_backward import RTOL_G, flow_grad_map
    from tf_mnf_b200.models import MNFLeNet
    from tf_mnf_b200.train import Trainer

    math_ctx.set_math(math)
    M.set_seed(2)
    B, n_train = 32, 600.0
    model = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=2.0, flow_type=flow)
    rng = np.random.RandomState(3)
    x = rng.uniform(size=(B, 28, 28, 1)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=B)]
    tr = Trainer(model, loss="xent", kl_weight=1.0 / n_train)
    tr._ensure_flat(math_ctx.to_device(x))
    calls = [(l.call_index, l.kl_index) for l in model.mnf_layers]
    loss, kl = tr.loss_and_grads(x, y)
    loss, kl = float(loss.numpy()[0]), float(kl.numpy()[0])

    Ps = [params_from_layer(l) for l in model.mnf_layers]
    outs = [(B, 24, 24, 20), (B, 8, 8, 50), (B, 500), (B, 10)]
    nb = [call_noise(l, B, o, ci, 0) for l, o, (ci, _) in zip(model.mnf_layers, outs, calls)]
    x64 = x.astype(np.float64)
    h, c1 = O.conv_call(Ps[0], x64, nb[0][0], nb[0][1], 1.0, bern_q=nb[0][2])
    h, p1 = _pool_fwd(h)
    h, c2 = O.conv_call(Ps[1], h, nb[1][0], nb[1][1], 1.0, bern_q=nb[1][2])
    h, p2 = _pool_fwd(h)
    fshape = h.shape
    h, c3 = O.dense_call(Ps[2], h.reshape(B, -1), nb[2][0], nb[2][1], 1.0, bern_q=nb[2][2])
    a3 = h
    h, c4 = O.dense_call(Ps[3], np.maximum(h, 0), nb[3][0], nb[3][1], 1.0, bern_q=nb[3][2])
    probs = O.softmax(h)
    ref_loss = float(-np.mean(np.log(probs[y > 0])))
    g = (probs - y) / B
    g, g4 = O.dense_call_backward(Ps[3], c4, g)
    g, g3 = O.dense_call_backward(Ps[2], c3, g * (a3 > 0))
    g, g2 = O.conv_call_backward(Ps[1], c2, _pool_bwd(g.reshape(fshape), p2))
    _, g1 = O.conv_call_backward(Ps[0], c1, _pool_bwd(g, p1))
    like = [g1, g2, g3, g4]
    ref_kl, klg = 0.0, []
    for lyr, P, (_, ki) in zip(model.mnf_layers, Ps, calls):
        ez, ea, eb, bq, br = kl_noise(lyr, ki)
        if lyr._conv:
            k, c = O.conv_kl(P, ez, ea, eb, bern_q=bq, bern_r=br)
            klg.append(O.conv_kl_backward(P, c, G=1.0 / n_train))
        else:
            k, c = O.dense_kl(P, ez, ea, bern_q=bq, bern_r=br)
            klg.append(O.dense_kl_backward(P, c, G=1.0 / n_train))
        ref_kl += float(k)
    assert abs(loss - ref_loss) <= RTOL * abs(ref_loss), (loss, ref_loss)
    assert abs(kl - ref_kl) <= RTOL * abs(ref_kl), (kl, ref_kl)

    def total(li, name):
        t = None
        for src in (like[li], klg[li]):
            if name in src and src[name] is not None:
                t = np.asarray(src[name], np.float64) if t is None else t + src[name]
        return t

    checked = 0
    for li, lyr in enumerate(model.mnf_layers):
        for name in ("mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var", "r0_mean", "r0_log_var",
                     "r0_apvar"):
            ref = total(li, name)
            assert ref is not None, name
            close(lyr.grads[name].numpy().reshape(ref.shape), ref, RTOL_G, f"layer {li} {name}")
            checked += 1
        for tag in ("flow_q", "flow_r"):
            for k in range(2):
                parts = [flow_grad_map(flow, s[tag][k]) for s in (like[li], klg[li]) if tag in s and s[tag]]
                for name in parts[0]:
                    ref = sum(np.asarray(p[name], np.float64) for p in parts)
                    close(lyr.grads[f"{tag}/{k}/{name}"].numpy().reshape(ref.shape), ref, RTOL_G,
                          f"layer {li} {tag}/{k}/{name}")
                    checked += 1
    assert checked >= 4 * 9 + 4 * 2 * 2 * 4
