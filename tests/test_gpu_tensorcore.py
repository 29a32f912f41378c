"""The tcgen05 (3xTF32) tensor-core path of the paired mean/variance contraction must meet the
same fp32 parity bar (1e-5 relative) as the FFMA path, against the float64 oracle."""
from __future__ import annotations

import numpy as np
import pytest
from gpu_util import close, make_layer

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture()
def tc_ctx():
    import tf_mnf_b200 as M

    ctx = M.default_context()
    ctx.set_math("tf32x3")
    yield ctx
    ctx.set_math("tf32x3")


def _f32(a):
    return np.asarray(a, np.float32)


@pytest.mark.parametrize("K,N,B,use_z", [(800, 500, 300, True), (500, 10, 257, True), (136, 100, 64, True),
                                         (25, 3, 1, True), (64, 64, 128, False), (130, 70, 129, True),
                                         (500, 10, 1029, True), (60, 3, 1100, False), (96, 14, 1024, True),
                                         (1000, 10, 1040, True)])
def test_dense_tensorcore(tc_ctx, K, N, B, use_z):
    rng = np.random.RandomState(K * 3 + N)
    P = O.make_layer_params(rng, (K, N), flow="iaf", std_init=10.0, dtype=np.float64)
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    x = rng.standard_normal((B, K))
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    ref, _ = O.dense_call(P, x, eps_z, eps, 1.0, use_z=use_z)
    lyr = make_layer(P, x.shape, use_z=use_z)
    inj = dict(eps_out=_f32(eps))
    want_path = "paired_umma<128>" if N >= 256 else "paired_umma<64>"
    if N <= 16 and K % 4 == 0 and K * ((N + 3) // 4 * 4) <= 12288:
        # FFMA kernels of the classifier head: warp per row; many rows: four rows per warp
        want_path = "dense_small" if B < 1024 else "dense_small_rows"
    if use_z:
        inj["eps_z"] = _f32(eps_z)
    close(lyr(_f32(x), noise=inj).numpy(), ref, RTOL, "dense call (tf32x3)")
    from tf_mnf_b200 import _lib as L

    assert L.last_path() == want_path, L.last_path()


@pytest.mark.parametrize("shape,wshape,stride,padding", [
    ((5, 12, 12, 20), (5, 5, 20, 50), 1, "VALID"), ((3, 9, 7, 3), (3, 3, 3, 64), 1, "SAME"),
    ((2, 11, 10, 4), (3, 2, 4, 36), 2, "SAME"), ((4, 8, 8, 8), (3, 3, 8, 24), 1, "VALID"),
    ((37, 12, 12, 20), (5, 5, 20, 50), 1, "VALID"),
    # single input channel: sliding-window kernel (conv_win_umma.cu); odd batch, kw <= 4 and kw > 4
    ((5, 28, 28, 1), (5, 5, 1, 20), 1, "VALID"), ((4, 10, 18, 1), (3, 3, 1, 32), 1, "VALID"),
    ((3, 15, 15, 1), (8, 8, 1, 4), 1, "VALID"), ((2, 9, 32, 1), (2, 1, 1, 12), 1, "VALID"),
    # raw-tile kernels with other images-per-tile counts: Ho = 4 (I = 4), Ho = 16 (I = 1), Ho = 1 (I = 16),
    # channel counts that are / are not multiples of 8, N = 64
    ((6, 8, 12, 8), (5, 5, 8, 32), 1, "VALID"), ((3, 20, 12, 4), (5, 5, 4, 24), 1, "VALID"),
    ((21, 5, 12, 16), (5, 5, 16, 64), 1, "VALID"), ((9, 9, 10, 12), (2, 3, 12, 40), 1, "VALID"),
])
def test_conv_tensorcore(tc_ctx, shape, wshape, stride, padding):
    rng = np.random.RandomState(sum(shape) + sum(wshape))
    P = O.make_layer_params(rng, wshape, flow="iaf", std_init=10.0, dtype=np.float64)
    F = wshape[-1]
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(F)
    P["mean_b"][...] = 0.1 * rng.standard_normal(F)
    x = rng.uniform(size=shape)
    y0 = O.conv2d(x, P["mean_W"], stride, padding)
    B = shape[0]
    eps_z, eps = rng.standard_normal((B, F)), rng.standard_normal(y0.shape)
    ref, _ = O.conv_call(P, x, eps_z, eps, 1.0, stride=stride, padding=padding)
    lyr = make_layer(P, x.shape, stride=stride, padding=padding)
    got = lyr(_f32(x), noise=dict(eps_z=_f32(eps_z), eps_out=_f32(eps))).numpy()
    close(got, ref, RTOL, "conv call (tf32x3)")
    # which kernel family served the shape (mnf_last_path): the parity above is about THAT kernel
    from tf_mnf_b200 import _lib as L

    C_in, Ho, Wo = shape[3], got.shape[1], got.shape[2]
    raw_ok = padding == "VALID" and stride == 1 and Wo == 8 and C_in % 4 == 0 and 20 < F <= 64 and 16 % Ho == 0
    chunks = -(-C_in // 8) * (Ho - 1 + wshape[0]) * (16 // max(Ho, 1)) * (7 + wshape[1]) if raw_ok else 0
    if padding == "VALID" and stride == 1 and C_in == 1:
        assert L.last_path() == "conv_win_umma", L.last_path()
    elif raw_ok and chunks <= 5 * 192:  # the producers hold a whole tile in registers (MAXC chunks per thread)
        assert L.last_path() == "conv_raw_f16", L.last_path()
    else:
        assert L.last_path() in ("paired_umma<64>", "conv_raw_umma", "conv_direct"), L.last_path()


@pytest.mark.parametrize("wshape,shape", [((5, 5, 20, 50), (7, 12, 12, 20)), ((3, 3, 8, 40), (9, 10, 10, 8)),
                                          ((5, 5, 12, 64), (4, 12, 12, 12))])
def test_conv_f16_split_dynamic_range(tc_ctx, wshape, shape):
    """conv_raw_f16.cu scales every image and every weight column by a power of two before the
    fp16 hi/lo split: images whose magnitudes differ by 1e8, zero images and weight columns of
    very different size must each meet the parity bar relative to THEIR OWN output scale."""
    rng = np.random.RandomState(5)
    P = O.make_layer_params(rng, wshape, flow="iaf", std_init=10.0, dtype=np.float64)
    F = wshape[-1]
    P["mean_W"] *= np.exp(rng.uniform(-8, 8, size=F))            # column scales over 7 decades
    P["mean_b"][...] = 0.0
    P["log_var_b"][...] = -60.0                                  # bias variance ~ 0: sd comes from x^2 * sigma^2
    x = rng.standard_normal(shape)
    x *= np.array([1e-4, 1.0, 300.0, 1e4, 0.0, 3e-3, 7.0, 1e2, 1e-2][: shape[0]]).reshape(-1, 1, 1, 1)
    B = shape[0]
    y0 = O.conv2d(x, P["mean_W"], 1, "VALID")
    eps_z, eps = rng.standard_normal((B, F)), rng.standard_normal(y0.shape)
    ref, _ = O.conv_call(P, x, eps_z, eps, 1.0, stride=1, padding="VALID")
    lyr = make_layer(P, x.shape)
    got = lyr(_f32(x), noise=dict(eps_z=_f32(eps_z), eps_out=_f32(eps))).numpy()
    for b in range(B):
        for n in range(0, F, 7):
            r = ref[b, ..., n]
            if np.max(np.abs(r)) > 1e-30:
                close(got[b, ..., n], r, 2e-5, f"image {b} filter {n}")
    close(got, ref, RTOL, "whole tensor")


@pytest.mark.parametrize("shape,wshape", [((64, 12, 12, 20), (5, 5, 20, 50)), ((301, 28, 28, 1), (5, 5, 1, 20))])
def test_tensorcore_matches_ffma_with_philox(tc_ctx, shape, wshape):
    """Same seed => the tensor-core and FFMA paths draw identical in-kernel noise; outputs agree to 1e-5."""
    rng = np.random.RandomState(3)
    P = O.make_layer_params(rng, wshape, flow="rnvp", std_init=10.0, dtype=np.float64)
    x = _f32(rng.uniform(size=shape))
    lyr = make_layer(P, x.shape)
    y_tc = lyr(x).numpy()
    tc_ctx.set_math("fp32")
    lyr.call_index = 0
    y_ff = lyr(x).numpy()
    close(y_tc, y_ff, RTOL, "tf32x3 vs fp32 (Philox noise)")


@pytest.mark.parametrize("kind", ["iaf", "rnvp"])
@pytest.mark.parametrize("D,B,masks", [(20, 37, "ones"), (50, 300, "autoregressive"), (800, 130, "ones"),
                                       (136, 64, "autoregressive"), (10, 129, "ones"), (500, 257, "ones"),
                                       (2048, 40, "ones")])
def test_flow_chain_tensorcore(tc_ctx, kind, D, B, masks):
    """IAF / RNVP steps on tcgen05 (flows_umma.cu) vs the float64 oracle, incl. saved activations.
    The fp32-parity mode keeps tcgen05 to D <= 800 (one TMEM accumulator sums D terms and the
    tensor core's fp32 accumulation is not round-to-nearest: csrc/flows.cu); wider rows take the
    FFMA kernel, or tcgen05 under the bf16 mode (test_gpu_bf16.py)."""
    from gpu_util import make_flows

    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.flows import NormalizingFlow

    if kind != "iaf" and masks == "autoregressive":
        pytest.skip("masks only apply to IAF")
    rng = np.random.RandomState(D + B)
    steps = O.make_flow(rng, kind, D, 3, (50,), masks, np.float64)
    for st in steps:
        for k, v in st.items():
            if k.startswith("b"):
                for b in v if isinstance(v, list) else [v]:
                    b[...] = 0.1 * rng.standard_normal(b.shape)
        if kind == "iaf":
            st["W"][1] *= 0.25
    z = rng.standard_normal((B, D))
    bern = [(rng.uniform(size=(B, D)) <= 0.5).astype(np.float64) for _ in range(3)]
    ref_states, ref_ld, caches = O.flow_forward(z, steps, bern)
    nf = NormalizingFlow(make_flows(steps, D))
    states, ld = nf.forward(_f32(z), bern=[_f32(b) for b in bern] if kind == "rnvp" else None, record=True)
    for k, (a, b) in enumerate(zip(states, ref_states)):
        close(a.numpy(), b, RTOL, f"state {k}")
    close(ld.numpy(), ref_ld, RTOL, "log_dets")
    # hidden activations saved for backward: first step's h
    h_ref = caches[0]["acts"][1] if kind == "iaf" else caches[0]["hs"][0]
    h_got = nf.last_run.saved.numpy()[: B * 50].reshape(B, 50)
    close(h_got, h_ref, RTOL, "saved hidden")
    f16 = kind == "rnvp" and D % 4 == 0 and B >= 32  # flows_f16.cu: scaled fp16 hi/lo operands, bulk-copied z tiles
    assert L.last_path() == (("flow_f16" if f16 else "flow_umma") if D <= 800 else "flow_gemm" if kind == "iaf" else "flow_mlp"), L.last_path()


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("shape,wshape", [((6, 28, 28, 1), (5, 5, 1, 20)), ((5, 12, 12, 20), (5, 5, 20, 50)),
                                          ((3, 10, 12, 4), (3, 3, 4, 8))])
def test_conv_fused_relu_pool_equals_unfused(tc_ctx, math, shape, wshape):
    """ReLU + MaxPool folded into the conv epilogue == conv, then mnf_relu_maxpool2 (bit-exact:
    same arithmetic and the same Philox draws)."""
    from tf_mnf_b200.models import ReLUMaxPool2D

    tc_ctx.set_math(math)
    rng = np.random.RandomState(11)
    P = O.make_layer_params(rng, wshape, flow="rnvp", std_init=10.0, dtype=np.float64)
    x = _f32(rng.uniform(size=shape))
    lyr = make_layer(P, x.shape)
    y_fused = lyr.call(x, fuse_relu_pool=True)
    fused = lyr.last_fused
    lyr.call_index = 0
    y_ref = ReLUMaxPool2D(padding="SAME")(lyr.call(x))
    if fused:
        assert y_fused.shape == y_ref.shape
        np.testing.assert_array_equal(y_fused.numpy(), y_ref.numpy())
    else:  # FFMA path with F > 20 has no fused epilogue: the call must have stayed unfused
        assert math == "fp32" and wshape[-1] > 20 and y_fused.shape[1] == shape[1] - wshape[0] + 1


@pytest.mark.parametrize("math", ["tf32x3", "fp32"])
@pytest.mark.parametrize("K,N,B,act", [(500, 10, 257, "softmax"), (500, 10, 64, "relu"), (800, 500, 130, "relu"),
                                       (64, 16, 33, "softmax"), (24, 4, 5, None)])
def test_dense_fused_activation(tc_ctx, math, K, N, B, act):
    """MNFDense with the following ReLU / Softmax folded into the kernel's epilogue
    (mnf_dense_forward_act) == MNFDense, then the stock layer; the small-N kernel (N <= 16) is also
    pinned against the float64 oracle."""
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.models import ReLU, Softmax

    tc_ctx.set_math(math)
    rng = np.random.RandomState(K + N)
    P = O.make_layer_params(rng, (K, N), flow="rnvp", std_init=10.0, dtype=np.float64)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    x = rng.standard_normal((B, K))
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    bern = [(rng.uniform(size=(B, K)) <= 0.5).astype(np.float64) for _ in range(2)]
    lyr = make_layer(P, x.shape)
    inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern])
    y_plain = lyr.call(_f32(x), noise=inj)
    if N <= 16:
        assert L.last_path() == "dense_small", L.last_path()
        ref, _ = O.dense_call(P, x, eps_z, eps, 1.0, bern_q=bern)
        close(y_plain.numpy(), ref, RTOL, "small-N dense vs oracle")
    if act is None:
        return
    y_fused = lyr.call(_f32(x), noise=inj, activation=act)
    stock = ReLU() if act == "relu" else Softmax()
    y_ref = stock(y_plain).numpy()
    if lyr.last_fused:
        if act == "relu":
            np.testing.assert_array_equal(y_fused.numpy(), y_ref)
        else:
            np.testing.assert_allclose(y_fused.numpy(), y_ref, rtol=2e-6, atol=1e-7)
    else:  # FFMA path for wide layers has no fused epilogue: the call stayed unfused
        assert math == "fp32" and N > 16
        np.testing.assert_array_equal(y_fused.numpy(), y_plain.numpy())


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
@pytest.mark.parametrize("D,B", [(1536, 70), (4096, 33)])
def test_wide_iaf_chain_as_gemms(tc_ctx, math, D, B):
    """Wide IAF steps on few rows run their two MLP contractions as whole-chip GEMMs (csrc/flows.cu
    `flow_gemm`; split-K keeps every tensor-core accumulation chain <= 1024 terms): fp32 parity in
    both math modes, sampled z0 included, parity reversal, log-dets and the saved activations."""
    from gpu_util import make_flows

    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.flows import NormalizingFlow

    tc_ctx.set_math(math)
    rng = np.random.RandomState(D + B)
    steps = O.make_flow(rng, "iaf", D, 2, (50,), "ones", np.float64)
    for st in steps:
        for k, v in st.items():
            if k.startswith("b"):
                for b in v if isinstance(v, list) else [v]:
                    b[...] = 0.1 * rng.standard_normal(b.shape)
        st["W"][1] *= 0.25
    z = rng.standard_normal((B, D))
    ref_states, ref_ld, caches = O.flow_forward(z, steps, None)
    nf = NormalizingFlow(make_flows(steps, D))
    states, ld = nf.forward(_f32(z), record=True)
    assert L.last_path() == "flow_gemm", L.last_path()
    for k, (a, b) in enumerate(zip(states, ref_states)):
        close(a.numpy(), b, RTOL, f"state {k}")
    close(ld.numpy(), ref_ld, RTOL, "log_dets")
    close(nf.last_run.saved.numpy()[: B * 50].reshape(B, 50), caches[0]["acts"][1], RTOL, "saved hidden")


@pytest.mark.parametrize("K,N,B,use_z,act", [(4096, 512, 256, True, None), (1500, 70, 33, True, "relu"), (2048, 24, 5, False, None),
                                             (4096, 512, 1100, True, "relu"), (1504, 72, 1024, False, None)])
def test_dense_long_contraction_keeps_fp32_parity(tc_ctx, K, N, B, use_z, act):
    """K > 1280 in the fp32-parity tensor-core mode: one TMEM accumulation chain of K terms would
    drift past 1e-5 (2.7e-5 measured at K = 4096), so the layer runs as split-K tcgen05 GEMMs with
    chains of <= 1024 terms plus an epilogue kernel (`dense_long_gemm`) or, for many rows, as the packed
    fp16 GEMM launched once per 1024-term segment with fp32 partial sums in between (`paired_f16x3<128>/segments`)."""
    from tf_mnf_b200 import _lib as L

    rng = np.random.RandomState(K + N)
    P = O.make_layer_params(rng, (K, N), flow="planar", std_init=10.0, dtype=np.float64)
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    x = rng.standard_normal((B, K))
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    ref, _ = O.dense_call(P, x, eps_z, eps, 1.0, use_z=use_z)
    if act == "relu":
        ref = np.maximum(ref, 0)
    lyr = make_layer(P, x.shape, use_z=use_z)
    inj = dict(eps_out=_f32(eps))
    if use_z:
        inj["eps_z"] = _f32(eps_z)
    y = lyr.call(_f32(x), noise=inj, activation=act)
    assert L.last_path() == ("paired_f16x3<128>/segments" if B >= 1024 else "dense_long_gemm"), L.last_path()
    assert act is None or lyr.last_fused
    close(y.numpy(), ref, RTOL, "dense call (long K, tf32x3)")


@pytest.mark.parametrize("K,N,B,use_z", [(800, 500, 1300, True), (64, 17, 1024, False), (1280, 130, 1025, True)])
def test_dense_f16x3_many_rows(tc_ctx, K, N, B, use_z):
    """dense_f16x3.cu (M >= 1024): injected noise against the float64 oracle, sd_out included; rows whose
    magnitudes differ by 1e8 and weight columns 7 decades apart must each meet the bar relative to
    THEIR OWN output scale (per-row and per-column power-of-two scales)."""
    from tf_mnf_b200 import _lib as L

    rng = np.random.RandomState(K + N)
    P = O.make_layer_params(rng, (K, N), flow="iaf", std_init=10.0, dtype=np.float64)
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.0
    P["log_var_b"][...] = -60.0
    P["mean_W"] *= np.exp(rng.uniform(-8, 8, size=N))
    x = rng.standard_normal((B, K))
    x *= np.exp(rng.uniform(-9, 9, size=(B, 1)))
    x[5] = 0.0
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    ref, c = O.dense_call(P, x, eps_z, eps, 1.0, use_z=use_z)
    lyr = make_layer(P, x.shape, use_z=use_z)
    lyr.training = True  # sd_out is written
    inj = dict(eps_out=_f32(eps))
    if use_z:
        inj["eps_z"] = _f32(eps_z)
    got = lyr(_f32(x), noise=inj).numpy()
    assert L.last_path() == "paired_f16x3<128>", L.last_path()
    close(got, ref, RTOL, "whole tensor")
    close(lyr._tape["sd"].numpy(), c["sq"], RTOL, "sd_out")
    for b in range(0, B, 97):
        if np.max(np.abs(ref[b])) > 1e-30:
            close(got[b], ref[b], 2e-5, f"row {b}")
    for n in range(0, N, 7):
        close(got[:, n], ref[:, n], 2e-5, f"column {n}")
