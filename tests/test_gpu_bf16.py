"""MNF_MATH_BF16: BASELINE.json's "wide MNFDense 4096 -> 4096 with planar flows, batch 8192, bf16
tensor-core path" (SURVEY.md section 8d, C4).  MNFDense.call (mnf_dense.py:159-170) with bf16
operands and fp32 accumulation against the float64 oracle.

Stated tolerance.  x*z, x^2, W and sigma^2 are each rounded to nearest bf16 (relative error
<= 2^-9), so every product carries <= 2^-8 and
    |y - y_ref| <= 2^-8 * (|x*z| @ |W|) + 2^-9 * sd * |eps|   (+ fp32 accumulation)
holds element by element; the tests assert it with a factor 2 of slack (2^-7), and
rel-L2(y) <= 1e-2 (the rounding errors of a K-term sum add up like sqrt(K): ~2e-3 typically)."""
from __future__ import annotations

import numpy as np
import pytest
from gpu_util import make_layer

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bf_ctx():
    import tf_mnf_b200 as M

    ctx = M.default_context()
    ctx.set_math("bf16")
    yield ctx
    ctx.set_math("tf32x3")


def _f32(a):
    return np.asarray(a, np.float32)


def _check(y, ref, c, P):
    y = np.asarray(y, np.float64)
    bound = 2.0 ** -7 * (np.abs(c["A"]) @ np.abs(P["mean_W"]) + c["sq"] * np.abs(c["eps"])) + 1e-6
    err = np.abs(y - ref)
    worst = np.max(err / bound)
    assert worst <= 1.0, f"per-element bf16 bound exceeded: max err/bound = {worst:.3f}"
    rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    assert rel <= 1e-2, f"rel-L2 = {rel:.3e}"
    return rel


@pytest.mark.parametrize("flow", ["planar", "iaf"])
@pytest.mark.parametrize("K,N,B,use_z", [(512, 384, 300, True), (200, 300, 77, True), (136, 100, 1024, True),
                                         (64, 128, 128, False), (130, 70, 129, True), (1024, 17, 5, True)])
def test_dense_bf16(bf_ctx, flow, K, N, B, use_z):
    rng = np.random.RandomState(K * 7 + N)
    P = O.make_layer_params(rng, (K, N), flow=flow, std_init=10.0, dtype=np.float64)
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    x = rng.standard_normal((B, K))
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    ref, c = O.dense_call(P, x, eps_z, eps, 1.0, use_z=use_z)
    lyr = make_layer(P, x.shape, use_z=use_z)
    inj = dict(eps_out=_f32(eps))
    if use_z:
        inj["eps_z"] = _f32(eps_z)
    y = lyr(_f32(x), noise=inj).numpy()
    from tf_mnf_b200 import _lib as L

    assert L.last_path() == "paired_bf16<128>", L.last_path()
    rel = _check(y, ref, c, P)
    assert rel > 1e-6, "suspiciously exact: the bf16 path did not run"


def test_bf16_is_deterministic_and_close_to_tf32x3(bf_ctx):
    """Same Philox noise in both math modes: bf16 must land within its tolerance of the
    fp32-parity tensor-core result, with the fused ReLU, frozen weights, and repeated calls."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.layers import MNFDense

    M.set_seed(11)
    rng = np.random.RandomState(0)
    x = _f32(rng.standard_normal((513, 768)))
    lyr = MNFDense(640, flow_type="planar", std_init=10.0)
    lyr.build(x.shape)
    outs = {}
    for mode in ("tf32x3", "bf16"):
        bf_ctx.set_math(mode)
        lyr.call_index = 0
        outs[mode] = lyr(x).numpy()
    bf_ctx.freeze_weights(True)
    try:
        lyr.call_index = 0
        a = lyr(x).numpy()
        lyr.call_index = 0
        b = lyr(x).numpy()
    finally:
        bf_ctx.freeze_weights(False)
    assert np.array_equal(a, outs["bf16"]) and np.array_equal(a, b)
    rel = np.linalg.norm(outs["bf16"].astype(np.float64) - outs["tf32x3"]) / np.linalg.norm(outs["tf32x3"])
    assert 1e-6 < rel <= 1e-2, rel

