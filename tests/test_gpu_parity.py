"""GPU parity: the CUDA path (through the ctypes C-ABI and the host classes) against
(1) the committed golden vectors (reference source run on the TF-shim, float64) and
(2) the NumPy oracle on seeded inputs.  fp32 bar: 1e-5 relative to the tensor's max
magnitude (north_star: "within 1e-5 relative in fp32")."""
from __future__ import annotations

import numpy as np
import pytest
from conftest import fix_params, fix_steps, load_golden, split_draws
from gpu_util import box_muller_f64, close, make_flows, make_layer, philox_words

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module", params=["tf32x3", "fp32"])
def ctx(request):
    """Both arithmetic modes: the library default (tcgen05, hi/lo operand splitting) and the FFMA kernels."""
    import tf_mnf_b200 as M

    c = M.default_context()
    c.set_math(request.param)
    yield c
    c.set_math("tf32x3")


# ------------------------------------------------------------------------------- noise
def test_philox_raw_bit_exact(ctx):
    import ctypes as C

    from tf_mnf_b200 import _lib as L

    for seed, stream, off, rows, outer, inner in [(0, 0, 0, 3, 1, 7), (0x1234567890ABCDEF, 77, 5, 4, 3, 10),
                                                  (42, 0x10203, (1 << 33) + 9, 2, 2, 4)]:
        out = ctx.empty((rows, outer, inner), "u4")
        L.check(ctx.lib.mnf_philox_raw(ctx.handle, seed, stream, off, rows, outer, inner, C.c_void_p(out.ptr)))
        np.testing.assert_array_equal(out.numpy(), philox_words(seed, stream, off, rows, outer, inner))


def test_philox_normal_values_and_moments(ctx):
    import ctypes as C

    from tf_mnf_b200 import _lib as L

    out = ctx.empty((5, 3, 10))
    L.check(ctx.lib.mnf_philox_normal(ctx.handle, 7, 3, 11, 5, 3, 10, C.c_void_p(out.ptr)))
    w = philox_words(7, 3, 11, 5, 3, 12)
    ref = box_muller_f64(w)[:, :, :10]
    np.testing.assert_allclose(out.numpy(), ref, atol=2e-5)
    big = ctx.empty((4096, 1, 1024))
    L.check(ctx.lib.mnf_philox_normal(ctx.handle, 1, 2, 0, 4096, 1, 1024, C.c_void_p(big.ptr)))
    v = big.numpy().ravel().astype(np.float64)
    n = v.size
    assert abs(v.mean()) < 5 / np.sqrt(n)
    assert abs(v.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs((v**3).mean()) < 5 * np.sqrt(15 / n)
    assert abs((v**4).mean() - 3) < 5 * np.sqrt(96 / n)
    from scipy import stats

    assert stats.kstest(v[:200000], "norm").pvalue > 1e-4
    b = ctx.empty((2048, 512))
    L.check(ctx.lib.mnf_philox_bernoulli(ctx.handle, 1, 9, 0, 2048, 512, 0.5, C.c_void_p(b.ptr)))
    bm = b.numpy()
    assert set(np.unique(bm)) == {0.0, 1.0} and abs(bm.mean() - 0.5) < 5 * 0.5 / np.sqrt(bm.size)


# ------------------------------------------------------------------------------- flows
@pytest.mark.parametrize("name", ["iaf_h8", "rnvp_h8", "planar"])
def test_flow_chain_golden(ctx, name):
    from tf_mnf_b200.flows import NormalizingFlow

    fx = load_golden("flows")[name]
    steps = fix_steps(fx["steps"])
    _, berns = split_draws(fx["draws"])
    nf = NormalizingFlow(make_flows(steps, fx["z"].shape[1]))
    states, ld = nf.forward(fx["z"].astype(np.float32), bern=[b.astype(np.float32) for b in berns] or None)
    assert len(states) == len(fx["states"])
    for k, (a, b) in enumerate(zip(states, fx["states"])):
        close(a.numpy(), b, RTOL, f"{name} state {k}")
    close(ld.numpy(), fx["ld"], RTOL, f"{name} log_dets")


@pytest.mark.parametrize("kind", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("D,B,masks", [(20, 37, "ones"), (50, 1, "autoregressive"), (800, 70, "ones"), (136, 5, "autoregressive")])
def test_flow_chain_oracle(ctx, kind, D, B, masks):
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
        if kind == "iaf":  # keep exp(s) of three chained un-masked steps inside the fp32 range
            st["W"][1] *= 0.25
    z = rng.standard_normal((B, D))
    bern = [(rng.uniform(size=(B, D)) <= 0.5).astype(np.float64) for _ in range(3)]
    ref_states, ref_ld, _ = O.flow_forward(z, steps, bern)
    nf = NormalizingFlow(make_flows(steps, D))
    states, ld = nf.forward(z.astype(np.float32), bern=[b.astype(np.float32) for b in bern] if kind == "rnvp" else None)
    for k, (a, b) in enumerate(zip(states, ref_states)):
        close(a.numpy(), b, RTOL, f"state {k}")
    close(ld.numpy(), ref_ld, RTOL, "log_dets")


def test_flow_philox_bernoulli_matches_materialised(ctx):
    """In-kernel Bernoulli draws == mnf_philox_bernoulli with the same (seed, stream, row)."""
    import ctypes as C

    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200 import init, noise
    from tf_mnf_b200.flows import RNVP

    fl = RNVP(24, h_sizes=(16,))
    fl.build(24)
    z = ctx.to_device(np.random.RandomState(0).standard_normal((9, 24)).astype(np.float32))
    x1, ld1 = fl.forward(z)  # call 0, Philox mask
    m = ctx.empty((9, 24))
    seed = (init.global_seed() & 0xFFFFFFFF) | (0 << 32)
    L.check(ctx.lib.mnf_philox_bernoulli(ctx.handle, seed, noise.stream_id(fl.uid, noise.FLOW_BERN), 0, 9, 24, 0.5,
                                         C.c_void_p(m.ptr)))
    fl._calls = 0
    x2, ld2 = fl.forward(z, mask=m)
    np.testing.assert_array_equal(x1.numpy(), x2.numpy())
    np.testing.assert_array_equal(ld1.numpy(), ld2.numpy())


# ------------------------------------------------------------------------------ layers
DENSE = ["dense_iaf", "dense_iaf_clip", "dense_iaf_nf0", "dense_iaf_f3", "dense_noz", "dense_rnvp", "dense_planar"]
CONV = ["conv_iaf", "conv_iaf_clip", "conv_rnvp", "conv_planar"]


def _f32(a):
    return None if a is None else np.asarray(a, np.float32)


@pytest.mark.parametrize("name", DENSE + CONV)
def test_layer_golden(ctx, name):
    fx = load_golden(name)
    P = fix_params(fx["P"])
    use_z = bool(fx["use_z"])
    lyr = make_layer(P, fx["x"].shape, max_std=float(fx["max_std"]), use_z=use_z)
    # call
    normals, berns = split_draws(fx["call_draws"])
    inj = dict(eps_out=_f32(normals[-1]))
    if use_z:
        inj.update(eps_z=_f32(normals[0]), bern_q=[_f32(b) for b in berns] or None)
    y = lyr(fx["x"].astype(np.float32), noise=inj)
    close(y.numpy(), fx["y"], RTOL, f"{name} call")
    # sample_z
    normals, berns = split_draws(fx["sz_draws"])
    inj = dict(eps_z=_f32(normals[0]), bern_q=[_f32(b) for b in berns] or None) if use_z else None
    z, ld = lyr.sample_z(4, noise=inj)
    close(z.numpy(), fx["sz_z"], RTOL, f"{name} sample_z")
    if np.max(np.abs(fx["sz_ld"])) > 0:
        close(ld.numpy(), fx["sz_ld"], RTOL, f"{name} log_dets")
    else:
        assert np.all(ld.numpy() == 0)
    # kl_div
    normals, berns = split_draws(fx["kl_draws"])
    nq = sum(1 for s in P["flow_q"] if s["kind"] == "rnvp")
    inj = None
    if use_z:
        inj = dict(eps_z=_f32(normals[0]), eps_a=_f32(normals[1]), bern_q=[_f32(b) for b in berns[:nq]] or None,
                   bern_r=[_f32(b) for b in berns[nq:]] or None)
        if name.startswith("conv"):
            inj["eps_b"] = _f32(normals[2])
    kl = lyr.kl_div(noise=inj)
    close(kl.numpy(), np.asarray(fx["kl"]).reshape(1), RTOL, f"{name} kl")


def test_lenet_small_golden(ctx):
    from tf_mnf_b200.models import MNFLeNet

    fx = load_golden("lenet_small")
    Ps = [fix_params(P) for P in fx["Ps"]]
    model = MNFLeNet(3, 4, 6, 5, max_std=1, flow_h_sizes=[10], std_init=10.0)
    shapes = [(3, 28, 28, 1), (3, 12, 12, 3), (3, 64), (3, 6)]
    from gpu_util import assign_flow

    for lyr, P, shp in zip(model.mnf_layers, Ps, shapes):
        lyr.build(shp)
        for n in ("mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var", "r0_mean", "r0_log_var",
                  "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"):
            getattr(lyr, n).assign(np.asarray(P[n], np.float32))
        for fl, st in zip(lyr.flow_q.flows, P["flow_q"]):
            assign_flow(fl, st)
        for fl, st in zip(lyr.flow_r.flows, P["flow_r"]):
            assign_flow(fl, st)
    normals, _ = split_draws(fx["call_draws"])
    inj = [dict(eps_z=_f32(normals[2 * i]), eps_out=_f32(normals[2 * i + 1])) for i in range(4)]
    probs = model(fx["x"].astype(np.float32), noise=inj)
    close(probs.numpy(), fx["probs"], RTOL, "lenet probs")
    kn, _ = split_draws(fx["kl_draws"])
    kinj = [dict(eps_z=_f32(kn[0]), eps_a=_f32(kn[1]), eps_b=_f32(kn[2])),
            dict(eps_z=_f32(kn[3]), eps_a=_f32(kn[4]), eps_b=_f32(kn[5])),
            dict(eps_z=_f32(kn[6]), eps_a=_f32(kn[7])), dict(eps_z=_f32(kn[8]), eps_a=_f32(kn[9]))]
    kl = model.kl_div(noise=kinj)
    close(kl.numpy(), np.asarray(fx["kl"]).reshape(1), RTOL, "lenet kl")


@pytest.mark.parametrize("flow", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("K,N,B", [(800, 500, 130), (500, 10, 257), (136, 100, 64), (25, 3, 1)])
def test_dense_oracle(ctx, flow, K, N, B):
    rng = np.random.RandomState(K + N)
    P = O.make_layer_params(rng, (K, N), flow=flow, std_init=10.0, dtype=np.float64)
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    x = rng.standard_normal((B, K))
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    bern = [(rng.uniform(size=(B, K)) <= 0.5).astype(np.float64) for _ in range(2)]
    ref, _ = O.dense_call(P, x, eps_z, eps, 1.0, bern_q=bern)
    lyr = make_layer(P, x.shape)
    inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern] if flow == "rnvp" else None)
    close(lyr(_f32(x), noise=inj).numpy(), ref, RTOL, "dense call")
    # KL at full layer size
    eps_z1, eps_a = rng.standard_normal((1, K)), rng.standard_normal(N)
    b1 = [(rng.uniform(size=(1, K)) <= 0.5).astype(np.float64) for _ in range(4)]
    for r_term in ("all_states", "last"):
        lyr.r_term = r_term
        ref_kl, _ = O.dense_kl(P, eps_z1, eps_a, r_term=r_term, bern_q=b1[:2], bern_r=b1[2:])
        kinj = dict(eps_z=_f32(eps_z1), eps_a=_f32(eps_a))
        if flow == "rnvp":
            kinj.update(bern_q=[_f32(b) for b in b1[:2]], bern_r=[_f32(b) for b in b1[2:]])
        close(lyr.kl_div(noise=kinj).numpy(), np.asarray(ref_kl).reshape(1), RTOL, f"dense kl {r_term}")


@pytest.mark.parametrize("flow", ["iaf", "rnvp"])
@pytest.mark.parametrize("shape,wshape,stride,padding", [
    ((6, 28, 28, 1), (5, 5, 1, 20), 1, "VALID"), ((5, 12, 12, 20), (5, 5, 20, 50), 1, "VALID"),
    ((3, 9, 7, 3), (3, 3, 3, 64), 1, "SAME"), ((2, 11, 10, 4), (3, 2, 4, 6), 2, "SAME"), ((2, 8, 8, 2), (3, 3, 2, 5), 2, "VALID"),
])
def test_conv_oracle(ctx, flow, shape, wshape, stride, padding):
    rng = np.random.RandomState(sum(shape) + sum(wshape))
    P = O.make_layer_params(rng, wshape, flow=flow, std_init=10.0, dtype=np.float64)
    F = wshape[-1]
    P["q0_log_var"][...] = -3 + 0.5 * rng.standard_normal(F)
    P["mean_b"][...] = 0.1 * rng.standard_normal(F)
    x = rng.uniform(size=shape)
    y0 = O.conv2d(x, P["mean_W"], stride, padding)
    B = shape[0]
    eps_z, eps = rng.standard_normal((B, F)), rng.standard_normal(y0.shape)
    bern = [(rng.uniform(size=(B, F)) <= 0.5).astype(np.float64) for _ in range(2)]
    ref, _ = O.conv_call(P, x, eps_z, eps, 1.0, stride=stride, padding=padding, bern_q=bern)
    lyr = make_layer(P, x.shape, stride=stride, padding=padding)
    inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern] if flow == "rnvp" else None)
    close(lyr(_f32(x), noise=inj).numpy(), ref, RTOL, "conv call")
    Kc = int(np.prod(wshape[:3]))
    eps_z1, eps_a, eps_b = rng.standard_normal((1, F)), rng.standard_normal(Kc), rng.standard_normal()
    b1 = [(rng.uniform(size=(1, F)) <= 0.5).astype(np.float64) for _ in range(4)]
    ref_kl, _ = O.conv_kl(P, eps_z1, eps_a, eps_b, bern_q=b1[:2], bern_r=b1[2:])
    kinj = dict(eps_z=_f32(eps_z1), eps_a=_f32(eps_a), eps_b=np.float32(eps_b))
    if flow == "rnvp":
        kinj.update(bern_q=[_f32(b) for b in b1[:2]], bern_r=[_f32(b) for b in b1[2:]])
    close(lyr.kl_div(noise=kinj).numpy(), np.asarray(ref_kl).reshape(1), RTOL, "conv kl")


# ------------------------------------------------ full-size properties (BASELINE config C3)
def test_lenet_full_size_properties(ctx):
    """B = 10240 (1024 samples x 10 images): shard invariance (same seed + global row =>
    bit-identical rows for any row partition), in-kernel Philox == injected materialised
    noise, rows of a softmax sum to 1, and rows sharing an image differ (fresh z per row)."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFLeNet

    M.set_seed(0)
    model = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=10.0, flow_type="rnvp")
    rng = np.random.RandomState(0)
    imgs = rng.uniform(size=(10, 28, 28, 1)).astype(np.float32)
    x = np.repeat(imgs, 1024, axis=0)
    xd = ctx.to_device(x)
    full = model(xd).numpy()
    assert full.shape == (10240, 10) and np.all(np.isfinite(full))
    np.testing.assert_allclose(full.sum(1), 1.0, atol=1e-5)
    assert np.abs(full[0] - full[1]).max() > 0  # same image, different posterior sample
    parts = []
    for lo, hi in [(0, 2560), (2560, 5120), (5120, 10240)]:
        for lyr in model.mnf_layers:
            lyr.call_index = 0
        model.set_row_offset(lo)
        parts.append(model(xd[lo:hi]).numpy())
    np.testing.assert_array_equal(np.concatenate(parts), full)
    kl = float(model.kl_div().numpy()[0])
    assert np.isfinite(kl)
