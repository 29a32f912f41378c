"""GPU parity of the hand-written backward kernels against the oracle's analytic backward
(float64; itself pinned by finite differences in test_oracle_grad.py).  The reference has
no gradients of its own: it relies on TF autodiff of the forward the oracle restates."""
from __future__ import annotations

import os

import numpy as np
import pytest
from gpu_util import close, make_flows, make_layer

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5   # activations-like tensors (dx, dz)
RTOL_G = float(os.environ.get("MNF_RTOL_G", 1e-5))  # parameter gradients: the same bar as activations (north_star)


@pytest.fixture(scope="module", params=["fp32", "tf32x3"])
def ctx(request):
    """Both arithmetic modes: FFMA everywhere, and the tensor-core mode, whose backward
    contractions run as 3xTF32 on tcgen05 (csrc/gemm_umma.cu) -- same fp32 parity bar."""
    import tf_mnf_b200 as M

    c = M.default_context()
    c.set_math(request.param)
    yield c
    c.set_math("tf32x3")


def _f32(a):
    return None if a is None else np.asarray(a, np.float32)


def flow_grad_map(kind, g):
    """oracle grad dict -> host variable names (tf_mnf_b200.flows.*.variables())."""
    if kind == "iaf":
        out = {}
        for i, (W, b) in enumerate(zip(g["W"], g["b"]), 1):
            out[f"w{i}"], out[f"b{i}"] = W, b
        return out
    if kind == "rnvp":
        out = {"net_kernel": g["Wn"][0], "net_bias": g["bn"][0]}
        for i in range(1, len(g["Wn"])):
            out[f"net_kernel_{i}"], out[f"net_bias_{i}"] = g["Wn"][i], g["bn"][i]
        out.update({"t_kernel": g["Wt"], "t_bias": g["bt"], "s_kernel": g["Ws"], "s_bias": g["bs"]})
        return out
    return {"u": g["u"], "w": g["w"], "b": g["b"]}


def perturb_biases(steps, rng):
    for st in steps:
        for k, v in st.items():
            if k.startswith("b"):
                for b in v if isinstance(v, list) else [v]:
                    b[...] = 0.1 * rng.standard_normal(b.shape)
        if st["kind"] == "iaf":
            st["W"][1] *= 0.25


@pytest.mark.parametrize("kind", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("D,B,masks", [(20, 37, "ones"), (136, 70, "autoregressive"), (50, 1, "ones"), (800, 33, "ones")])
def test_flow_chain_backward(ctx, kind, D, B, masks):
    from tf_mnf_b200.flows import NormalizingFlow

    if kind != "iaf" and masks == "autoregressive":
        pytest.skip("masks only apply to IAF")
    rng = np.random.RandomState(D * 7 + B)
    T = 3
    steps = O.make_flow(rng, kind, D, T, (50,), masks, np.float64)
    perturb_biases(steps, rng)
    z = rng.standard_normal((B, D))
    bern = [(rng.uniform(size=(B, D)) <= 0.5).astype(np.float64) for _ in range(T)]
    cs = [rng.standard_normal((B, D)) for _ in range(T + 1)]
    czT = rng.standard_normal((B, D))
    cl = rng.standard_normal(B)
    states, ld, caches = O.flow_forward(z, steps, bern)
    dcs = [c.copy() for c in cs]
    dcs[-1] = dcs[-1] + czT
    ref_dz, ref_g = O.flow_backward(dcs, cl, steps, caches)

    nf = NormalizingFlow(make_flows(steps, D))
    nf.forward(_f32(z), bern=[_f32(b) for b in bern] if kind == "rnvp" else None, record=True)
    d_z0, grads = nf.backward(_f32(np.stack(cs)), _f32(czT), _f32(cl))
    close(d_z0.numpy(), ref_dz, RTOL, "d_z0")
    for k in range(T):
        for name, ref in flow_grad_map(kind, ref_g[k]).items():
            close(grads[k][name].numpy().reshape(np.shape(ref)), ref, RTOL_G, f"step {k} {name}")


def _check_layer_grads(lyr, ref_g, kind, skip=()):
    for name in ("mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var", "r0_mean", "r0_log_var",
                 "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"):
        if name in ref_g and name not in skip:
            close(lyr.grads[name].numpy(), ref_g[name], RTOL_G, name)
    for tag in ("flow_q", "flow_r"):
        for k, g in enumerate(ref_g.get(tag, [])):
            for name, ref in flow_grad_map(kind, g).items():
                close(lyr.grads[f"{tag}/{k}/{name}"].numpy().reshape(np.shape(ref)), ref, RTOL_G, f"{tag}/{k}/{name}")


@pytest.mark.parametrize("flow", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("K,N,B,max_std", [(50, 30, 33, 1.0), (800, 500, 64, 1.0), (136, 100, 128, 1.3e-4), (500, 10, 1, 1.0)])
def test_dense_backward(ctx, flow, K, N, B, max_std):
    rng = np.random.RandomState(K + N + B)
    P = O.make_layer_params(rng, (K, N), flow=flow, std_init=2.0, dtype=np.float64)
    P["q0_log_var"][...] = -2 + 0.5 * rng.standard_normal(K)
    P["mean_b"][...] = 0.1 * rng.standard_normal(N)
    perturb_biases(P["flow_q"], rng)
    x = rng.standard_normal((B, K))
    eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    bern = [(rng.uniform(size=(B, K)) <= 0.5).astype(np.float64) for _ in range(2)]
    dy = rng.standard_normal((B, N))
    _, c = O.dense_call(P, x, eps_z, eps, max_std, bern_q=bern)
    ref_dx, ref_g = O.dense_call_backward(P, c, dy)

    lyr = make_layer(P, x.shape, max_std=max_std)
    lyr.training = True
    inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern] if flow == "rnvp" else None)
    lyr(_f32(x), noise=inj)
    lyr.zero_grads()
    dx = lyr.backward(_f32(dy))
    close(dx.numpy(), ref_dx, RTOL, "dx")
    _check_layer_grads(lyr, ref_g, flow)


@pytest.mark.parametrize("flow", ["iaf", "rnvp"])
@pytest.mark.parametrize("shape,wshape,stride,padding,max_std", [
    ((6, 28, 28, 1), (5, 5, 1, 20), 1, "VALID", 1.0), ((5, 12, 12, 20), (5, 5, 20, 50), 1, "VALID", 1.0),
    ((3, 9, 7, 3), (3, 3, 3, 64), 1, "SAME", 1.3e-4), ((2, 11, 10, 4), (3, 2, 4, 6), 2, "SAME", 1.0),
    ((2, 8, 8, 2), (3, 3, 2, 5), 2, "VALID", 1.0),
])
def test_conv_backward(ctx, flow, shape, wshape, stride, padding, max_std):
    rng = np.random.RandomState(sum(shape) * 3 + sum(wshape))
    P = O.make_layer_params(rng, wshape, flow=flow, std_init=2.0, dtype=np.float64)
    F = wshape[-1]
    P["q0_log_var"][...] = -2 + 0.5 * rng.standard_normal(F)
    P["mean_b"][...] = 0.1 * rng.standard_normal(F)
    perturb_biases(P["flow_q"], rng)
    x = rng.standard_normal(shape)
    B = shape[0]
    y0 = O.conv2d(x, P["mean_W"], stride, padding)
    eps_z, eps = rng.standard_normal((B, F)), rng.standard_normal(y0.shape)
    bern = [(rng.uniform(size=(B, F)) <= 0.5).astype(np.float64) for _ in range(2)]
    dy = rng.standard_normal(y0.shape)
    _, c = O.conv_call(P, x, eps_z, eps, max_std, stride=stride, padding=padding, bern_q=bern)
    ref_dx, ref_g = O.conv_call_backward(P, c, dy)

    lyr = make_layer(P, x.shape, max_std=max_std, stride=stride, padding=padding)
    lyr.training = True
    inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern] if flow == "rnvp" else None)
    lyr(_f32(x), noise=inj)
    lyr.zero_grads()
    dx = lyr.backward(_f32(dy))
    close(dx.numpy(), ref_dx, RTOL, "dx")
    _check_layer_grads(lyr, ref_g, flow)


@pytest.mark.parametrize("flow", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("r_term", ["all_states", "last"])
@pytest.mark.parametrize("wshape", [(60, 35), (800, 500), (5, 5, 1, 20), (5, 5, 20, 50)])
def test_kl_backward(ctx, flow, r_term, wshape):
    rng = np.random.RandomState(len(wshape) * 11 + wshape[0])
    conv = len(wshape) == 4
    P = O.make_layer_params(rng, wshape, flow=flow, std_init=1.0, dtype=np.float64)
    D = wshape[-1] if conv else wshape[0]
    P["q0_log_var"][...] = -2 + 0.5 * rng.standard_normal(D)
    P["mean_b"][...] = 0.1 * rng.standard_normal(wshape[-1])
    P["log_std_W"] += 7
    P["log_var_b"] += 7
    perturb_biases(P["flow_q"] + P["flow_r"], rng)
    eps_z = rng.standard_normal((1, D))
    bq = [(rng.uniform(size=(1, D)) <= 0.5).astype(np.float64) for _ in range(2)]
    br = [(rng.uniform(size=(1, D)) <= 0.5).astype(np.float64) for _ in range(2)]
    scale = 1.0 / 600.0
    if conv:
        Kc = int(np.prod(wshape[:3]))
        eps_a, eps_b = rng.standard_normal(Kc), rng.standard_normal()
        kl, c = O.conv_kl(P, eps_z, eps_a, eps_b, r_term=r_term, bern_q=bq, bern_r=br)
        ref_g = O.conv_kl_backward(P, c, scale, learn_p=True)
        x_shape = (1, 8, 8, wshape[2])
    else:
        eps_a, eps_b = rng.standard_normal(wshape[1]), None
        kl, c = O.dense_kl(P, eps_z, eps_a, r_term=r_term, bern_q=bq, bern_r=br)
        ref_g = O.dense_kl_backward(P, c, scale, learn_p=True)
        x_shape = (1, wshape[0])
    lyr = make_layer(P, x_shape, r_term=r_term, learn_p=True)
    lyr.training = True
    inj = dict(eps_z=_f32(eps_z), eps_a=_f32(eps_a))
    if conv:
        inj["eps_b"] = np.float32(eps_b)
    if flow == "rnvp":
        inj.update(bern_q=[_f32(b) for b in bq], bern_r=[_f32(b) for b in br])
    got = lyr.kl_div(noise=inj)
    close(got.numpy(), np.asarray(kl).reshape(1), RTOL, "kl")
    lyr.zero_grads()
    lyr.kl_backward(scale)
    _check_layer_grads(lyr, ref_g, flow)
