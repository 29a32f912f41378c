"""Coupling networks beyond the layers' default (one hidden layer, ReLU / tanh) -- made.py:77-88 builds one
MaskedDense per entry of h_sizes, rnvp.py:20-31 one Dense(h, activation) -- and the flow directions the
layers never call: MAF.forward / IAF.inverse (maf.py:33-45), NormalizingFlow.inverse
(flows/__init__.py:31-38), NormalizingFlowModel (:41-57).  CUDA (csrc/flows_generic.cu) against the golden
fixtures generated from the reference's own source and against the float64 oracle; bar 1e-5 of the maximum."""
from __future__ import annotations

import numpy as np
import pytest
from conftest import fix_steps, load_golden, split_draws
from gpu_util import close, make_flows, make_layer
from test_gpu_backward import RTOL_G, flow_grad_map, perturb_biases

from oracle import mnf_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _f32(a):
    return None if a is None else np.asarray(a, np.float32)


@pytest.mark.parametrize("name", ["iaf_h6_5", "rnvp_h7_4"])
def test_multi_hidden_chain_golden(name):
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.flows import NormalizingFlow

    fx = load_golden("flows")[name]
    steps = fix_steps(fx["steps"])
    _, berns = split_draws(fx["draws"])
    nf = NormalizingFlow(make_flows(steps, fx["z"].shape[1]))
    states, ld = nf.forward(fx["z"].astype(np.float32), bern=[b.astype(np.float32) for b in berns] or None)
    assert L.last_path() == "flow_generic", L.last_path()
    assert len(states) == len(fx["states"])
    for k, (a, b) in enumerate(zip(states, fx["states"])):
        close(a.numpy(), b, RTOL, f"{name} state {k}")
    close(ld.numpy(), fx["ld"], RTOL, f"{name} log_dets")


CASES = [("iaf", (16, 12), "ones", None), ("iaf", (9, 8, 7), "autoregressive", None), ("rnvp", (20, 9), None, "tanh"),
         ("rnvp", (11,), None, "relu"), ("rnvp", (10, 6), None, "sigmoid"), ("rnvp", (7, 5, 4, 3), None, "linear")]


@pytest.mark.parametrize("kind,h_sizes,masks,act", CASES)
@pytest.mark.parametrize("D,B", [(24, 37), (130, 5), (7, 1)])
def test_general_networks_forward_backward(kind, h_sizes, masks, act, D, B):
    from tf_mnf_b200.flows import NormalizingFlow

    rng = np.random.RandomState(D + B + len(h_sizes))
    T = 2
    steps = O.make_flow(rng, kind, D, T, h_sizes, masks or "ones", np.float64)
    perturb_biases(steps, rng)
    for st in steps:
        if act:
            st["act"] = act
    z = rng.standard_normal((B, D))
    bern = [(rng.uniform(size=(B, D)) <= 0.5).astype(np.float64) for _ in range(T)]
    cs = [rng.standard_normal((B, D)) for _ in range(T + 1)]
    cl = rng.standard_normal(B)
    ref_states, ref_ld, caches = O.flow_forward(z, steps, bern)
    ref_dz, ref_g = O.flow_backward([c.copy() for c in cs], cl, steps, caches)
    nf = NormalizingFlow(make_flows(steps, D))
    states, ld = nf.forward(_f32(z), bern=[_f32(b) for b in bern] if kind == "rnvp" else None, record=True)
    for k, (a, b) in enumerate(zip(states, ref_states)):
        close(a.numpy(), b, RTOL, f"state {k}")
    close(ld.numpy(), ref_ld, RTOL, "log_dets")
    d_z0, grads = nf.backward(_f32(np.stack(cs)), None, _f32(cl))
    close(d_z0.numpy(), ref_dz, RTOL, "d_z0")
    for k in range(T):
        for name, ref in flow_grad_map(kind, ref_g[k]).items():
            close(grads[k][name].numpy().reshape(np.shape(ref)), ref, RTOL_G, f"step {k} {name}")


@pytest.mark.parametrize("flow,h_sizes", [("iaf", (8, 6)), ("rnvp", (9, 5))])
def test_layer_with_two_hidden_layers(flow, h_sizes):
    """MNFDense(flow_h_sizes=(h1, h2)): call, backward and kl_div through the general kernels vs the oracle."""
    rng = np.random.RandomState(5)
    K, N, B = 30, 12, 21
    P = O.make_layer_params(rng, (K, N), flow=flow, h_sizes=h_sizes, std_init=2.0, dtype=np.float64)
    P["q0_log_var"][...] = -2 + 0.5 * rng.standard_normal(K)
    perturb_biases(P["flow_q"] + P["flow_r"], rng)
    x = rng.standard_normal((B, K))
    eps_z, eps, dy = rng.standard_normal((B, K)), rng.standard_normal((B, N)), rng.standard_normal((B, N))
    bern = [(rng.uniform(size=(B, K)) <= 0.5).astype(np.float64) for _ in range(2)]
    ref, c = O.dense_call(P, x, eps_z, eps, 1.0, bern_q=bern)
    ref_dx, ref_g = O.dense_call_backward(P, c, dy)
    lyr = make_layer(P, x.shape)
    assert lyr.flow_h_sizes == h_sizes
    lyr.training = True
    inj = dict(eps_z=_f32(eps_z), eps_out=_f32(eps), bern_q=[_f32(b) for b in bern] if flow == "rnvp" else None)
    close(lyr(_f32(x), noise=inj).numpy(), ref, RTOL, "call")
    lyr.zero_grads()
    close(lyr.backward(_f32(dy)).numpy(), ref_dx, RTOL, "dx")
    for name in ("mean_W", "log_std_W", "q0_mean", "q0_log_var"):
        close(lyr.grads[name].numpy(), ref_g[name], RTOL_G, name)
    for k, g in enumerate(ref_g["flow_q"]):
        for name, r in flow_grad_map(flow, g).items():
            close(lyr.grads[f"flow_q/{k}/{name}"].numpy().reshape(np.shape(r)), r, RTOL_G, f"flow_q/{k}/{name}")
    eps_z1, eps_a = rng.standard_normal((1, K)), rng.standard_normal(N)
    b1 = [(rng.uniform(size=(1, K)) <= 0.5).astype(np.float64) for _ in range(4)]
    ref_kl, ck = O.dense_kl(P, eps_z1, eps_a, bern_q=b1[:2], bern_r=b1[2:])
    kinj = dict(eps_z=_f32(eps_z1), eps_a=_f32(eps_a))
    if flow == "rnvp":
        kinj.update(bern_q=[_f32(b) for b in b1[:2]], bern_r=[_f32(b) for b in b1[2:]])
    close(lyr.kl_div(noise=kinj).numpy(), np.asarray(ref_kl).reshape(1), RTOL, "kl")
    ref_kg = O.dense_kl_backward(P, ck, 1.0 / 600)
    lyr.zero_grads()
    lyr.kl_backward(1.0 / 600)
    for tag in ("flow_q", "flow_r"):
        for k, g in enumerate(ref_kg[tag]):
            for name, r in flow_grad_map(flow, g).items():
                close(lyr.grads[f"{tag}/{k}/{name}"].numpy().reshape(np.shape(r)), r, RTOL_G, f"kl {tag}/{k}/{name}")


@pytest.mark.parametrize("h_sizes,masks", [((50,), "ones"), ((12, 9), "autoregressive"), ((30,), "autoregressive")])
@pytest.mark.parametrize("D,B,parity", [(20, 33, 0), (64, 7, 1), (5, 1, 1)])
def test_maf_sequential_decode(h_sizes, masks, D, B, parity):
    """MAF.forward / IAF.inverse (maf.py:33-45) against the oracle's loop; with autoregressive masks it
    inverts the single MADE pass (MAF.inverse / IAF.forward) and the log-determinants are opposite."""
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.flows import IAF, MAF

    rng = np.random.RandomState(D * 3 + B)
    st = O.make_flow(rng, "iaf", D, 1, h_sizes, masks, np.float64)[0]
    st["parity"] = parity
    perturb_biases([st], rng)
    z = rng.standard_normal((B, D))
    ref_x, ref_ld = O.maf_forward_sequential(z, st)
    iaf = make_flows([st], D)[0]
    assert isinstance(iaf, IAF)
    x, ld = iaf.inverse(_f32(z))
    assert L.last_path() == "maf_decode"
    close(x.numpy(), ref_x, RTOL, "decode x")
    close(ld.numpy(), ref_ld, RTOL, "decode log_dets")
    maf = MAF(parity, h_sizes=h_sizes, mask_mode=masks)
    maf.build(D)
    from gpu_util import assign_flow

    assign_flow(maf, st)
    x2, ld2 = maf.forward(_f32(z))
    np.testing.assert_array_equal(x2.numpy(), x.numpy())
    if masks == "autoregressive":
        y, ldy = maf.inverse(x2)  # the one-pass direction
        close(y.numpy(), z, 2e-5, "inverse(forward(z)) == z")
        close(ldy.numpy(), -ref_ld, 2e-5, "opposite log-determinants")


def test_normalizing_flow_inverse_and_model():
    """NormalizingFlow.inverse (flows/__init__.py:31-38) over MAF steps and NormalizingFlowModel (:41-57)."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.flows import MAF, RNVP, DiagNormal, NormalizingFlow, NormalizingFlowModel

    M.set_seed(3)
    D, B = 10, 64
    rng = np.random.RandomState(0)
    steps = O.make_flow(rng, "iaf", D, 3, (14,), "autoregressive", np.float64)
    perturb_biases(steps, rng)
    flows = []
    from gpu_util import assign_flow

    for st in steps:
        fl = MAF(int(st["parity"]), h_sizes=(14,), mask_mode="autoregressive")
        fl.build(D)
        assign_flow(fl, st)
        flows.append(fl)
    model = NormalizingFlowModel(DiagNormal(D), flows)
    x = rng.standard_normal((B, D))
    zs, ld = model.inverse(_f32(x))
    # oracle: inverse applies MAF.inverse (= the MADE pass) of the flows in reverse order
    ref, ref_ld = x, np.zeros(B)
    ref_states = [x]
    for st in reversed(steps):
        ref, l, _ = O.iaf_forward(ref, st)
        ref_ld = ref_ld + l
        ref_states.append(ref)
    assert len(zs) == 4
    for k, (a, b) in enumerate(zip(zs, ref_states)):
        close(a.numpy(), b, RTOL, f"inverse state {k}")
    close(ld.numpy(), ref_ld, RTOL, "inverse log_dets")
    lp = model.base_log_prob(_f32(x))
    want = float((-0.5 * ref_states[-1] ** 2 - 0.5 * np.log(2 * np.pi)).sum())
    assert abs(lp - want) <= 1e-5 * abs(want)
    xs = model.sample(500)
    assert len(xs) == 4 and xs[0].shape == (500, D) and np.all(np.isfinite(xs[-1].numpy()))
    assert abs(float(xs[0].numpy().mean())) < 0.1 and abs(float(xs[0].numpy().std()) - 1) < 0.1
    # forward (sequential decode) undoes inverse
    back, _ = NormalizingFlow(flows).forward(zs[-1])
    close(back[-1].numpy(), x, 5e-5, "forward(inverse(x)) == x")
    with pytest.raises(NotImplementedError):
        NormalizingFlow([RNVP(D)]).inverse(_f32(x))
