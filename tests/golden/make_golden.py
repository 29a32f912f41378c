#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN SOURCE (imported unmodified
from /root/reference) on the NumPy TF-shim in tests/golden/tfshim, in float64.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The resulting fixtures are committed; tests never read /root/reference.

Each fixture stores: the layer/flow parameters (in the oracle's dict layout), the inputs,
the ordered list of random draws the reference made (kind, shape, values), and the
reference outputs.  Script-side interventions (documented, none touches arithmetic):
  * PlanarFlow: `flow.dense.build(...)` is called before forward (planar.py:23 reads the
    kernel before Keras has built it).
  * RNVP / planar variants of the layers: the reference layers hard-wire IAF
    (mnf_dense.py:78-86); we swap `layer.flow_q/.flow_r` for NormalizingFlow([RNVP...])
    built from the reference's own flow classes after `build`.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tfshim"))
sys.path.insert(1, "/root/reference")

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)
from tf_mnf.flows import IAF, RNVP, NormalizingFlow, PlanarFlow  # noqa: E402
from tf_mnf.flows.made import MADE, MaskedDense  # noqa: E402
from tf_mnf.layers import MNFConv2D, MNFDense  # noqa: E402
from tf_mnf.models import MNFFeedForward, MNFLeNet  # noqa: E402


def flatten(obj, prefix="", out=None):
    out = {} if out is None else out
    if isinstance(obj, dict):
        for k, v in obj.items():
            flatten(v, f"{prefix}{k}/", out)
    elif isinstance(obj, (list, tuple)):
        out[f"{prefix}__len__"] = np.array(len(obj))
        for i, v in enumerate(obj):
            flatten(v, f"{prefix}{i}/", out)
    elif obj is None:
        out[f"{prefix}__none__"] = np.array(0)
    elif isinstance(obj, str):
        out[f"{prefix}__str__"] = np.array(obj)
    else:
        out[prefix.rstrip("/")] = np.asarray(obj)
    return out


def flow_step_params(fl, as_intended_masks=None):
    if isinstance(fl, IAF):
        mds = [l for l in fl.net.net.layers if isinstance(l, MaskedDense)]
        return dict(kind="iaf", parity=int(fl.parity), W=[np.array(m.kernel) for m in mds],
                    b=[np.array(m.bias) for m in mds], masks=[np.array(m.mask) for m in mds])
    if isinstance(fl, RNVP):
        return dict(kind="rnvp", Wn=[np.array(l.kernel) for l in fl.net.layers],
                    bn=[np.array(l.bias) for l in fl.net.layers], Wt=np.array(fl.t.kernel),
                    bt=np.array(fl.t.bias), Ws=np.array(fl.s.kernel), bs=np.array(fl.s.bias))
    if isinstance(fl, PlanarFlow):
        return dict(kind="planar", u=np.array(fl.u), w=np.array(fl.dense.kernel), b=np.array(fl.dense.bias))
    raise TypeError(fl)


def layer_params(lyr):
    names = ["mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var", "r0_mean", "r0_log_var",
             "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"]
    P = {n: np.array(getattr(lyr, n)) for n in names if hasattr(lyr, n)}
    P["flow_q"] = [flow_step_params(f) for f in lyr.flow_q.flows]
    P["flow_r"] = [flow_step_params(f) for f in lyr.flow_r.flows]
    return P


def perturb_flow_biases(flows, rng):
    """Keras inits biases to zero; give them values so the fixtures exercise them."""
    for fl in flows:
        if isinstance(fl, IAF):
            for m in fl.net.net.layers:
                if isinstance(m, MaskedDense):
                    m.bias[...] = 0.1 * rng.standard_normal(m.bias.shape)
        elif isinstance(fl, RNVP):
            for l in [*fl.net.layers, fl.t, fl.s]:
                l.bias[...] = 0.1 * rng.standard_normal(l.bias.shape)
        elif isinstance(fl, PlanarFlow):
            fl.dense.bias[...] = 0.1 * rng.standard_normal(fl.dense.bias.shape)


def draws():
    return [dict(kind=k, value=v) for k, _, v in tf.NOISE.draws]


def warm_flows(lyr, D):
    """Build flow sub-networks (Keras builds at first call) so their params exist."""
    for nf in (lyr.flow_q, lyr.flow_r):
        for fl in nf.flows:
            if isinstance(fl, PlanarFlow):
                fl.dense.build((None, D))
                fl.dense.built = True
        nf.forward(np.zeros((1, D)))


def swap_flows(lyr, kind, D, h):
    def mk(n):
        if kind == "rnvp":
            return NormalizingFlow([RNVP(D, h_sizes=h) for _ in range(n)])
        return NormalizingFlow([PlanarFlow(D) for _ in range(n)])
    lyr.flow_q, lyr.flow_r = mk(lyr.n_flows_q), mk(lyr.n_flows_r)


def gen_layer(conv, kind, seed, x, max_std=1.0, std_init=1.0, h=(8,), **kw):
    tf.NOISE.reset(seed, seed + 1)
    rng = np.random.RandomState(seed + 2)
    if conv:
        lyr = MNFConv2D(3, 3, max_std=max_std, std_init=std_init, flow_h_sizes=h, **kw)
        D = 3
    else:
        lyr = MNFDense(7, max_std=max_std, std_init=std_init, flow_h_sizes=h, **kw)
        D = x.shape[-1]
    lyr.build(tf.TensorShape(x.shape))
    lyr.built = True
    if kind != "iaf":
        swap_flows(lyr, kind, D, h)
    warm_flows(lyr, D)
    perturb_flow_biases([*lyr.flow_q.flows, *lyr.flow_r.flows], rng)
    lyr.mean_b[...] = 0.3 * rng.standard_normal(lyr.mean_b.shape)
    # make z matter: q0 variance is exp(-9) by default -> widen it
    if lyr.use_z:
        lyr.q0_log_var[...] = -2 + 0.5 * rng.standard_normal(lyr.q0_log_var.shape)
    out = dict(P=layer_params(lyr), x=x, max_std=np.array(max_std), use_z=np.array(int(lyr.use_z)))
    tf.NOISE.reset(seed + 3)
    out["y"] = lyr(x)
    out["call_draws"] = draws()
    tf.NOISE.reset(seed + 4)
    z, ld = lyr.sample_z(4)
    out["sz_z"], out["sz_ld"], out["sz_draws"] = z, ld, draws()
    tf.NOISE.reset(seed + 5)
    out["kl"] = np.array(lyr.kl_div())
    out["kl_draws"] = draws()
    return out


def gen_flows(seed=50):
    out = {}
    rng = np.random.RandomState(seed)
    D, B = 6, 5
    z = rng.standard_normal((B, D))
    cases = {
        "iaf_h8": [IAF(parity=i % 2, h_sizes=(8,)) for i in range(3)],
        "iaf_h6_5": [IAF(parity=i % 2, h_sizes=(6, 5)) for i in range(2)],
        "rnvp_h8": [RNVP(D, h_sizes=(8,)) for _ in range(2)],
        "rnvp_h7_4": [RNVP(D, h_sizes=(7, 4)) for _ in range(2)],
        "planar": [PlanarFlow(D) for _ in range(3)],
    }
    for name, flows in cases.items():
        tf.NOISE.reset(seed + 1, seed + 2 + len(out))
        nf = NormalizingFlow(flows)
        for fl in flows:
            if isinstance(fl, PlanarFlow):
                fl.dense.build((None, D))
                fl.dense.built = True
        nf.forward(np.zeros((1, D)))
        perturb_flow_biases(flows, rng)
        tf.NOISE.reset(seed + 7)
        xs, ld = nf.forward(z)
        out[name] = dict(steps=[flow_step_params(f) for f in flows], z=z, states=list(xs), ld=ld,
                         draws=draws())
    # single-row planar: tf.squeeze turns log_det into a scalar (planar.py:35)
    return out


def gen_masks():
    """Masks as MADE.set_masks constructs them (made.py:93-125), read BEFORE the first
    call overwrites them with ones (made.py:20-23)."""
    out = {}
    for D, h in [(6, (8,)), (20, (50,)), (50, (50,)), (10, (50,)), (12, (6, 5))]:
        made = MADE(n_outputs=2, h_sizes=h)
        made.build(tf.TensorShape((1, D)))
        mds = [l for l in made.net.layers if isinstance(l, MaskedDense)]
        out[f"D{D}_h{'_'.join(map(str, h))}"] = dict(intended=[np.array(m.mask) for m in mds])
        made.built = True
        made(np.zeros((1, D)))
        out[f"D{D}_h{'_'.join(map(str, h))}"]["as_run_sum"] = np.array([float(np.sum(m.mask)) for m in mds])
    return out


def gen_lenet(seed=90):
    tf.NOISE.reset(seed, seed + 1)
    rng = np.random.RandomState(seed + 2)
    model = MNFLeNet(3, 4, 6, 5, max_std=1, flow_h_sizes=[10], std_init=10.0)
    x = rng.uniform(size=(3, 28, 28, 1))
    model(x)  # builds layers + q flows
    model.kl_div()  # builds r flows
    mnf = [l for l in model.layers if hasattr(l, "kl_div")]
    for l in mnf:
        perturb_flow_biases([*l.flow_q.flows, *l.flow_r.flows], rng)
        l.q0_log_var[...] = -3 + 0.5 * rng.standard_normal(l.q0_log_var.shape)
        l.mean_b[...] = 0.1 * rng.standard_normal(l.mean_b.shape)
    out = dict(Ps=[layer_params(l) for l in mnf], x=x)
    tf.NOISE.reset(seed + 3)
    out["probs"] = model(x)
    out["call_draws"] = draws()
    tf.NOISE.reset(seed + 4)
    out["kl"] = np.array(model.kl_div())
    out["kl_draws"] = draws()
    return out


def gen_ff(seed=120):
    tf.NOISE.reset(seed, seed + 1)
    rng = np.random.RandomState(seed + 2)
    model = MNFFeedForward(layer_sizes=(6, 5, 3), flow_h_sizes=[7], std_init=1e-2)
    x = rng.standard_normal((4, 9))
    model(x)
    model.kl_div()
    mnf = [l for l in model.layers if hasattr(l, "kl_div")]
    for l in mnf:
        perturb_flow_biases([*l.flow_q.flows, *l.flow_r.flows], rng)
        l.q0_log_var[...] = -3 + 0.5 * rng.standard_normal(l.q0_log_var.shape)
    out = dict(Ps=[layer_params(l) for l in mnf], x=x)
    tf.NOISE.reset(seed + 3)
    out["y"] = model(x)
    out["call_draws"] = draws()
    tf.NOISE.reset(seed + 4)
    out["kl"] = np.array(model.kl_div())
    out["kl_draws"] = draws()
    return out


def main():
    rng = np.random.RandomState(7)
    xd = rng.standard_normal((5, 12))
    xc = rng.standard_normal((4, 6, 6, 2))
    fix = {
        "dense_iaf": gen_layer(False, "iaf", 10, xd),
        "dense_iaf_clip": gen_layer(False, "iaf", 20, xd, max_std=1.5e-4, std_init=2.0),
        "dense_iaf_nf0": gen_layer(False, "iaf", 25, xd, n_flows_q=0, n_flows_r=0),
        "dense_iaf_f3": gen_layer(False, "iaf", 27, xd, n_flows_q=3, n_flows_r=1, learn_p=True),
        "dense_noz": gen_layer(False, "iaf", 28, xd, use_z=False),
        "dense_rnvp": gen_layer(False, "rnvp", 30, xd),
        "dense_planar": gen_layer(False, "planar", 40, xd),
        "conv_iaf": gen_layer(True, "iaf", 11, xc),
        "conv_iaf_clip": gen_layer(True, "iaf", 21, xc, max_std=1.5e-4, std_init=2.0),
        "conv_rnvp": gen_layer(True, "rnvp", 31, xc),
        "conv_planar": gen_layer(True, "planar", 41, xc),
        "flows": gen_flows(),
        "masks": gen_masks(),
        "lenet_small": gen_lenet(),
        "ff_small": gen_ff(),
    }
    for name, obj in fix.items():
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **flatten(obj))
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
