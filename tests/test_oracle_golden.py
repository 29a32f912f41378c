"""Oracle (oracle/mnf_oracle.py) vs the golden vectors produced by running the reference's
own source on the NumPy TF-shim (tests/golden/make_golden.py), float64, rtol 1e-10."""
from __future__ import annotations

import numpy as np
import pytest
from conftest import fix_params, fix_steps, load_golden, split_draws

from oracle import mnf_oracle as O

RT = dict(rtol=1e-10, atol=1e-12)
DENSE = ["dense_iaf", "dense_iaf_clip", "dense_iaf_nf0", "dense_iaf_f3", "dense_noz", "dense_rnvp", "dense_planar"]
CONV = ["conv_iaf", "conv_iaf_clip", "conv_rnvp", "conv_planar"]


def _nq(P):
    return sum(1 for s in P["flow_q"] if s["kind"] == "rnvp")


@pytest.mark.parametrize("name", DENSE + CONV)
def test_layer_call(name):
    fx = load_golden(name)
    P = fix_params(fx["P"])
    use_z = bool(fx["use_z"])
    normals, berns = split_draws(fx["call_draws"])
    eps_z, eps_out = (normals if use_z else [None, normals[0]])
    fn = O.conv_call if name.startswith("conv") else O.dense_call
    y, _ = fn(P, fx["x"], eps_z, eps_out, max_std=float(fx["max_std"]), use_z=use_z, bern_q=berns)
    np.testing.assert_allclose(y, fx["y"], **RT)


@pytest.mark.parametrize("name", DENSE + CONV)
def test_sample_z(name):
    fx = load_golden(name)
    P = fix_params(fx["P"])
    use_z = bool(fx["use_z"])
    normals, berns = split_draws(fx["sz_draws"])
    z, ld, _ = O.sample_z(P, 4, normals[0] if use_z else None, use_z, berns)
    np.testing.assert_allclose(z, fx["sz_z"], **RT)
    np.testing.assert_allclose(ld, fx["sz_ld"], **RT)


@pytest.mark.parametrize("name", DENSE + CONV)
def test_layer_kl(name):
    fx = load_golden(name)
    P = fix_params(fx["P"])
    use_z = bool(fx["use_z"])
    normals, berns = split_draws(fx["kl_draws"])
    nq = _nq(P)
    if not use_z:
        normals = [None, None, None]
    if name.startswith("conv"):
        kl, _ = O.conv_kl(P, normals[0], normals[1], normals[2], use_z=use_z, bern_q=berns[:nq], bern_r=berns[nq:])
    else:
        kl, _ = O.dense_kl(P, normals[0], normals[1], use_z=use_z, bern_q=berns[:nq], bern_r=berns[nq:])
    np.testing.assert_allclose(kl, fx["kl"], **RT)


@pytest.mark.parametrize("name", ["iaf_h8", "iaf_h6_5", "rnvp_h8", "rnvp_h7_4", "planar"])
def test_flow_chain(name):
    fx = load_golden("flows")[name]
    steps = fix_steps(fx["steps"])
    _, berns = split_draws(fx["draws"])
    states, ld, _ = O.flow_forward(fx["z"], steps, berns)
    assert len(states) == len(fx["states"]) == len(steps) + 1
    for a, b in zip(states, fx["states"]):
        np.testing.assert_allclose(a, b, **RT)
    np.testing.assert_allclose(ld, fx["ld"], **RT)


def test_made_masks_vs_reference():
    fx = load_golden("masks")
    for key, val in fx.items():
        D = int(key.split("_")[0][1:])
        h = tuple(int(v) for v in key.split("_h")[1].split("_"))
        masks = O.made_masks(D, h, 2, "autoregressive")
        for a, b in zip(masks, val["intended"]):
            np.testing.assert_array_equal(a, b)
        # as-run (App. B-2): every mask is all-ones after the first call
        ones = O.made_masks(D, h, 2, "ones")
        np.testing.assert_array_equal(val["as_run_sum"], [m.sum() for m in ones])


# SURVEY.md App. A-3 known answers (NumPy RandomState(0), h=(50,))
A3 = {
    20: ([12, 15, 0, 3, 3, 7, 9, 18], 461, 1078),
    50: ([44, 47, 0, 3, 3, 39, 9, 19], 1124, 2752),
    500: ([172, 47, 117, 192, 323, 251, 195, 359], 12082, 25836),
    800: ([684, 559, 629, 192, 763, 707, 359, 9], 23608, 32784),
    136: ([47, 117, 67, 103, 9, 21, 36, 87], 3427, 6746),
    100: ([44, 47, 64, 67, 67, 9, 83, 21], 2648, 4704),
    10: ([5, 0, 3, 3, 7, 3, 5, 2], 245, 510),
    4096: ([2732, 2607, 1653, 3264, 835, 763, 1731, 3431], 108048, 193504),
}


@pytest.mark.parametrize("D", sorted(A3))
def test_made_masks_known_answers(D):
    first, nnz_h, nnz_o = A3[D]
    rng = np.random.RandomState(0)
    m0 = rng.randint(0, D - 1, size=50)
    assert list(m0[:8]) == first
    masks = O.made_masks(D, (50,), 2)
    assert masks[0].shape == (D, 50) and masks[1].shape == (50, 2 * D)
    assert int(masks[0].sum()) == nnz_h and int(masks[1].sum()) == nnz_o


def _lenet_noise(fx):
    normals, _ = split_draws(fx["call_draws"])
    call = [(normals[2 * i], normals[2 * i + 1]) for i in range(4)]
    kn, _ = split_draws(fx["kl_draws"])
    kl = [(kn[0], kn[1], kn[2]), (kn[3], kn[4], kn[5]), (kn[6], kn[7]), (kn[8], kn[9])]
    return call, kl


def test_lenet_small():
    fx = load_golden("lenet_small")
    Ps = [fix_params(P) for P in fx["Ps"]]
    call, kl = _lenet_noise(fx)
    probs = O.lenet_forward(Ps, fx["x"], call, max_std=1.0)
    np.testing.assert_allclose(probs, fx["probs"], **RT)
    np.testing.assert_allclose(O.lenet_kl(Ps, kl), fx["kl"], **RT)


def test_feed_forward_small():
    fx = load_golden("ff_small")
    Ps = [fix_params(P) for P in fx["Ps"]]
    normals, _ = split_draws(fx["call_draws"])
    h = fx["x"]
    for i, P in enumerate(Ps):
        h, _ = O.dense_call(P, h, normals[2 * i], normals[2 * i + 1], max_std=1.0)
        if i < len(Ps) - 1:
            h = O.relu(h) / np.sqrt(1 + 1e-3)  # ReLU then inference-mode BatchNormalization
    np.testing.assert_allclose(h, fx["y"], **RT)
    kn, _ = split_draws(fx["kl_draws"])
    kl = sum(O.dense_kl(P, kn[2 * i], kn[2 * i + 1])[0] for i, P in enumerate(Ps))
    np.testing.assert_allclose(kl, fx["kl"], **RT)
