"""Oracle self-consistency: analytic backward vs central finite differences
in float64.  The reference has no hand-written gradients (TF autodiff), so this is what
pins the backward formulas the CUDA kernels are later compared against."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import mnf_oracle as O


def leaves(tree, path=()):
    if isinstance(tree, dict):
        for k, v in tree.items():
            yield from leaves(v, (*path, k))
    elif isinstance(tree, list):
        for i, v in enumerate(tree):
            yield from leaves(v, (*path, i))
    elif isinstance(tree, np.ndarray) and tree.dtype == np.float64:
        yield path, tree


def get(tree, path):
    for p in path:
        if tree is None:
            return None
        tree = tree[p] if not isinstance(tree, dict) else tree.get(p)
    return tree


def fd_check(f, P, grads, rng, n=6, eps=1e-6, rtol=2e-5, skip=()):
    for path, arr in leaves(P):
        if any(q in skip for q in path) or "masks" in path:
            continue
        g = get(grads, path)
        assert g is not None, f"missing grad for {path}"
        g = np.asarray(g)
        assert g.shape == arr.shape, (path, g.shape, arr.shape)
        flat = arr.reshape(-1)
        for idx in rng.choice(flat.size, size=min(n, flat.size), replace=False):
            old = flat[idx]
            flat[idx] = old + eps
            fp = f()
            flat[idx] = old - eps
            fm = f()
            flat[idx] = old
            num = (fp - fm) / (2 * eps)
            ana = g.reshape(-1)[idx]
            assert abs(num - ana) <= rtol * max(1.0, abs(num), abs(ana)), (path, idx, num, ana)


def widen(P, rng):
    if "q0_log_var" in P:
        P["q0_log_var"][...] = -1 + 0.3 * rng.standard_normal(P["q0_log_var"].shape)
    P["mean_b"][...] = 0.2 * rng.standard_normal(P["mean_b"].shape)
    for st in P["flow_q"] + P["flow_r"]:
        for k, v in st.items():
            if k.startswith("b"):
                for b in v if isinstance(v, list) else [v]:
                    b[...] = 0.1 * rng.standard_normal(b.shape)


@pytest.mark.parametrize("kind,h", [("iaf", (7,)), ("iaf", (6, 5)), ("rnvp", (7,)), ("rnvp", (5, 4)), ("planar", (1,))])
@pytest.mark.parametrize("masks", ["ones", "autoregressive"])
def test_flow_chain_grad(kind, h, masks):
    if kind != "iaf" and masks == "autoregressive":
        pytest.skip("masks only apply to IAF")
    rng = np.random.RandomState(3)
    D, B, T = 6, 4, 3
    steps = O.make_flow(rng, kind, D, T, h, masks, np.float64)
    for st in steps:
        for k, v in st.items():
            if k.startswith("b"):
                for b in v if isinstance(v, list) else [v]:
                    b[...] = 0.1 * rng.standard_normal(b.shape)
    z = rng.standard_normal((B, D))
    bern = [(rng.uniform(size=(B, D)) <= 0.5).astype(np.float64) for _ in range(T)]
    cs = [rng.standard_normal((B, D)) for _ in range(T + 1)]
    cl = rng.standard_normal(B)
    tree = dict(z=z, steps=steps)

    def f():
        states, ld, _ = O.flow_forward(tree["z"], tree["steps"], bern)
        return sum((c * s).sum() for c, s in zip(cs, states)) + (cl * ld).sum()

    states, ld, caches = O.flow_forward(z, steps, bern)
    dz, gs = O.flow_backward(list(cs), cl, steps, caches)
    fd_check(f, tree, dict(z=dz, steps=gs), rng, skip=("kind", "parity"))


@pytest.mark.parametrize("flow", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("max_std", [1.0, 1.3e-4])
def test_dense_call_grad(flow, max_std):
    rng = np.random.RandomState(5)
    P = O.make_layer_params(rng, (9, 5), flow=flow, h_sizes=(6,), std_init=1.5, dtype=np.float64)
    widen(P, rng)
    B = 4
    x = rng.standard_normal((B, 9))
    eps_z, eps = rng.standard_normal((B, 9)), rng.standard_normal((B, 5))
    bern = [(rng.uniform(size=(B, 9)) <= 0.5).astype(np.float64) for _ in range(2)]
    cy = rng.standard_normal((B, 5))
    tree = dict(P=P, x=x)

    def f():
        return (cy * O.dense_call(tree["P"], tree["x"], eps_z, eps, max_std, bern_q=bern)[0]).sum()

    _, c = O.dense_call(P, x, eps_z, eps, max_std, bern_q=bern)
    dx, g = O.dense_call_backward(P, c, cy)
    fd_check(f, tree, dict(P=g, x=dx), rng,
             skip=("flow_r", "r0_mean", "r0_log_var", "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"))


@pytest.mark.parametrize("flow", ["iaf", "rnvp"])
@pytest.mark.parametrize("padding,stride", [("VALID", 1), ("SAME", 1), ("VALID", 2)])
def test_conv_call_grad(flow, padding, stride):
    rng = np.random.RandomState(6)
    P = O.make_layer_params(rng, (3, 3, 2, 4), flow=flow, h_sizes=(6,), std_init=1.5, dtype=np.float64)
    widen(P, rng)
    B = 3
    x = rng.standard_normal((B, 7, 6, 2))
    y0 = O.conv2d(x, P["mean_W"], stride, padding)
    eps_z, eps = rng.standard_normal((B, 4)), rng.standard_normal(y0.shape)
    bern = [(rng.uniform(size=(B, 4)) <= 0.5).astype(np.float64) for _ in range(2)]
    cy = rng.standard_normal(y0.shape)
    tree = dict(P=P, x=x)
    kw = dict(max_std=1.3e-4, stride=stride, padding=padding, bern_q=bern)

    def f():
        return (cy * O.conv_call(tree["P"], tree["x"], eps_z, eps, **kw)[0]).sum()

    _, c = O.conv_call(P, x, eps_z, eps, **kw)
    dx, g = O.conv_call_backward(P, c, cy)
    fd_check(f, tree, dict(P=g, x=dx), rng,
             skip=("flow_r", "r0_mean", "r0_log_var", "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"))


@pytest.mark.parametrize("flow", ["iaf", "rnvp", "planar"])
@pytest.mark.parametrize("r_term", ["all_states", "last"])
@pytest.mark.parametrize("conv", [False, True])
def test_kl_grad(flow, r_term, conv):
    rng = np.random.RandomState(8)
    shape = (3, 2, 2, 4) if conv else (6, 5)
    P = O.make_layer_params(rng, shape, flow=flow, h_sizes=(6,), std_init=1.0, dtype=np.float64)
    widen(P, rng)
    P["log_std_W"] += 7  # make sigma terms non-negligible
    P["log_var_b"] += 7
    D = 4 if conv else 6
    eps_z = rng.standard_normal((1, D))
    nb = lambda: [(rng.uniform(size=(1, D)) <= 0.5).astype(np.float64) for _ in range(2)]  # noqa: E731
    bq, br = nb(), nb()
    if conv:
        eps_a, eps_b = rng.standard_normal(12), rng.standard_normal()
        fwd = lambda Q: O.conv_kl(Q, eps_z, eps_a, eps_b, r_term=r_term, bern_q=bq, bern_r=br)  # noqa: E731
        bwd = O.conv_kl_backward
    else:
        eps_a = rng.standard_normal(5)
        fwd = lambda Q: O.dense_kl(Q, eps_z, eps_a, r_term=r_term, bern_q=bq, bern_r=br)  # noqa: E731
        bwd = O.dense_kl_backward
    tree = dict(P=P)
    kl, c = fwd(P)
    g = bwd(P, c, 1.0, learn_p=True)
    fd_check(lambda: fwd(tree["P"])[0], tree, dict(P=g), rng, rtol=5e-5)


def test_kl_grad_no_z():
    rng = np.random.RandomState(9)
    P = O.make_layer_params(rng, (6, 5), dtype=np.float64)
    P["log_std_W"] += 8
    kl, c = O.dense_kl(P, None, None, use_z=False)
    g = O.dense_kl_backward(P, c, 1.0, learn_p=True)
    tree = dict(P={k: P[k] for k in ("mean_W", "log_std_W", "mean_b", "log_var_b", "prior_var_r_p", "prior_var_r_p_bias")})
    fd_check(lambda: O.dense_kl(P, None, None, use_z=False)[0], tree, dict(P=g), rng)


@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_backward_fd(training):
    """Stock BatchNormalization of MNFFeedForward (mnf_feed_forward.py:20): the oracle's analytic
    backward against finite differences, batch-statistics (model.fit) and moving-statistics modes."""
    rng = np.random.RandomState(3)
    B, N = 9, 5
    P = dict(x=rng.standard_normal((B, N)), gamma=1 + 0.3 * rng.standard_normal(N), beta=0.2 * rng.standard_normal(N))
    mm, mv = 0.1 * rng.standard_normal(N), 1 + 0.2 * rng.uniform(size=N)
    w = rng.standard_normal((B, N))

    def f():
        y, _, _ = O.batchnorm_forward(P["x"], P["gamma"], P["beta"], mm, mv, training=training)
        return float((y * w).sum())

    y, c, (nm, nv) = O.batchnorm_forward(P["x"], P["gamma"], P["beta"], mm, mv, training=training)
    dx, dg, db = O.batchnorm_backward(w, c)
    fd_check(f, P, dict(x=dx, gamma=dg, beta=db), rng, n=12)
    if training:
        np.testing.assert_allclose(nm, mm * 0.99 + P["x"].mean(0) * 0.01)
        np.testing.assert_allclose(nv, mv * 0.99 + P["x"].var(0) * 0.01)
        np.testing.assert_allclose(y.mean(0), P["beta"], atol=1e-12)
    else:
        assert nm is mm and nv is mv


def test_maf_sequential_decode_inverts_the_made_pass():
    """maf.py:33-45 (oracle: maf_forward_sequential) is the inverse of maf.py:47-53 (iaf_forward = MAF.inverse)
    when the MADE masks are autoregressive -- the property the two directions exist for -- with and without
    the feature reversal, one and two hidden layers; log-determinants are opposite."""
    rng = np.random.RandomState(4)
    for h_sizes in ((9,), (8, 7)):
        for parity in (0, 1):
            D = 6
            st = O.make_flow(rng, "iaf", D, 1, h_sizes, "autoregressive", np.float64)[0]
            st["parity"] = parity
            for b in st["b"]:
                b[...] = 0.1 * rng.standard_normal(b.shape)
            x = rng.standard_normal((5, D))
            y, ld, _ = O.iaf_forward(x, st)
            x2, ld2 = O.maf_forward_sequential(y, st)
            np.testing.assert_allclose(x2, x, atol=1e-10)
            np.testing.assert_allclose(ld2, -ld, atol=1e-10)


@pytest.mark.parametrize("act", ["tanh", "relu", "sigmoid", "linear"])
def test_rnvp_activations_backward_fd(act):
    """RNVP with the constructor's `activation` (rnvp.py:20-29) and two hidden layers: analytic backward vs FD."""
    rng = np.random.RandomState(7)
    D, B = 5, 4
    st = O.make_flow(rng, "rnvp", D, 1, (6, 4), "ones", np.float64)[0]
    st["act"] = act
    for b in [*st["bn"], st["bt"], st["bs"]]:
        b[...] = 0.3 * rng.standard_normal(b.shape)
    z = rng.standard_normal((B, D)) + 0.3
    m = (rng.uniform(size=(B, D)) <= 0.5).astype(np.float64)
    wx, wl = rng.standard_normal((B, D)), rng.standard_normal(B)
    P = dict(z=z, st={k: v for k, v in st.items() if k not in ("kind", "act")})

    def f():
        s2 = dict(P["st"], kind="rnvp", act=act)
        x, ld, _ = O.rnvp_forward(P["z"], s2, m)
        return float((x * wx).sum() + (ld * wl).sum())

    x, ld, c = O.rnvp_forward(z, st, m)
    dz, g = O.rnvp_backward(wx, wl, st, c)
    fd_check(f, P, dict(z=dz, st=g), rng, n=8)
