"""Helpers for the GPU parity tests: build host layers / flows whose parameters come from a
fixture or from the oracle's parameter dicts, and Philox in NumPy (bit-exact integer part)."""
from __future__ import annotations

import numpy as np


def philox4x32_10(c, k):
    """Vectorised Philox4x32-10.  c: uint32 [...,4], k: (k0,k1)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def philox_words(seed64, stream, row_offset, rows, outer, inner):
    """The raw words libmnf_b200 draws for a [rows, outer, inner] tensor (include/mnf_b200.h)."""
    inner4 = (inner + 3) // 4
    r, o, q = np.meshgrid(np.arange(rows, dtype=np.uint64), np.arange(outer, dtype=np.uint64),
                          np.arange(inner4, dtype=np.uint64), indexing="ij")
    grow = r + np.uint64(row_offset)
    c3 = np.uint64(stream) ^ ((grow >> np.uint64(32)) << np.uint64(24))
    c = np.stack([q, o, grow & np.uint64(0xFFFFFFFF), c3 & np.uint64(0xFFFFFFFF)], -1).astype(np.uint32)
    w = philox4x32_10(c, (seed64 & 0xFFFFFFFF, (seed64 >> 32) & 0xFFFFFFFF))
    return w.reshape(rows, outer, inner4 * 4)[:, :, :inner]


def box_muller_f64(words):
    """float64 restatement of csrc/common.cuh:box_muller on [..., 4k] words: the top 23 bits of a word are
    the mantissa of f in [1, 2); u = f - (1 - 2^-24), angle = 2 pi (f - 1.5)."""
    w = words.reshape(*words.shape[:-1], -1, 2).astype(np.uint64)
    fa = 1.0 + (w[..., 0] >> np.uint64(9)).astype(np.float64) / 8388608.0
    fb = 1.0 + (w[..., 1] >> np.uint64(9)).astype(np.float64) / 8388608.0
    u = fa - (1.0 - 2.0 ** -24)
    r = np.sqrt(-2 * np.log(u))
    th = 2 * np.pi * (fb - 1.5)
    return np.stack([r * np.cos(th), r * np.sin(th)], -1).reshape(words.shape)


def assign_flow(fl, st):
    """Copy an oracle-layout flow step (dict of float arrays) into a host flow object."""
    f32 = lambda a: np.asarray(a, np.float32)  # noqa: E731
    kind = st["kind"]
    if kind == "iaf":
        for lyr, W, b in zip(fl.net.layers, st["W"], st["b"]):
            lyr.kernel.assign(f32(W)), lyr.bias.assign(f32(b))
        masks = st.get("masks")
        if masks is not None and not all(np.all(np.asarray(m) == 1) for m in masks):
            for lyr, m in zip(fl.net.layers, masks):
                lyr.set_mask(m, lyr.kernel.ctx)
    elif kind == "rnvp":
        for k, bb, W, b in zip(fl.net_kernels, fl.net_biases, st["Wn"], st["bn"]):
            k.assign(f32(W)), bb.assign(f32(b))
        fl.t_kernel.assign(f32(st["Wt"])), fl.t_bias.assign(f32(st["bt"]))
        fl.s_kernel.assign(f32(st["Ws"])), fl.s_bias.assign(f32(st["bs"]))
    else:
        fl.u.assign(f32(st["u"])), fl.w.assign(f32(st["w"])), fl.b.assign(f32(st["b"]))


def make_flows(steps, D):
    from tf_mnf_b200.flows import IAF, RNVP, PlanarFlow

    out = []
    for st in steps:
        if st["kind"] == "iaf":
            fl = IAF(parity=int(st["parity"]), h_sizes=tuple(W.shape[1] for W in st["W"][:-1]))
        elif st["kind"] == "rnvp":
            fl = RNVP(D, h_sizes=tuple(W.shape[1] for W in st["Wn"]), activation=st.get("act", "tanh"))
        else:
            fl = PlanarFlow(D)
        fl.build(D)
        assign_flow(fl, st)
        out.append(fl)
    return out


def make_layer(P, x_shape, max_std=1.0, use_z=True, r_term="all_states", stride=1, padding="VALID", **kw):
    """Host MNFDense / MNFConv2D carrying the parameters of oracle dict P."""
    from tf_mnf_b200.layers import MNFConv2D, MNFDense

    conv = P["mean_W"].ndim == 4
    fq, fr = P["flow_q"], P["flow_r"]
    kind = (fq or fr or [dict(kind="iaf")])[0]["kind"]
    h = (50,)
    for st in [*fq, *fr]:
        h = (tuple(W.shape[1] for W in st["W"][:-1]) if kind == "iaf" else
             tuple(W.shape[1] for W in st["Wn"]) if kind == "rnvp" else (50,))
    common = dict(n_flows_q=len(fq), n_flows_r=len(fr), use_z=use_z, flow_h_sizes=h, max_std=max_std,
                  flow_type=kind, r_term=r_term, **kw)
    if conv:
        kh, kw_, _, F = P["mean_W"].shape
        lyr = MNFConv2D(F, (kh, kw_), strides=stride, padding=padding, **common)
    else:
        lyr = MNFDense(P["mean_W"].shape[1], **common)
    lyr.build(x_shape)
    for name in ("mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var", "r0_mean", "r0_log_var",
                 "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"):
        if name in P and hasattr(lyr, name):
            getattr(lyr, name).assign(np.asarray(P[name], np.float32))
    for fl, st in zip(lyr.flow_q.flows, fq):
        assign_flow(fl, st)
    for fl, st in zip(lyr.flow_r.flows, fr):
        assign_flow(fl, st)
    return lyr


def close(a, b, rtol=1e-5, what=""):
    """|a-b| <= rtol * max|b| elementwise-scale bound (fp32 parity bar, north_star: 1e-5 relative)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.max(np.abs(b)), 1e-30)
    err = np.max(np.abs(a - b)) / scale
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol:g}"
    return err


# ----------------------------------------------------------------- the timed (in-kernel Philox) mode
def params_from_layer(lyr, dtype=np.float64):
    """Oracle-layout parameter dict read back from a host layer's device variables."""
    P = {}
    for n in ("mean_W", "log_std_W", "mean_b", "log_var_b", "q0_mean", "q0_log_var", "r0_mean", "r0_log_var",
              "r0_apvar", "prior_var_r_p", "prior_var_r_p_bias"):
        if hasattr(lyr, n) and getattr(lyr, n) is not None:
            P[n] = getattr(lyr, n).numpy().astype(dtype)
    for tag in ("flow_q", "flow_r"):
        steps = []
        for fl in getattr(lyr, tag).flows:
            v = {k: a.numpy().astype(dtype) for k, a in fl.variables().items()}
            name = type(fl).__name__
            if name in ("IAF", "MAF"):
                ls = fl.net.layers
                steps.append(dict(kind="iaf", parity=fl.parity, W=[l.kernel.numpy().astype(dtype) for l in ls],
                                  b=[l.bias.numpy().astype(dtype) for l in ls],
                                  masks=None if ls[0].mask is None else [l.mask_host.astype(dtype) for l in ls]))
            elif name == "RNVP":
                steps.append(dict(kind="rnvp", Wn=[k.numpy().astype(dtype) for k in fl.net_kernels],
                                  bn=[b.numpy().astype(dtype) for b in fl.net_biases], Wt=v["t_kernel"],
                                  bt=v["t_bias"], Ws=v["s_kernel"], bs=v["s_bias"], act=fl.activation))
            else:
                steps.append(dict(kind="planar", u=v["u"].reshape(-1, 1), w=v["w"].reshape(-1, 1), b=v["b"].reshape(1)))
        P[tag] = steps
    return P


def draw_normal(lyr, draw, call_index, rows, outer, inner, row_offset=0):
    """Exactly the normals a kernel draws for this layer's (draw id, call index): mnf_philox_normal."""
    import ctypes as C

    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200 import noise as nz

    ctx = lyr.ctx
    out = ctx.empty((rows, outer, inner))
    seed = (int(lyr.seed) & 0xFFFFFFFF) | ((int(call_index) & 0xFFFFFFFF) << 32)
    L.check(ctx.lib.mnf_philox_normal(ctx.handle, seed, nz.stream_id(lyr.uid, draw), int(row_offset), rows, outer, inner,
                                      C.c_void_p(out.ptr)))
    return out.numpy().astype(np.float64)


def draw_bern(lyr, draw, call_index, rows, inner, row_offset=0):
    import ctypes as C

    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200 import noise as nz

    ctx = lyr.ctx
    out = ctx.empty((rows, inner))
    seed = (int(lyr.seed) & 0xFFFFFFFF) | ((int(call_index) & 0xFFFFFFFF) << 32)
    L.check(ctx.lib.mnf_philox_bernoulli(ctx.handle, seed, nz.stream_id(lyr.uid, draw), int(row_offset), rows, inner,
                                         C.c_float(0.5), C.c_void_p(out.ptr)))
    return out.numpy().astype(np.float64)


def call_noise(lyr, B, out_shape, call_index, row_offset=0):
    """(eps_z [B,D], eps_out [B,...], bern_q list) of one call() of `lyr`, as its kernels draw them in
    the timed mode (no injection): draw ids of tf_mnf_b200/noise.py, counters of include/mnf_b200.h."""
    from tf_mnf_b200 import noise as nz

    D = lyr._flow_dim()
    eps_z = draw_normal(lyr, nz.CALL_Z, call_index, B, 1, D, row_offset).reshape(B, D)
    inner = out_shape[-1]
    outer = int(np.prod(out_shape[1:-1])) if len(out_shape) > 2 else 1
    eps_out = draw_normal(lyr, nz.CALL_OUT, call_index, B, outer, inner, row_offset).reshape(out_shape)
    bern = None
    if lyr.flow_type == "rnvp":
        bern = [draw_bern(lyr, nz.BERN_Q_CALL + k, call_index, B, D, row_offset) for k in range(lyr.n_flows_q)]
    return eps_z, eps_out, bern


def kl_noise(lyr, kl_index):
    """Draws of one kl_div() call: (eps_z [1,D], eps_a, eps_b | None, bern_q, bern_r)."""
    from tf_mnf_b200 import noise as nz

    D = lyr._flow_dim()
    rows, cols = lyr._kl_shape()
    eps_z = draw_normal(lyr, nz.KL_Z, kl_index, 1, 1, D).reshape(1, D)
    eps_a = draw_normal(lyr, nz.KL_A, kl_index, 1, 1, rows if lyr._conv else cols).reshape(-1)
    eps_b = float(draw_normal(lyr, nz.KL_B, kl_index, 1, 1, 1).reshape(())) if lyr._conv else None
    bq = br = None
    if lyr.flow_type == "rnvp":
        bq = [draw_bern(lyr, nz.BERN_Q_KL + k, kl_index, 1, D) for k in range(lyr.n_flows_q)]
        br = [draw_bern(lyr, nz.BERN_R_KL + k, kl_index, 1, D) for k in range(lyr.n_flows_r)]
    return eps_z, eps_a, eps_b, bq, br
