"""CPU (NumPy) restatement of the tf-mnf MNF-layer hot path.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (tf_mnf_b200) never does and has no CPU fallback.

PARITY PIN: the reference (janosh/tf-mnf) ships no tests or golden vectors and TensorFlow
is not installable here, so parity is *unpinned by the reference's own tests*.  The pin we
do have: tests/golden/make_golden.py executes the reference's unmodified source files on a
NumPy TF-shim (tests/golden/tfshim) in float64 and this module is checked against those
outputs (tests/test_oracle_golden.py), plus the deterministic MADE-mask known answers.

All functions are dtype-generic (pass float64 or float32 arrays) and take every random
draw as an explicit argument (draw order: SURVEY.md App. A-6).

Reference files followed (paths relative to /root/reference):
  tf_mnf/layers/mnf_dense.py:48-170, tf_mnf/layers/mnf_conv.py:55-197,
  tf_mnf/flows/__init__.py:13-38, tf_mnf/flows/maf.py:47-63, tf_mnf/flows/made.py:14-125,
  tf_mnf/flows/rnvp.py:20-47, tf_mnf/flows/planar.py:16-36,
  tf_mnf/models/mnf_lenet.py:14-39, tf_mnf/models/mnf_feed_forward.py:12-28.

Quirk flags; defaults reproduce the reference *as it runs*:
  made_masks = "ones" | "autoregressive"   (B-2: as-run the masks are all ones)
  r_term     = "all_states" | "last"       (B-3: Gaussian r-term summed over all states)
"""
from __future__ import annotations

import numpy as np

LOG2PI = float(np.log(2 * np.pi))


# ------------------------------------------------------------------ MADE masks (A-3)
def made_masks(n_in: int, h_sizes, n_outputs: int = 2, mode: str = "autoregressive"):
    """made.py:93-125 (shuffle=False, num_masks=1, seed 0).  mode="ones" gives what is in
    effect at run time (made.py:20-23 overwrites the masks at first call; App. B-2)."""
    h_sizes = tuple(int(h) for h in h_sizes)
    shapes = []
    prev = n_in
    for h in h_sizes:
        shapes.append((prev, h))
        prev = h
    shapes.append((prev, n_in * n_outputs))
    if mode == "ones":
        return [np.ones(s) for s in shapes]
    rng = np.random.RandomState(0)  # made.py:75,98
    m = {-1: np.arange(n_in)}  # made.py:102-104
    for lyr, size in enumerate(h_sizes):
        m[lyr] = rng.randint(m[lyr - 1].min(), n_in - 1, size=size)  # made.py:108
    L = len(h_sizes)
    masks = [m[lyr - 1][:, None] <= m[lyr][None, :] for lyr in range(L)]  # made.py:112-114
    masks.append(m[L - 1][:, None] < m[-1][None, :])  # made.py:115
    if n_outputs > 1:
        masks[-1] = np.concatenate([masks[-1]] * n_outputs, axis=1)  # made.py:120
    return [mk.astype(np.float64) for mk in masks]


# ------------------------------------------------------------------------ flow steps
def iaf_forward(x, st):
    """IAF.forward == MAF.inverse, maf.py:47-53,63; MADE made.py:77-91; MaskedDense :28-30."""
    W, b, masks = st["W"], st["b"], st.get("masks")
    acts, pres = [x], []
    h = x
    for l in range(len(W)):
        We = W[l] if masks is None else W[l] * masks[l].astype(W[l].dtype)
        pre = h @ We + b[l]
        pres.append(pre)
        h = np.maximum(pre, 0) if l < len(W) - 1 else pre
        acts.append(h)
    D = x.shape[1]
    s, t = h[:, :D], h[:, D:]  # tf.split order, made.py:91 / maf.py:49
    es = np.exp(s)
    y = x * es + t
    if st["parity"]:
        y = y[:, ::-1]
    ld = s.sum(axis=1)
    return y, ld, dict(x=x, es=es, acts=acts, pres=pres)


def iaf_backward(dy, dld, st, c):
    W, masks = st["W"], st.get("masks")
    x, es = c["x"], c["es"]
    if st["parity"]:
        dy = dy[:, ::-1]
    ds = dy * x * es + dld[:, None]
    do = np.concatenate([ds, dy], axis=1)
    dx = dy * es
    gW, gb = [None] * len(W), [None] * len(W)
    d = do
    for l in reversed(range(len(W))):
        mk = None if masks is None else masks[l].astype(W[l].dtype)
        We = W[l] if mk is None else W[l] * mk
        gW[l] = c["acts"][l].T @ d
        if mk is not None:
            gW[l] = gW[l] * mk
        gb[l] = d.sum(axis=0)
        d = d @ We.T
        if l > 0:
            d = d * (c["pres"][l - 1] > 0)
    dx = dx + d
    return dx, dict(W=gW, b=gb)


_ACT = {
    "tanh": (np.tanh, lambda h: 1 - h * h),
    "relu": (lambda x: np.maximum(x, 0), lambda h: (h > 0).astype(h.dtype)),
    "sigmoid": (lambda x: 1.0 / (1.0 + np.exp(-x)), lambda h: h * (1 - h)),
    "linear": (lambda x: x, lambda h: np.ones_like(h)),
}


def rnvp_forward(z, st, m):
    """rnvp.py:33-47.  m = Bernoulli(0.5) mask [B,D], fresh per call (rnvp.py:36).  st["act"]: the
    constructor's `activation` of the net's Dense layers (rnvp.py:20-29; default "tanh")."""
    act = _ACT[st.get("act", "tanh")][0]
    z1, z2 = (1 - m) * z, m * z
    h, hs = z2, []
    for Wn, bn in zip(st["Wn"], st["bn"]):
        h = act(h @ Wn + bn)
        hs.append(h)
    sh = h @ st["Wt"] + st["bt"]
    sc = h @ st["Ws"] + st["bs"]
    g = 1.0 / (1.0 + np.exp(-sc))
    ld = ((1 - m) * np.log(g)).sum(axis=1)
    x = (z1 * g + (1 - g) * sh) + z2
    return x, ld, dict(z=z, m=m, hs=hs, sh=sh, g=g)


def maf_forward_sequential(z, st):
    """MAF.forward == IAF.inverse, maf.py:33-45, exactly as the reference spells the loop out (it cannot
    run there: it assigns into a tensor slice, App. B-4): x = 0; z reversed when parity; for i in
    range(dim): (s, t) = net(x); x[:, i] = (z[:, i] - t[:, i]) exp(-s[:, i]); log_dets += -s[:, i].
    One full MADE pass per dimension (the slow direction)."""
    W, b, masks = st["W"], st["b"], st.get("masks")
    B, D = z.shape
    x = np.zeros_like(z)
    ld = np.zeros(B, z.dtype)
    zz = z[:, ::-1] if st["parity"] else z
    for i in range(D):
        h = x
        for l in range(len(W)):
            We = W[l] if masks is None else W[l] * masks[l].astype(W[l].dtype)
            h = h @ We + b[l]
            if l < len(W) - 1:
                h = np.maximum(h, 0)
        s, t = h[:, :D], h[:, D:]
        x[:, i] = (zz[:, i] - t[:, i]) * np.exp(-s[:, i])
        ld = ld - s[:, i]
    return x, ld


def rnvp_backward(dx, dld, st, c):
    z, m, hs, sh, g = c["z"], c["m"], c["hs"], c["sh"], c["g"]
    z1 = (1 - m) * z
    dg = dx * (z1 - sh) + dld[:, None] * (1 - m) / g
    dsc = dg * g * (1 - g)
    dsh = dx * (1 - g)
    y = hs[-1]
    grads = dict(Wt=y.T @ dsh, bt=dsh.sum(0), Ws=y.T @ dsc, bs=dsc.sum(0), Wn=[], bn=[])
    dh = dsh @ st["Wt"].T + dsc @ st["Ws"].T
    gWn, gbn = [None] * len(hs), [None] * len(hs)
    dact = _ACT[st.get("act", "tanh")][1]
    for l in reversed(range(len(hs))):
        dp = dh * dact(hs[l])
        inp = hs[l - 1] if l > 0 else m * z
        gWn[l] = inp.T @ dp
        gbn[l] = dp.sum(0)
        dh = dp @ st["Wn"][l].T
    grads["Wn"], grads["bn"] = gWn, gbn
    dz = dx * ((1 - m) * g + m) + m * dh
    return dz, grads


def _softplus(v):
    return np.logaddexp(0.0, v)


def planar_forward(z, st):
    """planar.py:22-36.  u,w: [D,1]; b: [1]."""
    u, w, b = st["u"], st["w"], st["b"]
    uw = (u * w).sum()
    suw = -1 + _softplus(uw)
    ww = (w**2).sum()
    u_hat = u + (suw - uw) * w / ww
    a = z @ w + b  # [B,1]
    th = np.tanh(a)
    x = z + th @ u_hat.T
    psi = (1 - th**2) @ w.T
    psi_u = psi @ u_hat  # [B,1]
    ld = np.log(np.abs(1 + psi_u))[:, 0]  # tf.squeeze -> [B] (scalar when B == 1)
    return x, ld, dict(z=z, th=th, u_hat=u_hat, uw=uw, ww=ww, suw=suw, psi_u=psi_u)


def planar_backward(dx, dld, st, c):
    u, w = st["u"], st["w"]
    z, th, u_hat, uw, ww, suw, psi_u = (c[k] for k in ("z", "th", "u_hat", "uw", "ww", "suw", "psi_u"))
    wu = (w * u_hat).sum()  # = w^T u_hat
    dth2 = 1 - th**2  # [B,1]
    # ld = log|1 + dth2 * wu|
    dq = (dld[:, None] / (1 + psi_u))  # d/d(psi_u)
    dwu = (dq * dth2).sum()
    d_dth2 = dq * wu
    # x = z + th u_hat^T
    dth = dx @ u_hat + d_dth2 * (-2 * th)
    du_hat = dx.T @ th + dwu * w
    dw = dwu * u_hat
    da = dth * dth2
    dz = dx + da @ w.T
    dw = dw + z.T @ da
    db = da.sum(0)
    # u_hat = u + (suw - uw) w / ww
    k = (suw - uw) / ww
    du = du_hat.copy()
    dw = dw + k * du_hat
    dk = (du_hat * w).sum()
    dsuw_m_uw = dk / ww
    dww = -dk * (suw - uw) / ww**2
    sig = 1.0 / (1.0 + np.exp(-uw))
    duw = dsuw_m_uw * (sig - 1)
    du = du + duw * w
    dw = dw + duw * u + dww * 2 * w
    return dz, dict(u=du, w=dw, b=db)


def flow_step_forward(z, st, m=None):
    k = st["kind"]
    if k == "iaf":
        return iaf_forward(z, st)
    if k == "rnvp":
        return rnvp_forward(z, st, m)
    if k == "planar":
        return planar_forward(z, st)
    raise ValueError(k)


def flow_step_backward(dy, dld, st, c):
    return {"iaf": iaf_backward, "rnvp": rnvp_backward, "planar": planar_backward}[st["kind"]](dy, dld, st, c)


def flow_forward(z, steps, bern=None):
    """NormalizingFlow.forward, flows/__init__.py:22-29: returns the LIST of T+1 states."""
    ld = np.zeros(z.shape[0], z.dtype)
    states, caches = [z], []
    bi = 0
    for st in steps:
        m = None
        if st["kind"] == "rnvp":
            m = bern[bi]
            bi += 1
        z, l, c = flow_step_forward(z, st, m)
        ld = ld + l
        states.append(z)
        caches.append(c)
    return states, ld, caches


def flow_backward(dstates, dld, steps, caches):
    """dstates: list of T+1 grads (None allowed).  Returns dz0, [grads per step]."""
    T = len(steps)
    d = dstates[T]
    if d is None:
        d = 0.0
    grads = [None] * T
    for k in reversed(range(T)):
        d = d + np.zeros_like(caches[k].get("x", caches[k].get("z")))
        d, grads[k] = flow_step_backward(d, dld, steps[k], caches[k])
        if dstates[k] is not None:
            d = d + dstates[k]
    return d, grads


# ----------------------------------------------------------------------- sample_z (A-2)
def sample_z(P, B, eps_z, use_z=True, bern_q=None):
    """mnf_dense.py:88-102 / mnf_conv.py:100-114.  Returns z [B,D], log_dets [B], cache."""
    dt = P["mean_W"].dtype
    if not use_z:
        W = P["mean_W"]
        D = W.shape[-1] if W.ndim == 4 else W.shape[0]
        return np.ones((B, D), dt), np.zeros(B, dt), None
    sd = np.exp(0.5 * P["q0_log_var"])  # sqrt(exp(.)) :95-96
    z0 = P["q0_mean"][None, :] + sd[None, :] * eps_z
    states, ld, caches = flow_forward(z0, P["flow_q"], bern_q)
    return states[-1], ld, dict(eps_z=eps_z, sd=sd, states=states, caches=caches)


def sample_z_backward(P, c, dz, dld):
    T = len(P["flow_q"])
    dstates = [None] * T + [dz]
    dz0, gflow = flow_backward(dstates, dld, P["flow_q"], c["caches"])
    g = dict(q0_mean=dz0.sum(0), q0_log_var=(dz0 * c["eps_z"] * 0.5 * c["sd"][None, :]).sum(0), flow_q=gflow)
    return g


# -------------------------------------------------------------------- MNFDense.call (A-1)
def _clipped_var(P, max_std):
    eW = np.exp(P["log_std_W"])
    std = np.minimum(eW, max_std)  # clip_by_value(.,0,max_std) :163
    eb = np.exp(P["log_var_b"])
    vb = np.minimum(eb, max_std**2)  # :164
    return std * std, vb, eW <= max_std, eb <= max_std**2


def dense_call(P, x, eps_z, eps_out, max_std=1.0, use_z=True, bern_q=None):
    """mnf_dense.py:159-170."""
    z, _, zc = sample_z(P, x.shape[0], eps_z, use_z, bern_q)
    A = x * z
    mu = A @ P["mean_W"] + P["mean_b"]
    Vw, vb, inW, inb = _clipped_var(P, max_std)
    V = (x * x) @ Vw + vb
    sq = np.sqrt(V)
    y = mu + sq * eps_out
    return y, dict(x=x, z=z, zc=zc, A=A, Vw=Vw, vb=vb, inW=inW, inb=inb, sq=sq, eps=eps_out)


def dense_call_backward(P, c, g):
    x, z, A, Vw, sq, eps = c["x"], c["z"], c["A"], c["Vw"], c["sq"], c["eps"]
    dV = g * eps / (2 * sq)
    dA = g @ P["mean_W"].T
    dz = dA * x
    dx = dA * z + 2 * x * (dV @ Vw.T)
    dVw = (x * x).T @ dV
    grads = dict(
        mean_W=A.T @ g,
        mean_b=g.sum(0),
        log_std_W=np.where(c["inW"], 2 * Vw * dVw, 0),
        log_var_b=np.where(c["inb"], c["vb"] * dV.sum(0), 0),
    )
    if c["zc"] is not None:
        grads.update(sample_z_backward(P, c["zc"], dz, np.zeros(x.shape[0], x.dtype)))
    return dx, grads


# ------------------------------------------------------------------ MNFConv2D.call (A-1c)
def _pad_same(x, kh, kw, s):
    B, H, W, C = x.shape
    Ho, Wo = -(-H // s), -(-W // s)
    ph = max((Ho - 1) * s + kh - H, 0)
    pw = max((Wo - 1) * s + kw - W, 0)
    pads = ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0))
    return np.pad(x, pads), pads


def _im2col(x, kh, kw, s):
    B, H, W, C = x.shape
    Ho, Wo = (H - kh) // s + 1, (W - kw) // s + 1
    sb, sh, sw, sc = x.strides
    v = np.lib.stride_tricks.as_strided(
        x, (B, Ho, Wo, kh, kw, C), (sb, sh * s, sw * s, sh, sw, sc), writeable=False
    )
    return v.reshape(B * Ho * Wo, kh * kw * C), Ho, Wo


def conv2d(x, w, stride=1, padding="VALID"):
    kh, kw, C, F = w.shape
    if padding == "SAME":
        x, _ = _pad_same(x, kh, kw, stride)
    cols, Ho, Wo = _im2col(np.ascontiguousarray(x), kh, kw, stride)
    return (cols @ w.reshape(-1, F)).reshape(x.shape[0], Ho, Wo, F)


def _conv_bwd(x, w, dout, stride, padding):
    """Returns dx, dw for out = conv2d(x, w)."""
    kh, kw, C, F = w.shape
    pads = None
    xp = x
    if padding == "SAME":
        xp, pads = _pad_same(x, kh, kw, stride)
    xp = np.ascontiguousarray(xp)
    cols, Ho, Wo = _im2col(xp, kh, kw, stride)
    d2 = dout.reshape(-1, F)
    dw = (cols.T @ d2).reshape(w.shape)
    dcols = (d2 @ w.reshape(-1, F).T).reshape(x.shape[0], Ho, Wo, kh, kw, C)
    dxp = np.zeros_like(xp)
    for i in range(kh):
        for j in range(kw):
            dxp[:, i : i + (Ho - 1) * stride + 1 : stride, j : j + (Wo - 1) * stride + 1 : stride, :] += dcols[:, :, :, i, j, :]
    if pads is not None:
        (_, _), (t, bt), (l, r), _ = pads
        dxp = dxp[:, t : dxp.shape[1] - bt, l : dxp.shape[2] - r, :]
    return dxp, dw


def conv_call(P, x, eps_z, eps_out, max_std=1.0, use_z=True, stride=1, padding="VALID", bern_q=None):
    """mnf_conv.py:182-197 (with the strides/padding the call omits, App. B-1)."""
    z, _, zc = sample_z(P, x.shape[0], eps_z, use_z, bern_q)
    Vw, vb, inW, inb = _clipped_var(P, max_std)
    m = conv2d(x, P["mean_W"], stride, padding) + P["mean_b"]
    v = conv2d(x * x, Vw, stride, padding) + vb
    sq = np.sqrt(v)
    y = m * z[:, None, None, :] + sq * eps_out
    return y, dict(x=x, z=z, zc=zc, m=m, Vw=Vw, vb=vb, inW=inW, inb=inb, sq=sq, eps=eps_out,
                   stride=stride, padding=padding)


def conv_call_backward(P, c, g):
    x, z, m, Vw, sq, eps = c["x"], c["z"], c["m"], c["Vw"], c["sq"], c["eps"]
    dm = g * z[:, None, None, :]
    dz = (g * m).sum(axis=(1, 2))
    dv = g * eps / (2 * sq)
    dx1, dW = _conv_bwd(x, P["mean_W"], dm, c["stride"], c["padding"])
    dx2, dVw = _conv_bwd(x * x, Vw, dv, c["stride"], c["padding"])
    dx = dx1 + 2 * x * dx2
    grads = dict(
        mean_W=dW,
        mean_b=dm.sum(axis=(0, 1, 2)),
        log_std_W=np.where(c["inW"], 2 * Vw * dVw, 0),
        log_var_b=np.where(c["inb"], c["vb"] * dv.sum(axis=(0, 1, 2)), 0),
    )
    if c["zc"] is not None:
        grads.update(sample_z_backward(P, c["zc"], dz, np.zeros(x.shape[0], x.dtype)))
    return dx, grads


# ------------------------------------------------------------------------ kl_div (A-4/5)
def _r_gauss(states, mean_r, lvr, r_term):
    """mnf_dense.py:151-155: as written the sum runs over ALL stacked states (App. B-3)."""
    S = np.stack(states, 0) if r_term == "all_states" else states[-1][None]
    d = S - mean_r
    return 0.5 * np.sum(-np.exp(lvr) * d * d - LOG2PI + lvr), S, d


def _kl_common_backward(P, c, G, dz_direct, dc_direct):
    """Shared tail of dense/conv KL backward: Gaussian r-term, flow_r, flow_q."""
    grads = {}
    abar, d, lvr = c["abar"], c["d"], c["lvr"]
    e = np.exp(lvr)
    # upstream on L (Gaussian term) is -G  (KL = ... - log_r)
    dS = -G * (-e * d)  # [S,1,D]
    dmean_r = -G * (e * d).sum(axis=(0, 1))
    dlvr = -G * 0.5 * (-e * d * d + 1).sum(axis=(0, 1))
    grads["r0_mean"] = dmean_r * abar
    grads["r0_log_var"] = dlvr * abar
    dabar = (dmean_r * P["r0_mean"] + dlvr * P["r0_log_var"]).sum()
    T = len(P["flow_r"])
    if c["r_term"] == "all_states":
        dstates = [dS[k] for k in range(T + 1)]
    else:
        dstates = [None] * T + [dS[0]]
    dldr = np.full(1, -G, d.dtype)
    dz_r, gfr = flow_backward(dstates, dldr, P["flow_r"], c["rcaches"])
    grads["flow_r"] = gfr
    return grads, dabar, dz_r


def dense_kl(P, eps_z, eps_a, use_z=True, r_term="all_states", bern_q=None, bern_r=None):
    """mnf_dense.py:104-157."""
    z, ldq, zc = sample_z(P, 1, eps_z, use_z, bern_q)
    Wm, Ls = P["mean_W"], P["log_std_W"]
    Mt = z.T * Wm  # :107
    Vt = np.exp(Ls) ** 2  # :108 (unclipped, App. B-10)
    p = P["prior_var_r_p"]
    Pi = np.exp(p)[:, None]
    klw = 0.5 * np.sum(p[:, None] - 2 * Ls + (Vt + Mt**2) / Pi - 1)  # :112-117
    pb = P["prior_var_r_p_bias"]
    lvb = P["log_var_b"]
    klb = 0.5 * np.sum(pb - lvb + (np.exp(lvb) + P["mean_b"] ** 2) / np.exp(pb) - 1)  # :118-124
    c = dict(z=z, zc=zc, Mt=Mt, Vt=Vt, use_z=use_z, r_term=r_term)
    if not use_z:
        return klw + klb, c
    log_q = -ldq[0] - 0.5 * np.sum(LOG2PI + P["q0_log_var"] + 1)  # :126-130
    states, ldr, rcaches = flow_forward(z, P["flow_r"], bern_r)  # :134-136
    log_r = ldr[0]
    cap = P["r0_apvar"]
    mw = Mt.T @ cap  # :139
    vw = Vt.T @ (cap**2)  # :140
    sv = np.sqrt(vw)
    a = np.tanh(mw + sv * eps_a)  # :144
    abar = a.mean()
    mean_r = abar * P["r0_mean"]  # :146
    lvr = abar * P["r0_log_var"]  # :147
    L, S, d = _r_gauss(states, mean_r, lvr, r_term)
    log_r = log_r + L
    c.update(a=a, abar=abar, lvr=lvr, d=d, sv=sv, eps_a=eps_a, rcaches=rcaches, states=states)
    return klw + klb - log_r + log_q, c


def dense_kl_backward(P, c, G=1.0, learn_p=False):
    z, Mt, Vt = c["z"], c["Mt"], c["Vt"]
    Wm = P["mean_W"]
    p = P["prior_var_r_p"]
    Pi = np.exp(p)[:, None]
    pb = P["prior_var_r_p_bias"]
    lvb = P["log_var_b"]
    g = {}
    dW = G * Mt * z.T / Pi
    dLs = G * (-1 + Vt / Pi)
    dz = G * (Mt * Wm / Pi).sum(1)[None, :]
    g["mean_b"] = G * P["mean_b"] / np.exp(pb)
    g["log_var_b"] = G * 0.5 * (-1 + np.exp(lvb) / np.exp(pb))
    if learn_p:
        g["prior_var_r_p"] = G * 0.5 * (1 - (Vt + Mt**2) / Pi).sum(1)
        g["prior_var_r_p_bias"] = G * 0.5 * np.sum(1 - (np.exp(lvb) + P["mean_b"] ** 2) / np.exp(pb), keepdims=True)
    if c["use_z"]:
        cap = P["r0_apvar"]
        gc, dabar, dz_r = _kl_common_backward(P, c, G, None, None)
        g.update(gc)
        a = c["a"]
        dpre = dabar / a.size * (1 - a * a)
        dmw = dpre
        dvw = dpre * c["eps_a"] / (2 * c["sv"])
        dW = dW + z.T * cap[:, None] * dmw[None, :]
        dLs = dLs + 2 * Vt * (cap**2)[:, None] * dvw[None, :]
        dz = dz + ((Wm * dmw[None, :]).sum(1) * cap)[None, :]
        g["r0_apvar"] = Mt @ dmw + 2 * cap * (Vt @ dvw)
        dz = dz + dz_r
        gq = sample_z_backward(P, c["zc"], dz, np.full(1, -G, z.dtype))
        gq["q0_log_var"] = gq["q0_log_var"] - 0.5 * G
        g.update(gq)
    g["mean_W"], g["log_std_W"] = dW, dLs
    return g


def conv_kl(P, eps_z, eps_a, eps_b, use_z=True, r_term="all_states", bern_q=None, bern_r=None):
    """mnf_conv.py:116-180."""
    z, ldq, zc = sample_z(P, 1, eps_z, use_z, bern_q)
    F = P["mean_W"].shape[-1]
    Ls = P["log_std_W"].reshape(-1, F)
    Wm = P["mean_W"].reshape(-1, F)  # :120-121
    Mt = Wm * z  # :122
    bt = P["mean_b"] * z[0]  # :123
    Vt = np.exp(Ls) ** 2
    p = P["prior_var_r_p"]
    Pi = np.exp(p)[:, None]
    klw = 0.5 * np.sum(p[:, None] - Ls + (Vt + Mt**2) / Pi - 1)  # :128-133 (-log sigma, B-6)
    pb = P["prior_var_r_p_bias"]
    lvb = P["log_var_b"]
    klb = 0.5 * np.sum(pb - lvb + (np.exp(lvb) + bt**2) / np.exp(pb) - 1)  # :134-140
    c = dict(z=z, zc=zc, Mt=Mt, Vt=Vt, bt=bt, use_z=use_z, r_term=r_term)
    if not use_z:
        return klw + klb, c
    log_q = -ldq[0] - 0.5 * np.sum(LOG2PI + P["q0_log_var"] + 1)
    states, ldr, rcaches = flow_forward(z, P["flow_r"], bern_r)
    log_r = ldr[0]
    cap = P["r0_apvar"]
    mw = Mt @ cap  # :154
    vw = Vt @ (cap**2)  # :155
    sv = np.sqrt(vw)
    a = mw + sv * eps_a  # :161 (no tanh)
    mu_b = np.sum(bt * cap)  # :163
    var_b = np.sum(np.exp(lvb) * cap**2)  # :164
    svb = np.sqrt(var_b)
    a = a + mu_b + svb * eps_b  # :165
    abar = a.mean()
    mean_r = abar * P["r0_mean"]
    lvr = abar * P["r0_log_var"]
    L, S, d = _r_gauss(states, mean_r, lvr, r_term)
    log_r = log_r + L
    c.update(a=a, abar=abar, lvr=lvr, d=d, sv=sv, svb=svb, eps_a=eps_a, eps_b=eps_b, rcaches=rcaches,
             states=states)
    return klw + klb - log_r + log_q, c


def conv_kl_backward(P, c, G=1.0, learn_p=False):
    z, Mt, Vt, bt = c["z"], c["Mt"], c["Vt"], c["bt"]
    F = P["mean_W"].shape[-1]
    Wm = P["mean_W"].reshape(-1, F)
    p = P["prior_var_r_p"]
    Pi = np.exp(p)[:, None]
    pb = P["prior_var_r_p_bias"]
    epb = np.exp(pb)
    lvb = P["log_var_b"]
    g = {}
    dW = G * Mt * z / Pi
    dLs = G * (-0.5 + Vt / Pi)
    dz = G * (Mt * Wm / Pi).sum(0)[None, :]
    dbt = G * bt / epb
    dlvb = G * 0.5 * (-1 + np.exp(lvb) / epb)
    if learn_p:
        g["prior_var_r_p"] = G * 0.5 * (1 - (Vt + Mt**2) / Pi).sum(1)
        g["prior_var_r_p_bias"] = G * 0.5 * np.sum(1 - (np.exp(lvb) + bt**2) / epb, keepdims=True)
    if c["use_z"]:
        cap = P["r0_apvar"]
        gc, dabar, dz_r = _kl_common_backward(P, c, G, None, None)
        g.update(gc)
        n = c["a"].size
        da = np.full(n, dabar / n, z.dtype)
        dmw = da
        dvw = da * c["eps_a"] / (2 * c["sv"])
        dmu_b = da.sum()
        dvar_b = da.sum() * c["eps_b"] / (2 * c["svb"])
        dW = dW + dmw[:, None] * (z * cap[None, :])
        dLs = dLs + 2 * Vt * dvw[:, None] * (cap**2)[None, :]
        dz = dz + ((Wm * dmw[:, None]).sum(0) * cap)[None, :]
        dbt = dbt + dmu_b * cap
        dlvb = dlvb + dvar_b * np.exp(lvb) * cap**2
        g["r0_apvar"] = Mt.T @ dmw + 2 * cap * (Vt.T @ dvw) + dmu_b * bt + dvar_b * 2 * cap * np.exp(lvb)
        dz = dz + dz_r
    g["mean_b"] = dbt * z[0]
    dz = dz + (dbt * P["mean_b"])[None, :]
    g["log_var_b"] = dlvb
    if c["use_z"]:
        gq = sample_z_backward(P, c["zc"], dz, np.full(1, -G, z.dtype))
        gq["q0_log_var"] = gq["q0_log_var"] - 0.5 * G
        g.update(gq)
    g["mean_W"] = dW.reshape(P["mean_W"].shape)
    g["log_std_W"] = dLs.reshape(P["mean_W"].shape)
    return g


# ---------------------------------------------------------------- stock layers + models
def relu(x):
    return np.maximum(x, 0)


def maxpool2x2_same(x):
    """keras MaxPool2D(padding="SAME") defaults: 2x2 window, stride 2 (mnf_lenet.py:24,27)."""
    B, H, W, C = x.shape
    Ho, Wo = -(-H // 2), -(-W // 2)
    xp = np.full((B, 2 * Ho, 2 * Wo, C), -np.inf, x.dtype)
    xp[:, :H, :W, :] = x
    return xp.reshape(B, Ho, 2, Wo, 2, C).max(axis=(2, 4))


def softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


# ------------------------------------------------ stock BatchNormalization (mnf_feed_forward.py:20)
def batchnorm_forward(x, gamma, beta, moving_mean, moving_var, eps=1e-3, momentum=0.99, training=True):
    """tf.keras.layers.BatchNormalization() on [B, N] (Keras defaults; third-party: tensorflow==2.1.0
    keras/layers/normalization.py, non-fused path for rank-2 inputs).  training=True is what
    model.fit runs (bandgap_regression.py:76): batch mean / biased variance normalise and the moving
    statistics move by (1 - momentum) towards them; training=False (model(x), :83,:107) uses the
    moving statistics.  Returns y, cache, (new_moving_mean, new_moving_var)."""
    if training:
        mean = x.mean(axis=0)
        var = ((x - mean) ** 2).mean(axis=0)
        new = (moving_mean * momentum + mean * (1 - momentum), moving_var * momentum + var * (1 - momentum))
    else:
        mean, var, new = moving_mean, moving_var, (moving_mean, moving_var)
    inv = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * inv
    return xhat * gamma + beta, dict(xhat=xhat, inv=inv, gamma=gamma, training=training), new


def batchnorm_backward(dy, c):
    """Gradient of batchnorm_forward: (dx, dgamma, dbeta)."""
    xhat, inv, gamma = c["xhat"], c["inv"], c["gamma"]
    dgamma, dbeta = (dy * xhat).sum(axis=0), dy.sum(axis=0)
    if c["training"]:
        B = dy.shape[0]
        dx = gamma * inv / B * (B * dy - dbeta - xhat * dgamma)
    else:
        dx = dy * gamma * inv
    return dx, dgamma, dbeta


def lenet_forward(Ps, x, noise, max_std=1.0, bern=None):
    """MNFLeNet, mnf_lenet.py:22-33.  Ps = [conv1, conv2, dense1, dense2] param dicts;
    noise = [(eps_z, eps_out)] * 4 in program order (App. A-6)."""
    bern = bern or [None] * 4
    h, _ = conv_call(Ps[0], x, *noise[0], max_std=max_std, bern_q=bern[0])
    h = maxpool2x2_same(relu(h))
    h, _ = conv_call(Ps[1], h, *noise[1], max_std=max_std, bern_q=bern[1])
    h = maxpool2x2_same(relu(h))
    h = h.reshape(h.shape[0], -1)
    h, _ = dense_call(Ps[2], h, *noise[2], max_std=max_std, bern_q=bern[2])
    h = relu(h)
    h, _ = dense_call(Ps[3], h, *noise[3], max_std=max_std, bern_q=bern[3])
    return softmax(h)


def lenet_kl(Ps, kl_noise, r_term="all_states", bern=None):
    """MNFLeNet.kl_div, mnf_lenet.py:35-39.  kl_noise = [(eps_z, eps_a, eps_b)]*2 + [(eps_z, eps_a)]*2."""
    bern = bern or [(None, None)] * 4
    tot = 0.0
    for i in range(2):
        tot = tot + conv_kl(Ps[i], *kl_noise[i], r_term=r_term, bern_q=bern[i][0], bern_r=bern[i][1])[0]
    for i in range(2, 4):
        tot = tot + dense_kl(Ps[i], *kl_noise[i], r_term=r_term, bern_q=bern[i][0], bern_r=bern[i][1])[0]
    return tot


# ------------------------------------------------------------------- parameter creation
def glorot_normal(rng, shape, dtype=np.float32):
    """keras GlorotNormal: truncated normal (+-2 sigma), sigma = sqrt(2/(fi+fo))/0.87962566."""
    shape = tuple(shape)
    if len(shape) == 1:
        fi = fo = shape[0]
    elif len(shape) == 2:
        fi, fo = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fi, fo = shape[-2] * rf, shape[-1] * rf
    std = np.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
    v = rng.standard_normal(int(np.prod(shape)))
    bad = np.abs(v) > 2
    while bad.any():
        v[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(v) > 2
    return (v * std).reshape(shape).astype(dtype)


def glorot_uniform(rng, shape, dtype=np.float32):
    fi, fo = shape
    lim = np.sqrt(6.0 / (fi + fo))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def make_flow(rng, kind, D, n_steps, h_sizes=(50,), made_masks_mode="ones", dtype=np.float32):
    steps = []
    for i in range(n_steps):
        if kind == "iaf":
            sizes = [D, *h_sizes, 2 * D]
            W = [glorot_uniform(rng, (sizes[l], sizes[l + 1]), dtype) for l in range(len(sizes) - 1)]
            b = [np.zeros(sizes[l + 1], dtype) for l in range(len(sizes) - 1)]
            masks = None if made_masks_mode == "ones" else [m.astype(dtype) for m in made_masks(D, h_sizes, 2)]
            steps.append(dict(kind="iaf", parity=i % 2, W=W, b=b, masks=masks))
        elif kind == "rnvp":
            sizes = [D, *h_sizes]
            Wn = [glorot_uniform(rng, (sizes[l], sizes[l + 1]), dtype) for l in range(len(h_sizes))]
            bn = [np.zeros(h, dtype) for h in h_sizes]
            H = h_sizes[-1]
            steps.append(dict(kind="rnvp", Wn=Wn, bn=bn, Wt=glorot_uniform(rng, (H, D), dtype),
                              bt=np.zeros(D, dtype), Ws=glorot_uniform(rng, (H, D), dtype), bs=np.zeros(D, dtype)))
        elif kind == "planar":
            steps.append(dict(kind="planar", u=glorot_normal(rng, (D, 1), dtype),
                              w=glorot_uniform(rng, (D, 1), dtype), b=np.zeros(1, dtype)))
        else:
            raise ValueError(kind)
    return steps


def make_layer_params(rng, w_shape, n_flows_q=2, n_flows_r=2, flow="iaf", h_sizes=(50,), std_init=1.0,
                      prior_var_w=1.0, prior_var_b=1.0, made_masks_mode="ones", dtype=np.float32):
    """MNFDense.build mnf_dense.py:48-86 (w_shape=(n_in,n_out), z on inputs) and
    MNFConv2D.build mnf_conv.py:55-98 (w_shape=(kh,kw,C,F), z on filters)."""
    conv = len(w_shape) == 4
    n_out = w_shape[-1]
    D = n_out if conv else w_shape[0]
    Dp = int(np.prod(w_shape[:-1])) if conv else w_shape[0]
    G = lambda s: glorot_normal(rng, s, dtype)  # noqa: E731
    P = dict(
        mean_W=G(w_shape),
        log_std_W=(G(w_shape) * std_init - 9).astype(dtype),
        mean_b=np.zeros(n_out, dtype),
        log_var_b=(G((n_out,)) * std_init - 9).astype(dtype),
        q0_mean=(G((D,)) + (0 if n_flows_q > 0 else 1)).astype(dtype),
        q0_log_var=(G((D,)) * std_init - 9).astype(dtype),
        r0_mean=G((D,)),
        r0_log_var=G((D,)),
        r0_apvar=G((D,)),
        prior_var_r_p=(G((Dp,)) * std_init + np.log(prior_var_w)).astype(dtype),
        prior_var_r_p_bias=(G((1,)) * std_init + np.log(prior_var_b)).astype(dtype),
    )
    P["flow_r"] = make_flow(rng, flow, D, n_flows_r, h_sizes, made_masks_mode, dtype)
    P["flow_q"] = make_flow(rng, flow, D, n_flows_q, h_sizes, made_masks_mode, dtype)
    return P


def make_lenet_params(seed=0, flow="iaf", std_init=10.0, dtype=np.float32, s=(20, 50, 500, 10), **kw):
    rng = np.random.RandomState(seed)
    shapes = [(5, 5, 1, s[0]), (5, 5, s[0], s[1]), (4 * 4 * s[1], s[2]), (s[2], s[3])]
    return [make_layer_params(rng, shp, flow=flow, std_init=std_init, dtype=dtype, **kw) for shp in shapes]


def cast_params(P, dtype):
    """Deep-cast a parameter dict (or list of them) to dtype."""
    if isinstance(P, dict):
        return {k: cast_params(v, dtype) for k, v in P.items()}
    if isinstance(P, (list, tuple)):
        return [cast_params(v, dtype) for v in P]
    if isinstance(P, np.ndarray) and P.dtype.kind == "f":
        return P.astype(dtype)
    return P
