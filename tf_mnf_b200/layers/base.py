"""Common host logic of MNFDense / MNFConv2D: constructor contract, variable creation,
sample_z and kl_div plumbing.  Arithmetic lives in libmnf_b200 (csrc/*.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from .. import _lib as L
from .. import init
from .. import noise as nz
from ..flows import IAF, RNVP, NormalizingFlow, PlanarFlow
from ..flows.base import ChainRun
from ..runtime import Context, DeviceArray, as_device, default_context, param_version

_VP = C.c_void_p


def _p(a: DeviceArray | None):
    return _VP(a.ptr) if a is not None else None


class MNFLayerBase:
    """Keyword contract of the reference layers (mnf_dense.py:22-35, mnf_conv.py:22-39) plus
    additive options: flow_type ("iaf" default | "rnvp" | "planar", App. B-7), made_masks
    ("ones" default = reference as-run | "autoregressive", App. B-2), r_term ("all_states"
    default = as written | "last", App. B-3), seed."""

    _conv = False

    def __init__(self, n_flows_q: int = 2, n_flows_r: int = 2, learn_p: bool = False, use_z: bool = True,
                 prior_var_w: float = 1, prior_var_b: float = 1, flow_h_sizes=(50,), max_std: float = 1,
                 std_init: float = 1, flow_type: str = "iaf", made_masks: str = "ones",
                 r_term: str = "all_states", seed: int | None = None, name: str | None = None, **kwargs: Any) -> None:
        if kwargs:  # keras.layers.Layer.__init__ rejects unknown keywords the same way
            raise TypeError(f"{type(self).__name__}: unexpected keyword arguments {sorted(kwargs)}")
        if flow_type not in ("iaf", "rnvp", "planar"):
            raise ValueError(f"flow_type must be 'iaf', 'rnvp' or 'planar', got {flow_type!r}")
        if r_term not in ("all_states", "last"):
            raise ValueError(f"r_term must be 'all_states' or 'last', got {r_term!r}")
        if made_masks not in ("ones", "autoregressive"):
            raise ValueError(f"made_masks must be 'ones' or 'autoregressive', got {made_masks!r}")
        self.learn_p, self.use_z = bool(learn_p), bool(use_z)
        self.prior_var_w, self.prior_var_b = prior_var_w, prior_var_b
        self.max_std, self.std_init = float(max_std), float(std_init)
        self.n_flows_q, self.n_flows_r = int(n_flows_q), int(n_flows_r)
        self.flow_h_sizes = tuple(int(h) for h in flow_h_sizes)
        self.flow_type, self.made_masks, self.r_term = flow_type, made_masks, r_term
        self.uid = init.next_uid()
        self.seed = init.global_seed() if seed is None else int(seed)
        self.name = name or f"{type(self).__name__.lower()}_{self.uid}"
        self.built = False
        self.ctx: Context | None = None
        self.row_offset = 0       # global index of local row 0 (set by the sharded driver)
        self.call_index = 0       # bumps every call() / sample_z(): fresh noise per call
        self.kl_index = 0         # bumps every kl_div() (its draws use their own stream ids, so the two
        #                           counters never collide; separate so that kl_div() between two
        #                           forward calls does not invalidate a z drawn ahead of time)
        self.training = False     # record what backward needs
        self.grads: dict[str, Any] = {}
        self._tape: dict[str, Any] = {}
        self._kl_tape: dict[str, Any] = {}
        self._pre: dict[str, Any] | None = None  # z sampled ahead of call() on a side context

    # ------------------------------------------------------------------ variables
    def _flow_dim(self) -> int:  # pragma: no cover - abstract
        raise NotImplementedError

    def _make_flow(self, n: int) -> NormalizingFlow:
        D = self._flow_dim()
        if self.flow_type == "iaf":
            fl = [IAF(parity=i % 2, h_sizes=self.flow_h_sizes, mask_mode=self.made_masks) for i in range(n)]
        elif self.flow_type == "rnvp":
            fl = [RNVP(D, h_sizes=self.flow_h_sizes) for _ in range(n)]
        else:
            fl = [PlanarFlow(D) for _ in range(n)]
        for f in fl:
            f.build(D, self.ctx)
        return NormalizingFlow(fl)

    def _build_common(self, w_shape, n_prior: int) -> None:
        """Variable creation in the reference's order (mnf_dense.py:50-86 / mnf_conv.py:57-98)."""
        ctx = self.ctx = self.ctx or default_context()
        rng = init.param_rng(self.uid)
        G = lambda s: init.glorot_normal(rng, s)  # noqa: E731
        n_out = w_shape[-1]
        D = self._flow_dim()
        si, mi = self.std_init, -9.0
        dev = lambda a: ctx.to_device(np.asarray(a, np.float32))  # noqa: E731
        self.mean_W = dev(G(w_shape))
        self.log_std_W = dev(G(w_shape) * si + mi)
        self.mean_b = ctx.zeros((n_out,))
        self.log_var_b = dev(G((n_out,)) * si + mi)
        if self.use_z:
            self.q0_mean = dev(G((D,)) + (0 if self.n_flows_q > 0 else 1))
            self.q0_log_var = dev(G((D,)) * si + mi)
            self.r0_mean = dev(G((D,)))
            self.r0_log_var = dev(G((D,)))
            self.r0_apvar = dev(G((D,)))
        self.prior_var_r_p = dev(G((n_prior,)) * si + np.log(self.prior_var_w))
        self.prior_var_r_p_bias = dev(G((1,)) * si + np.log(self.prior_var_b))
        self.flow_r = self._make_flow(self.n_flows_r)
        self.flow_q = self._make_flow(self.n_flows_q)
        self.built = True

    def variables(self) -> dict[str, DeviceArray]:
        names = ["mean_W", "log_std_W", "mean_b", "log_var_b"]
        if self.use_z:
            names += ["q0_mean", "q0_log_var", "r0_mean", "r0_log_var", "r0_apvar"]
        names += ["prior_var_r_p", "prior_var_r_p_bias"]
        out = {n: getattr(self, n) for n in names}
        for tag, nf in (("flow_q", self.flow_q), ("flow_r", self.flow_r)):
            for k, fl in enumerate(nf.flows):
                for n, v in fl.variables().items():
                    out[f"{tag}/{k}/{n}"] = v
        return out

    @property
    def trainable_variables(self) -> dict[str, DeviceArray]:
        v = self.variables()
        if not self.learn_p:
            v.pop("prior_var_r_p"), v.pop("prior_var_r_p_bias")
        return v

    # ------------------------------------------------------------------- sample_z
    def _sample_z(self, batch_size: int, z_draw: int, bern_draw: int, inject: dict | None, record: bool,
                  row_offset: int, ctx: Context | None = None, alloc: Context | None = None, index: int | None = None):
        """mnf_dense.py:88-102 / mnf_conv.py:100-114 -> (z, log_dets, ChainRun | None)."""
        ctx, D = ctx or self.ctx, self._flow_dim()
        B = int(batch_size)
        if not self.use_z:
            return None, ctx.zeros((B,)), None
        inject = inject or {}
        index = self.call_index if index is None else index
        pre, self._pre = (self._pre, None) if z_draw == nz.CALL_Z else (None, self._pre)
        if pre is not None:
            ctx.wait(pre["side"])  # also orders the buffers' (stream-ordered) release after the side work
            # ... drawn from the CURRENT parameters: a z sampled before an assign() / optimiser step is stale
            if (not inject and pre["key"] == (B, index, row_offset, bool(record), param_version()) and ctx is self.ctx):
                return pre["z"], pre["ld"], pre["run"]
        eps_inj = as_device(inject["eps_z"], ctx) if inject.get("eps_z") is not None else None
        if eps_inj is not None and eps_inj.shape != (B, D):
            raise ValueError(f"injected eps_z must have shape {(B, D)}, got {eps_inj.shape}")
        eps = nz.make(self.seed, index, self.uid, z_draw, row_offset, eps_inj)
        run = ChainRun(self.flow_q.flows)
        states, ld = run.forward(q0_mean=self.q0_mean, q0_log_var=self.q0_log_var, B=B, eps=eps,
                                 bern=inject.get("bern_q"), seed=self.seed, call_index=index,
                                 layer_id=self.uid, bern_draw=bern_draw, row_offset=row_offset, record=record,
                                 ctx=ctx, alloc=alloc)
        run.eps = eps
        run._eps_keep = eps_inj
        return states[-1], ld, run

    def presample(self, batch_size: int, side: Context) -> None:
        """Enqueue the z draw of the NEXT call() on `side` (a Context.fork() of this layer's
        context): z does not depend on x, so it can be computed beside the layers in front of
        this one.  Same Philox counters as the in-line draw -> bit-identical z.  The buffers
        belong to the layer's own context; call() joins the two streams."""
        self._require_built()
        if not self.use_z or side is self.ctx:
            return
        if self._pre is not None:
            self.ctx.wait(self._pre["side"])
        self._pre = None
        B = int(batch_size)
        z, ld, run = self._sample_z(B, nz.CALL_Z, nz.BERN_Q_CALL, None, self.training, self.row_offset, side, self.ctx)
        self._pre = dict(key=(B, self.call_index, self.row_offset, bool(self.training), param_version()), z=z, ld=ld,
                         run=run, side=side)

    def sample_z(self, batch_size: int, noise: dict | None = None):  # noqa: A002 - mirrors reference name
        self._require_built()
        z, ld, _ = self._sample_z(batch_size, 0, 16, noise, False, self.row_offset)
        if z is None:
            z = self.ctx.to_device(np.ones((int(batch_size), self._flow_dim()), np.float32))
        self.call_index += 1
        return z, ld

    def _require_built(self) -> None:
        if not self.built:
            raise RuntimeError(f"{self.name}: call build(input_shape) (or the layer on an input) first")

    # ---------------------------------------------------------------------- kl_div
    def _kl_shape(self) -> tuple[int, int]:  # pragma: no cover - abstract
        raise NotImplementedError

    def kl_div(self, noise: dict | None = None, ctx: Context | None = None) -> DeviceArray:  # noqa: A002
        """MNFDense.kl_div mnf_dense.py:104-157 / MNFConv2D.kl_div mnf_conv.py:116-180.
        Returns a 1-element device array (float(kl) copies it to the host).  `ctx`: run on
        another context/stream (kl_div does not depend on x); join with `main.wait(ctx)`."""


        self._require_built()
        ctx, D = ctx or self.ctx, self._flow_dim()
        inj = noise or {}
        rows, cols = self._kl_shape()
        z, ldq, qrun = self._sample_z(1, nz.KL_Z, nz.BERN_Q_KL, inj, self.training, 0, ctx, index=self.kl_index)
        a = L.KlArgs()
        a.conv, a.use_z = int(self._conv), int(self.use_z)
        a.r_all_states = int(self.r_term == "all_states")
        a.rows, a.cols = rows, cols
        a.mean_W, a.log_std_W, a.mean_b, a.log_var_b = _p(self.mean_W), _p(self.log_std_W), _p(self.mean_b), _p(self.log_var_b)
        a.prior_var_r_p, a.prior_var_r_p_bias = _p(self.prior_var_r_p), _p(self.prior_var_r_p_bias)
        keep = []
        rrun = None
        if self.use_z:
            rrun = ChainRun(self.flow_r.flows)
            rstates, ldr = rrun.forward(z=z.reshape(1, D), bern=inj.get("bern_r"), seed=self.seed,
                                        call_index=self.kl_index, layer_id=self.uid, bern_draw=nz.BERN_R_KL,
                                        record=self.training, ctx=ctx)
            a.n_r_states = len(rstates)
            a.q0_log_var, a.r0_mean, a.r0_log_var, a.r0_apvar = _p(self.q0_log_var), _p(self.r0_mean), _p(self.r0_log_var), _p(self.r0_apvar)
            a.z, a.ld_q, a.r_states, a.ld_r = _p(z), _p(ldq), _p(rrun.states), _p(ldr)
            L_a = rows if self._conv else cols
            ea = as_device(inj["eps_a"], ctx) if inj.get("eps_a") is not None else None
            if ea is not None and ea.size != L_a:
                raise ValueError(f"injected eps_a must have {L_a} elements, got shape {ea.shape}")
            eb = as_device(np.asarray(inj["eps_b"], np.float32).reshape(1), ctx) if inj.get("eps_b") is not None else None
            keep += [ea, eb, ldr]
            a.eps_a = nz.make(self.seed, self.kl_index, self.uid, nz.KL_A, 0, ea)
            a.eps_b = nz.make(self.seed, self.kl_index, self.uid, nz.KL_B, 0, eb)
        out = ctx.empty((1,))
        L.check(ctx.lib.mnf_kl_forward(ctx.handle, C.byref(a), _p(out)))
        if self.training:
            self._kl_tape = dict(args=a, qrun=qrun, rrun=rrun, keep=keep, z=z, ldq=ldq)
        self.kl_index += 1
        return out

    # ------------------------------------------------------------------- backward
    def zero_grads(self) -> dict[str, DeviceArray]:
        """(Re)create zeroed gradient buffers, one per variable (same keys as variables())."""
        ctx = self.ctx
        for name, v in self.variables().items():
            g = self.grads.get(name)
            if g is None or g.shape != v.shape:
                self.grads[name] = ctx.zeros(v.shape)
            else:
                L.check(ctx.lib.mnf_memset(ctx.handle, _p(g), 0, g.nbytes))
        return self.grads

    def _grad_steps(self, tag: str, nf: NormalizingFlow, grads: dict | None = None):
        """mnf_flow_step array whose pointer slots name the GRADIENT buffers of flow `tag`."""
        grads = self.grads if grads is None else grads
        steps = (L.FlowStep * max(len(nf.flows), 1))()
        for k, fl in enumerate(nf.flows):
            by_id = {id(v): grads[f"{tag}/{k}/{n}"] for n, v in fl.variables().items()}
            fl.fill_step(steps[k], ptr_of=lambda a, m=by_id: m[id(a)].ptr)
        return steps

    def _flow_backward(self, tag: str, nf: NormalizingFlow, run: ChainRun, d_states, d_zT, d_ld,
                       ctx: Context | None = None, grads: dict | None = None) -> DeviceArray:
        ctx = ctx or self.ctx
        B, D, T = run.B, run.D, len(nf.flows)
        d_z0 = ctx.empty((B, D))
        gsteps = self._grad_steps(tag, nf, grads)
        L.check(ctx.lib.mnf_flow_backward(ctx.handle, B, D, T, run.steps, _p(run.states), _p(run.saved),
                                          _p(d_states), _p(d_zT), _p(d_ld), _p(d_z0), gsteps))
        return d_z0

    def _sample_z_backward(self, run: ChainRun, dz: DeviceArray, d_ld: DeviceArray | None,
                           ctx: Context | None = None, grads: dict | None = None) -> None:
        """Through flow_q back to q0_mean / q0_log_var; accumulates."""
        ctx = ctx or self.ctx
        grads = self.grads if grads is None else grads
        d_z0 = self._flow_backward("flow_q", self.flow_q, run, None, dz, d_ld, ctx, grads)
        L.check(ctx.lib.mnf_sample_z_backward(ctx.handle, run.B, run.D, _p(d_z0), _p(self.q0_log_var),
                                              C.byref(run.eps), _p(grads["q0_mean"]), _p(grads["q0_log_var"])))

    def _add_grad(self, name: str, src: DeviceArray, ctx: Context | None = None, grads: dict | None = None) -> None:
        dst = (self.grads if grads is None else grads)[name]
        ctx = ctx or self.ctx
        L.check(ctx.lib.mnf_axpy(ctx.handle, dst.size, 1.0, _p(src), _p(dst)))

    def kl_backward(self, scale: float = 1.0, ctx: Context | None = None, grads: dict | None = None,
                    fresh_grads: bool = False) -> None:
        """Gradient of scale * kl_div() (the most recent kl_div() call made with
        `training=True`) accumulated into self.grads.
        `ctx` / `grads`: run on the (forked) context kl_div() ran on and accumulate into another
        set of buffers, so the KL branch of a training step runs beside the likelihood branch."""
        t = self._kl_tape
        if not t:
            raise RuntimeError(f"{self.name}: call kl_div() with layer.training = True before kl_backward()")
        if grads is None and not self.grads:
            self.zero_grads()
        ctx, D = ctx or self.ctx, self._flow_dim()
        a: L.KlArgs = t["args"]
        g = L.KlGrads()
        names = ["mean_W", "log_std_W", "mean_b", "log_var_b"]
        if self.learn_p:
            names += ["prior_var_r_p", "prior_var_r_p_bias"]
        if self.use_z:
            names += ["q0_log_var", "r0_mean", "r0_log_var", "r0_apvar"]
        # fresh_grads: the target buffers were just zeroed and nothing else has written them yet, so
        # the kernel's (overwriting) outputs go there directly instead of through temporaries + axpy
        target = self.grads if grads is None else grads
        scratch = {n: target[n] for n in names} if fresh_grads else {n: ctx.empty(getattr(self, n).shape) for n in names}
        for n, buf in scratch.items():
            setattr(g, n, buf.ptr)
        d_z = d_rst = None
        if self.use_z:
            d_z = ctx.empty((1, D))
            d_rst = ctx.empty((a.n_r_states, 1, D))
            g.d_z, g.d_r_states = d_z.ptr, d_rst.ptr
        L.check(ctx.lib.mnf_kl_backward(ctx.handle, C.byref(a), float(scale), C.byref(g)))
        if not fresh_grads:
            for n, buf in scratch.items():
                self._add_grad(n, buf, ctx, grads)
        if self.use_z:
            # dKL/d(ld_q) = dKL/d(ld_r) = -scale; the one-element array is cached per (context, scale):
            # uploading it synchronises the stream
            key = (id(ctx), float(scale))
            d_ld = self._dld_cache.get(key) if hasattr(self, "_dld_cache") else None
            if d_ld is None:
                if not hasattr(self, "_dld_cache"):
                    self._dld_cache = {}
                d_ld = self._dld_cache[key] = ctx.to_device(np.array([-float(scale)], np.float32))
            dz_r = self._flow_backward("flow_r", self.flow_r, t["rrun"], d_rst, None, d_ld, ctx, grads)
            L.check(ctx.lib.mnf_axpy(ctx.handle, D, 1.0, _p(dz_r), _p(d_z)))
            self._sample_z_backward(t["qrun"], d_z, d_ld, ctx, grads)
