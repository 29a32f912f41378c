"""tf_mnf.flows look-alike on libmnf_b200 (reference: tf_mnf/flows/__init__.py:13-57)."""
from __future__ import annotations

from typing import Any, Sequence

from .. import init, noise
from .base import ChainRun, Flow
from .made import MADE, MaskedDense
from .maf import IAF, MAF
from .planar import PlanarFlow
from .rnvp import RNVP

__all__ = ["IAF", "MAF", "MADE", "MaskedDense", "RNVP", "PlanarFlow", "NormalizingFlow", "NormalizingFlowModel", "DiagNormal"]


class NormalizingFlow:
    """A chain of flows is a flow.  `forward(z)` returns the LIST of all T+1 intermediate
    states and the summed per-row log-determinants (flows/__init__.py:22-29); the whole
    chain is one C-ABI call."""

    def __init__(self, flows: Sequence[Flow]) -> None:
        self.flows = list(flows)
        self.uid = init.next_uid()
        self._calls = 0
        self.last_run: ChainRun | None = None

    def forward(self, z: Any, bern: Sequence[Any] | None = None, record: bool = False):
        if any(isinstance(fl, MAF) and not isinstance(fl, IAF) for fl in self.flows):
            # a MAF's forward is the sequential decode (maf.py:33-45), not the MADE pass the fused chain
            # runs: step by step through the flows' own forward(), as flows/__init__.py:22-29 does
            import ctypes as C

            from .. import _lib as L
            from ..runtime import as_device, default_context

            if bern is not None or record:
                raise NotImplementedError("injected masks / recording are for chains of IAF, RNVP and PlanarFlow steps")
            z = as_device(z, default_context())
            xs, total = [z], z.ctx.zeros((z.shape[0],))
            for fl in self.flows:
                z, ld = fl.forward(z)
                xs.append(z)
                L.check(z.ctx.lib.mnf_axpy(z.ctx.handle, ld.size, 1.0, C.c_void_p(ld.ptr), C.c_void_p(total.ptr)))
            return xs, total
        run = ChainRun(self.flows)
        states, ld = run.forward(z=z, bern=bern, seed=init.global_seed(), call_index=self._calls,
                                 layer_id=self.uid, bern_draw=noise.FLOW_BERN, record=record)
        self._calls += 1
        self.last_run = run
        return states, ld

    def backward(self, d_states: Any = None, d_zT: Any = None, d_log_dets: Any = None):
        """Gradient of the last `forward(z, record=True)`.  d_states [T+1,B,D] (grads for every
        returned state), d_zT [B,D] (extra grad for the last state), d_log_dets [B].
        Returns (d_z0, [ {name: grad} per flow ]) -- what tape.gradient gives in the reference."""
        import ctypes as C

        from .. import _lib as L
        from ..runtime import as_device

        run = self.last_run
        if run is None or (run.saved is None and any(f.kind != L.MNF_FLOW_PLANAR for f in self.flows)):
            raise RuntimeError("call forward(z, record=True) before backward()")
        ctx = run.states.ctx
        p = lambda a: C.c_void_p(a.ptr) if a is not None else None  # noqa: E731
        ds = as_device(d_states, ctx) if d_states is not None else None
        dz = as_device(d_zT, ctx) if d_zT is not None else None
        dl = as_device(d_log_dets, ctx) if d_log_dets is not None else None
        grads = [{n: ctx.zeros(v.shape) for n, v in fl.variables().items()} for fl in self.flows]
        gsteps = (L.FlowStep * max(len(self.flows), 1))()
        for k, fl in enumerate(self.flows):
            by_id = {id(v): grads[k][n] for n, v in fl.variables().items()}
            fl.fill_step(gsteps[k], ptr_of=lambda a, m=by_id: m[id(a)].ptr)
        d_z0 = ctx.empty((run.B, run.D))
        L.check(ctx.lib.mnf_flow_backward(ctx.handle, run.B, run.D, len(self.flows), run.steps, p(run.states),
                                          p(run.saved), p(ds), p(dz), p(dl), p(d_z0), gsteps))
        return d_z0, grads

    def inverse(self, x: Any):
        """flows/__init__.py:31-38: x -> z through the flows in reverse order; returns the LIST of states
        (starting with x) and the summed log-determinants.  A flow must have an `inverse` (MAF: one
        MADE pass; IAF: the sequential decode); RNVP and PlanarFlow have none in the reference either."""
        import ctypes as C

        from .. import _lib as L
        from ..runtime import as_device, default_context

        x = as_device(x, default_context())
        zs, total = [x], None
        for fl in reversed(self.flows):
            if not hasattr(fl, "inverse"):
                raise NotImplementedError(f"{type(fl).__name__} has no inverse() (neither in the reference)")
            x, ld = fl.inverse(x)
            zs.append(x)
            if total is None:
                total = ld
            else:
                L.check(x.ctx.lib.mnf_axpy(x.ctx.handle, ld.size, 1.0, C.c_void_p(ld.ptr), C.c_void_p(total.ptr)))
        if total is None:
            total = x.ctx.zeros((x.shape[0],))
        return zs, total


class DiagNormal:
    """A base distribution for NormalizingFlowModel (the reference takes any object with log_prob /
    sample, e.g. tfp.distributions.MultivariateNormalDiag): independent N(loc, scale^2) per dimension.
    sample() draws on the device (mnf_philox_normal); log_prob() returns per-row log densities [B]."""

    def __init__(self, dim: int, loc: float = 0.0, scale: float = 1.0, seed: int | None = None) -> None:
        self.dim, self.loc, self.scale = int(dim), float(loc), float(scale)
        self.seed = init.global_seed() if seed is None else int(seed)
        self.uid = init.next_uid()
        self._calls = 0

    def sample(self, num_samples):
        import ctypes as C

        import numpy as np

        from .. import _lib as L
        from ..runtime import default_context

        n = int(np.prod(num_samples)) if not isinstance(num_samples, int) else int(num_samples)
        ctx = default_context()
        out = ctx.empty((n, self.dim))
        seed = (self.seed & 0xFFFFFFFF) | ((self._calls & 0xFFFFFFFF) << 32)
        L.check(ctx.lib.mnf_philox_normal(ctx.handle, seed, noise.stream_id(self.uid, noise.FLOW_BERN - 1), 0, n, 1, self.dim,
                                          C.c_void_p(out.ptr)))
        self._calls += 1
        if self.loc != 0.0 or self.scale != 1.0:
            host = out.numpy() * self.scale + self.loc
            out = ctx.to_device(host)
        return out

    def log_prob(self, z):
        import numpy as np

        zz = (z.numpy() if hasattr(z, "numpy") else np.asarray(z)).astype(np.float64)
        u = (zz - self.loc) / self.scale
        return (-0.5 * u * u - np.log(self.scale) - 0.5 * np.log(2 * np.pi)).sum(axis=-1)


class NormalizingFlowModel(NormalizingFlow):
    """(base distribution, flow) pair (flows/__init__.py:41-57).  `base`: any object with
    log_prob(z) and sample(num_samples), e.g. DiagNormal above."""

    def __init__(self, base: Any, flow: Sequence[Flow]) -> None:
        super().__init__(flow)
        self.base = base

    def base_log_prob(self, x: Any):
        """flows/__init__.py:49-51: sum over the batch of base.log_prob(inverse(x)[-1])."""
        import numpy as np

        zs, _ = self.inverse(x)
        return float(np.sum(self.base.log_prob(zs[-1])))

    def sample(self, *num_samples: int):
        """flows/__init__.py:53-57: z ~ base, then all the states of forward(z)."""
        z = self.base.sample(num_samples)
        xs, _ = self.forward(z)
        return xs
