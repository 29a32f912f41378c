"""tf_mnf.flows look-alike on libmnf_b200 (reference: tf_mnf/flows/__init__.py:13-57)."""
from __future__ import annotations

from typing import Any, Sequence

from .. import init, noise
from .base import ChainRun, Flow
from .made import MADE, MaskedDense
from .maf import IAF, MAF
from .planar import PlanarFlow
from .rnvp import RNVP

__all__ = ["IAF", "MAF", "MADE", "MaskedDense", "RNVP", "PlanarFlow", "NormalizingFlow", "NormalizingFlowModel"]


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
        raise NotImplementedError("NormalizingFlow.inverse is never called by the MNF layers (SURVEY.md 8f-4)")


class NormalizingFlowModel(NormalizingFlow):
    """(base distribution, flow) pair (flows/__init__.py:41-57); unused by the layers."""

    def __init__(self, base: Any, flow: Sequence[Flow]) -> None:
        super().__init__(flow)
        self.base = base

    def base_log_prob(self, x: Any):
        raise NotImplementedError("needs NormalizingFlow.inverse (SURVEY.md 8f-4)")

    def sample(self, *num_samples: int):
        raise NotImplementedError("NormalizingFlowModel.sample is outside the MNF hot path (SURVEY.md 8f-4)")
