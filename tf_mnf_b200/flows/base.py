"""Shared machinery of the flow host classes: marshal a list of flows into mnf_flow_step
structs and run the whole chain through ONE C-ABI call (mnf_flow_forward)."""
from __future__ import annotations

import ctypes as C
from typing import Any, Sequence

from .. import _lib as L
from .. import init, noise
from ..runtime import Context, DeviceArray, as_device, default_context


class Flow:
    """Base of IAF/MAF, RNVP, PlanarFlow.  Mirrors the tf.Module protocol loosely:
    `forward(z) -> (x, log_det[B])`, parameters as public attributes."""

    kind = -1

    def __init__(self) -> None:
        self.uid = init.next_uid()
        self.built = False
        self.dim: int | None = None
        self._calls = 0

    def build(self, dim: int, ctx: Context | None = None) -> None:  # pragma: no cover - abstract
        raise NotImplementedError

    def variables(self) -> dict:  # pragma: no cover - abstract
        raise NotImplementedError

    def fill_step(self, st: L.FlowStep, ptr_of=lambda a: a.ptr) -> None:  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def hidden(self) -> int:
        """Floats per row the step saves for its backward pass: the sum of its hidden widths."""
        if self.kind == L.MNF_FLOW_PLANAR:
            return 0
        return int(sum(self.h_sizes))

    def _single_step(self, z: Any, bern: Sequence[Any] | None = None):
        chain = ChainRun([self])
        states, ld = chain.forward(z=z, bern=bern, seed=init.global_seed(), call_index=self._calls,
                                   layer_id=self.uid, bern_draw=noise.FLOW_BERN)
        self._calls += 1
        return states[1], ld


class ChainRun:
    """One invocation of a flow chain on the device; keeps what backward needs."""

    def __init__(self, flows: Sequence[Flow]) -> None:
        self.flows = list(flows)
        self.states: DeviceArray | None = None  # [T+1,B,D]
        self.log_dets: DeviceArray | None = None
        self.saved: DeviceArray | None = None
        self.steps = None
        self.B = self.D = 0
        self._keep: list = []

    def _marshal(self, ctx: Context, D: int, seed: int, call_index: int, layer_id: int, bern_draw: int,
                 row_offset: int, bern: Sequence[Any] | None):
        T = len(self.flows)
        steps = (L.FlowStep * max(T, 1))()
        bi = 0
        self._keep = []
        for k, fl in enumerate(self.flows):
            if not fl.built:
                fl.build(D, ctx)
            if fl.dim is not None and fl.dim != D:
                raise ValueError(f"flow step {k} has dim {fl.dim} but the input has width {D}")
            fl.fill_step(steps[k])
            inj = None
            if fl.kind == L.MNF_FLOW_RNVP and bern is not None:
                inj = as_device(bern[bi], ctx)
                self._keep.append(inj)
                bi += 1
            steps[k].bern = noise.make(seed, call_index, layer_id, bern_draw + k, row_offset, inj)
        return steps

    def forward(self, z: Any = None, q0_mean: DeviceArray | None = None, q0_log_var: DeviceArray | None = None,
                B: int | None = None, eps: L.Noise | None = None, bern: Sequence[Any] | None = None,
                seed: int = 0, call_index: int = 0, layer_id: int = 0, bern_draw: int = noise.FLOW_BERN,
                row_offset: int = 0, record: bool = False, ctx: Context | None = None,
                alloc: Context | None = None):
        """`alloc`: context whose stream owns the output buffers when the chain itself runs on
        another (`ctx`); the caller orders `ctx` after `alloc` before and `alloc` after `ctx`
        behind this call (MNFLayerBase.presample)."""
        if z is not None:
            z = as_device(z, ctx)
            ctx = z.ctx
            if z.ndim != 2:
                raise ValueError(f"flow input must be [batch, dim], got shape {z.shape}")
            B, D = z.shape
        else:
            ctx = ctx or q0_mean.ctx
            D = q0_mean.shape[0]
        T = len(self.flows)
        steps = self._marshal(ctx, D, seed, call_index, layer_id, bern_draw, row_offset, bern)
        actx = alloc or ctx
        states = actx.empty((T + 1, B, D))
        ld = actx.empty((B,))
        saved = None
        if record and T:
            saved = actx.empty((sum(B * max(f.hidden, 1) for f in self.flows),))
        if alloc is not None and alloc is not ctx:
            ctx.wait(alloc)  # the buffers exist on ctx's stream only behind alloc's allocation point
        L.check(ctx.lib.mnf_flow_forward(
            ctx.handle, B, D, T, steps,
            C.c_void_p(z.ptr) if z is not None else None,
            C.c_void_p(q0_mean.ptr) if q0_mean is not None else None,
            C.c_void_p(q0_log_var.ptr) if q0_log_var is not None else None,
            C.byref(eps) if eps is not None else None,
            C.c_void_p(states.ptr), C.c_void_p(ld.ptr), C.c_void_p(saved.ptr) if saved is not None else None))
        self.states, self.log_dets, self.saved, self.steps, self.B, self.D = states, ld, saved, steps, B, D
        self._z_in = z
        return [states[k] for k in range(T + 1)], ld
