"""A whole step (forward + kl_div, on however many forked streams) captured ONCE into a CUDA
graph and replayed with one launch.

What makes a stochastic step replayable: every noise plan made during capture points at the
graph's device-side replay counter (mnf_noise.call_delta), which the graph itself bumps at its
end, so replay i draws with call index + i -- the same Philox counters an eager loop would use,
bit for bit (tests/test_gpu_streams.py).  Requirements checked here: the model's weights are
frozen (no host-side packing inside the step), one eager warm-up call has sized the workspace
and the buffer pools, and every forked context is joined back before the capture ends.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Sequence

from . import _lib as L
from . import noise as nz
from .runtime import Context, DeviceArray


class CapturedStep:
    """fn() -> DeviceArray | tuple of DeviceArrays, enqueued on `ctx` (and on `forks`, contexts
    fn uses besides ctx).  `layers`: the MNF layers fn calls, whose Python-side counters advance
    by (n_calls, n_kl) per replay so that eager code afterwards keeps drawing fresh noise."""

    def __init__(self, ctx: Context, fn: Callable[[], Any], forks: Sequence[Context] = (), layers: Sequence[Any] = (),
                 calls_per_step: int = 1, kl_per_step: int = 1, models: Sequence[Any] = ()) -> None:
        self.ctx, self.forks, self.layers = ctx, list(forks), list(layers)
        self.calls_per_step, self.kl_per_step = calls_per_step, kl_per_step
        self.models = list(models)
        for m in models:
            if not all(getattr(lyr.ctx, "frozen", False) for lyr in m.mnf_layers if lyr.ctx is not None):
                raise RuntimeError("CapturedStep: call model.freeze_weights() first (weight packing uploads tables "
                                   "from host memory, which a graph cannot replay)")
        every = [ctx, *self.forks]

        def sync_all() -> None:
            for c in every:
                c.sync()

        fn()  # eager warm-up: variables, workspaces, function attributes, packed weights
        for m in self.models:  # side contexts the models forked for themselves (created lazily)
            for c in getattr(m, "_sides", []):
                if all(c is not e for e in every):
                    self.forks.append(c)
                    every.append(c)
        sync_all()
        # A captured step re-uses no buffer (each stays reserved for the graph's lifetime), so it
        # needs more blocks than an eager step ever holds at once: one eager pass in the same
        # keep-everything mode stocks the pools, and the capture below then allocates nothing.
        stock: list = []
        for c in every:
            c._capture_keep = stock
        for _ in range(2):  # twice: a z drawn one call ahead outlives its pass, the capture needs a spare
            fn()
            sync_all()
        for c in every:
            c._capture_keep = None
        del stock
        self.counter = ctx.zeros((1,), "u4")
        ctx.sync()
        keep: list = []
        for c in every:
            c._capture_keep = keep
            c._capturing = True
        nz._capture_delta = self.counter.ptr
        L.check(ctx.lib.mnf_graph_begin(ctx.handle))
        try:
            for c in self.forks:
                c.wait(ctx)  # fork: the other streams join the capture
                # ... and must fail as clearly as the capturing context if they would grow a workspace
                # or miss the packed-weight cache now (cudaFree / cudaMalloc would invalidate the capture)
                L.check(ctx.lib.mnf_ctx_set_capturing(c.handle, 1))
            # z drawn one call ahead (Sequential.presample_first): the captured step reads the z the
            # previous replay drew (buffer X, filled eagerly for replay 0 by the stocking pass above)
            # and draws the next one into another buffer Y; Y is copied onto X at the end of the graph.
            carried = [(lyr, lyr._pre) for m in self.models for lyr in m.mnf_layers if lyr._pre is not None]
            self.outputs = fn()
            for c in self.forks:
                ctx.wait(c)  # join whatever fn left pending on them
            for lyr, x_pre in carried:
                y_pre = lyr._pre
                if y_pre is None or y_pre is x_pre:
                    continue
                for src, dst in ((y_pre["run"].states, x_pre["run"].states), (y_pre["ld"], x_pre["ld"])):
                    L.check(ctx.lib.mnf_memcpy_d2d(ctx.handle, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), src.nbytes))
                keep.append((x_pre, y_pre))
                lyr._pre = None  # eager calls after the graph draw their z in line
            L.check(ctx.lib.mnf_u32_add(ctx.handle, C.c_void_p(self.counter.ptr), 1))
        finally:
            nz._capture_delta = None
            for c in every:
                c._capture_keep = None
                c._capturing = False
            for c in self.forks:
                ctx.lib.mnf_ctx_set_capturing(c.handle, 0)
            g = C.c_void_p()
            rc = ctx.lib.mnf_graph_end(ctx.handle, C.byref(g))
        L.check(rc)
        self._exec = g
        self._keep = keep  # every buffer the step touches stays reserved while the graph lives
        self.replays = 0
        # the capture advanced the layers' counters once without running anything: replay 0 uses
        # exactly those indices, so nothing to undo

    def launch(self) -> Any:
        """Enqueue one replay on ctx's stream; returns the (re-used) output arrays."""
        L.check(self.ctx.lib.mnf_graph_launch(self.ctx.handle, self._exec))
        if self.replays:  # the capture itself already counted the first step
            for lyr in self.layers:
                lyr.call_index += self.calls_per_step
                lyr.kl_index += self.kl_per_step
        self.replays += 1
        return self.outputs

    def close(self) -> None:
        if getattr(self, "_exec", None):
            self.ctx.sync()
            self.ctx.lib.mnf_graph_destroy(self._exec)
            self._exec = None
            self._keep = []

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
