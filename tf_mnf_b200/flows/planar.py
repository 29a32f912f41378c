"""PlanarFlow host object (reference: tf_mnf/flows/planar.py:8-36); forward only.
The Dense(1) is built eagerly from `dim` (the reference reads its kernel before Keras has
built it, App. B-5).  Kernel: csrc/flows.cu (warp-per-row dot, tanh, rank-1 update)."""
from __future__ import annotations

from .. import _lib as L
from .. import init
from ..runtime import default_context
from .base import Flow


class PlanarFlow(Flow):
    kind = L.MNF_FLOW_PLANAR

    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = int(dim)
        self.u = self.w = self.b = None

    def build(self, dim: int | None = None, ctx=None) -> None:
        ctx = ctx or default_context()
        if dim is not None and dim != self.dim:
            raise ValueError(f"PlanarFlow was created for dim={self.dim} but got input of width {dim}")
        rng = init.param_rng(self.uid)
        self.u = ctx.to_device(init.glorot_normal(rng, (self.dim, 1)))  # planar.py:19
        self.w = ctx.to_device(init.glorot_uniform(rng, (self.dim, 1)))  # Dense(1).kernel, planar.py:20
        self.b = ctx.zeros((1,))
        self.built = True

    def variables(self) -> dict:
        return {"u": self.u, "w": self.w, "b": self.b}

    def fill_step(self, st: L.FlowStep, ptr_of=lambda a: a.ptr) -> None:
        st.kind, st.parity, st.hidden = self.kind, 0, 1
        st.w1, st.b1, st.wt = ptr_of(self.w), ptr_of(self.b), ptr_of(self.u)
        st.ws = st.bs = st.bt = None
        st.ld2 = 0
        st.m1 = st.m2 = None

    def forward(self, z):
        return self._single_step(z)
