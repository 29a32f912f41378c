"""RNVP host object (reference: tf_mnf/flows/rnvp.py:8-47): gated coupling with a fresh
Bernoulli(0.5) mask per call and element; forward only.  Kernel: csrc/flows.cu."""
from __future__ import annotations

from .. import _lib as L
from .. import init
from ..runtime import default_context
from .base import Flow


class RNVP(Flow):
    kind = L.MNF_FLOW_RNVP

    def __init__(self, dim: int, h_sizes=(30,), activation: str = "tanh") -> None:
        super().__init__()
        h_sizes = tuple(int(h) for h in h_sizes)
        if len(h_sizes) != 1:
            raise NotImplementedError(f"h_sizes={h_sizes}: the fused sm_100a flow kernel supports exactly one hidden layer")
        if activation != "tanh":
            raise NotImplementedError(f"activation={activation!r}: only the reference default 'tanh' is implemented")
        self.dim, self.h_sizes, self.activation = int(dim), h_sizes, activation
        self.net_kernel = self.net_bias = self.t_kernel = self.t_bias = self.s_kernel = self.s_bias = None

    def build(self, dim: int | None = None, ctx=None) -> None:
        ctx = ctx or default_context()
        if dim is not None and dim != self.dim:
            raise ValueError(f"RNVP was created for dim={self.dim} but got input of width {dim}")
        rng = init.param_rng(self.uid)
        D, H = self.dim, self.h_sizes[0]
        self.net_kernel = ctx.to_device(init.glorot_uniform(rng, (D, H)))  # rnvp.py:28-29
        self.net_bias = ctx.zeros((H,))
        self.t_kernel = ctx.to_device(init.glorot_uniform(rng, (H, D)))  # rnvp.py:30
        self.t_bias = ctx.zeros((D,))
        self.s_kernel = ctx.to_device(init.glorot_uniform(rng, (H, D)))  # rnvp.py:31
        self.s_bias = ctx.zeros((D,))
        self.built = True

    def variables(self) -> dict:
        return {"net_kernel": self.net_kernel, "net_bias": self.net_bias, "t_kernel": self.t_kernel,
                "t_bias": self.t_bias, "s_kernel": self.s_kernel, "s_bias": self.s_bias}

    def fill_step(self, st: L.FlowStep, ptr_of=lambda a: a.ptr) -> None:
        st.kind, st.parity, st.hidden = self.kind, 0, self.h_sizes[0]
        st.w1, st.b1 = ptr_of(self.net_kernel), ptr_of(self.net_bias)
        st.ws, st.bs = ptr_of(self.s_kernel), ptr_of(self.s_bias)
        st.wt, st.bt = ptr_of(self.t_kernel), ptr_of(self.t_bias)
        st.ld2 = self.dim
        st.m1 = st.m2 = None

    def forward(self, z, mask=None):
        return self._single_step(z, bern=None if mask is None else [mask])
