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
        if not 1 <= len(h_sizes) <= L.MNF_FLOW_MAX_HIDDEN:
            raise NotImplementedError(f"h_sizes={h_sizes}: 1 to {L.MNF_FLOW_MAX_HIDDEN} hidden layers are supported")
        if activation not in L.ACTIVATIONS or activation in (None, "default"):
            raise NotImplementedError(f"activation={activation!r}: one of 'tanh', 'relu', 'sigmoid', 'linear'")
        self.dim, self.h_sizes, self.activation = int(dim), h_sizes, activation
        self.net_kernels: list = []
        self.net_biases: list = []
        self.t_kernel = self.t_bias = self.s_kernel = self.s_bias = None

    # the first net layer under the names the one-hidden-layer version always had
    @property
    def net_kernel(self):
        return self.net_kernels[0] if self.net_kernels else None

    @property
    def net_bias(self):
        return self.net_biases[0] if self.net_biases else None

    def build(self, dim: int | None = None, ctx=None) -> None:
        ctx = ctx or default_context()
        if dim is not None and dim != self.dim:
            raise ValueError(f"RNVP was created for dim={self.dim} but got input of width {dim}")
        rng = init.param_rng(self.uid)
        D = self.dim
        sizes = [D, *self.h_sizes]
        self.net_kernels = [ctx.to_device(init.glorot_uniform(rng, (sizes[i], sizes[i + 1]))) for i in range(len(self.h_sizes))]  # rnvp.py:28-29
        self.net_biases = [ctx.zeros((h,)) for h in self.h_sizes]
        H = self.h_sizes[-1]
        self.t_kernel = ctx.to_device(init.glorot_uniform(rng, (H, D)))  # rnvp.py:30
        self.t_bias = ctx.zeros((D,))
        self.s_kernel = ctx.to_device(init.glorot_uniform(rng, (H, D)))  # rnvp.py:31
        self.s_bias = ctx.zeros((D,))
        self.built = True

    def variables(self) -> dict:
        out = {"net_kernel": self.net_kernels[0], "net_bias": self.net_biases[0]}
        for i in range(1, len(self.h_sizes)):
            out[f"net_kernel_{i}"], out[f"net_bias_{i}"] = self.net_kernels[i], self.net_biases[i]
        out.update({"t_kernel": self.t_kernel, "t_bias": self.t_bias, "s_kernel": self.s_kernel, "s_bias": self.s_bias})
        return out

    def fill_step(self, st: L.FlowStep, ptr_of=lambda a: a.ptr) -> None:
        st.kind, st.parity, st.hidden = self.kind, 0, self.h_sizes[0]
        st.w1, st.b1 = ptr_of(self.net_kernels[0]), ptr_of(self.net_biases[0])
        st.ws, st.bs = ptr_of(self.s_kernel), ptr_of(self.s_bias)
        st.wt, st.bt = ptr_of(self.t_kernel), ptr_of(self.t_bias)
        st.ld2 = self.dim
        st.m1 = st.m2 = None
        st.n_extra = len(self.h_sizes) - 1
        st.activation = 0 if self.activation == "tanh" else L.ACTIVATIONS[self.activation]
        for i in range(1, len(self.h_sizes)):
            st.extra_h[i - 1] = self.h_sizes[i]
            st.extra_w[i - 1], st.extra_b[i - 1] = ptr_of(self.net_kernels[i]), ptr_of(self.net_biases[i])
            st.extra_m[i - 1] = None

    def forward(self, z, mask=None):
        return self._single_step(z, bern=None if mask is None else [mask])
