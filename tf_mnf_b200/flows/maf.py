"""IAF / MAF host objects (reference: tf_mnf/flows/maf.py:13-63).

`IAF.forward` (== `MAF.inverse`, maf.py:47-53,63) is one MADE pass, y = x*exp(s)+t with an
optional feature reversal and log-det sum(s); it runs as the fused kernel in csrc/flows.cu.
`MAF.forward` / `IAF.inverse` (the sequential decode, maf.py:33-45; it cannot run in the reference,
App. B-4) is the loop the reference spells out, as one kernel (mnf_maf_decode)."""
from __future__ import annotations

from .. import _lib as L
from .base import Flow
from .made import MADE


class MAF(Flow):
    kind = L.MNF_FLOW_IAF

    def __init__(self, parity, net: MADE | None = None, h_sizes=(30,), mask_mode: str = "ones") -> None:
        super().__init__()
        self.parity = int(bool(parity))
        self.net = net or MADE(n_outputs=2, h_sizes=h_sizes, mask_mode=mask_mode)
        if not 1 <= len(self.net.h_sizes) <= L.MNF_FLOW_MAX_HIDDEN:
            raise NotImplementedError(f"h_sizes={self.net.h_sizes}: 1 to {L.MNF_FLOW_MAX_HIDDEN} hidden layers are supported")

    def build(self, dim: int, ctx=None) -> None:
        if not self.net.built:
            self.net.build(dim, None, ctx)
        self.dim = dim
        self.built = True

    @property
    def h_sizes(self):
        return self.net.h_sizes

    def variables(self) -> dict:
        """w1, b1, ..., w{L+1}, b{L+1}: the MaskedDense layers of the MADE net in order (made.py:77-88)."""
        out = {}
        for i, lyr in enumerate(self.net.layers, 1):
            out[f"w{i}"], out[f"b{i}"] = lyr.kernel, lyr.bias
        return out

    def fill_step(self, st: L.FlowStep, ptr_of=lambda a: a.ptr) -> None:
        layers = self.net.layers
        l1, last, mid = layers[0], layers[-1], layers[1:-1]
        D = self.dim
        st.kind, st.parity, st.hidden = self.kind, self.parity, self.net.h_sizes[0]
        st.w1, st.b1 = ptr_of(l1.kernel), ptr_of(l1.bias)
        st.ws, st.wt = ptr_of(last.kernel), ptr_of(last.kernel) + 4 * D  # [H,2D]: s = cols [0,D), t = cols [D,2D)
        st.bs, st.bt = ptr_of(last.bias), ptr_of(last.bias) + 4 * D
        st.ld2 = 2 * D
        st.m1 = l1.mask.ptr if l1.mask is not None else None
        st.m2 = last.mask.ptr if last.mask is not None else None
        st.n_extra, st.activation = len(mid), 0
        for i, lyr in enumerate(mid):  # further hidden layers: the general kernels (csrc/flows_generic.cu)
            st.extra_h[i] = lyr.units
            st.extra_w[i], st.extra_b[i] = ptr_of(lyr.kernel), ptr_of(lyr.bias)
            st.extra_m[i] = lyr.mask.ptr if lyr.mask is not None else None

    # maf.py:47-53
    def inverse(self, x):
        return self._single_step(x)

    def _decode(self, z):
        """maf.py:33-45, the dim-pass sequential decode (mnf_maf_decode): returns (x, log_dets)."""
        import ctypes as C

        from ..runtime import as_device, default_context

        z = as_device(z, default_context())
        if z.ndim != 2:
            raise ValueError(f"flow input must be [batch, dim], got shape {z.shape}")
        B, D = z.shape
        if not self.built:
            self.build(D, z.ctx)
        if D != self.dim:
            raise ValueError(f"flow has dim {self.dim} but the input has width {D}")
        st = L.FlowStep()
        self.fill_step(st)
        x, ld = z.ctx.empty((B, D)), z.ctx.empty((B,))
        L.check(z.ctx.lib.mnf_maf_decode(z.ctx.handle, B, D, C.byref(st), C.c_void_p(z.ptr), C.c_void_p(x.ptr), C.c_void_p(ld.ptr)))
        return x, ld

    # maf.py:33-45 (as the reference spells it out; it cannot run there, App. B-4)
    def forward(self, z):
        return self._decode(z)


class IAF(MAF):
    """maf.py:56-63: swaps forward and inverse."""

    def forward(self, z):
        return self._single_step(z)

    def inverse(self, x):
        return self._decode(x)
