"""IAF / MAF host objects (reference: tf_mnf/flows/maf.py:13-63).

`IAF.forward` (== `MAF.inverse`, maf.py:47-53,63) is one MADE pass, y = x*exp(s)+t with an
optional feature reversal and log-det sum(s); it runs as the fused kernel in csrc/flows.cu.
`MAF.forward` / `IAF.inverse` (the sequential decode, maf.py:33-45) cannot run in the
reference either (App. B-4) and is outside the hot path."""
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
        if len(self.net.h_sizes) != 1:
            raise NotImplementedError(
                f"h_sizes={self.net.h_sizes}: the fused sm_100a flow kernel supports exactly one hidden layer"
            )

    def build(self, dim: int, ctx=None) -> None:
        if not self.net.built:
            self.net.build(dim, None, ctx)
        self.dim = dim
        self.built = True

    def variables(self) -> dict:
        l1, l2 = self.net.layers
        return {"w1": l1.kernel, "b1": l1.bias, "w2": l2.kernel, "b2": l2.bias}

    def fill_step(self, st: L.FlowStep, ptr_of=lambda a: a.ptr) -> None:
        l1, l2 = self.net.layers
        D = self.dim
        st.kind, st.parity, st.hidden = self.kind, self.parity, self.net.h_sizes[0]
        st.w1, st.b1 = ptr_of(l1.kernel), ptr_of(l1.bias)
        st.ws, st.wt = ptr_of(l2.kernel), ptr_of(l2.kernel) + 4 * D  # [H,2D]: s = cols [0,D), t = cols [D,2D)
        st.bs, st.bt = ptr_of(l2.bias), ptr_of(l2.bias) + 4 * D
        st.ld2 = 2 * D
        st.m1 = l1.mask.ptr if l1.mask is not None else None
        st.m2 = l2.mask.ptr if l2.mask is not None else None

    # maf.py:47-53
    def inverse(self, x):
        return self._single_step(x)

    # maf.py:33-45 (broken in the reference, App. B-4)
    def forward(self, z):
        raise NotImplementedError("MAF.forward (sequential decode) is outside the MNF hot path (SURVEY.md 8f-4)")


class IAF(MAF):
    """maf.py:56-63: swaps forward and inverse."""

    def forward(self, z):
        return self._single_step(z)

    def inverse(self, x):
        raise NotImplementedError("IAF.inverse (sequential decode) is outside the MNF hot path (SURVEY.md 8f-4)")
