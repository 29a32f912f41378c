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
