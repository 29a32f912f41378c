"""MNFLeNet (reference: tf_mnf/models/mnf_lenet.py:11-39)."""
from __future__ import annotations

from typing import Any

from ..layers import MNFConv2D, MNFDense
from .stock import Flatten, ReLU, ReLUMaxPool2D, Sequential, Softmax


class MNFLeNet(Sequential):
    """conv(s1,5,VALID) ReLU MaxPool(SAME) conv(s2,5,VALID) ReLU MaxPool(SAME) Flatten
    MNFDense(s3) ReLU MNFDense(s4) Softmax.  ReLU+MaxPool run as one fused kernel."""

    def __init__(self, s1: int = 20, s2: int = 50, s3: int = 500, s4: int = 10, **kwargs: Any) -> None:
        c1 = MNFConv2D(s1, 5, padding="VALID", **kwargs)
        c2 = MNFConv2D(s2, 5, padding="VALID", **kwargs)
        d1 = MNFDense(s3, **kwargs)
        d2 = MNFDense(s4, **kwargs)
        super().__init__([c1, ReLUMaxPool2D(padding="SAME"), c2, ReLUMaxPool2D(padding="SAME"), Flatten(),
                          d1, ReLU(), d2, Softmax()])
