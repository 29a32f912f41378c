"""MNF VGG-style conv net -- BASELINE.json configs[4] ("MNF VGG-style conv net, synthetic 32x32x3
batch 256 data-parallel training with NCCL grad allreduce").  The reference ships no such model
(tf_mnf/models/__init__.py:1-3); SURVEY.md section 8(d) C5 defines it from the reference's layer
classes the way tf_mnf/models/mnf_lenet.py:14-33 builds LeNet:

    [MNFConv2D(f, 3, SAME) ReLU] x 2, MaxPool 2x2   for f in (64, 128, 256)
    Flatten (4*4*256 = 4096)  MNFDense(512) ReLU  MNFDense(n_classes) Softmax
"""
from __future__ import annotations

from typing import Any, Sequence

from ..layers import MNFConv2D, MNFDense
from .stock import Flatten, ReLU, ReLUMaxPool2D, Sequential, Softmax


class MNFVGG(Sequential):
    def __init__(self, filters: Sequence[int] = (64, 128, 256), n_hidden: int = 512, n_classes: int = 10,
                 **kwargs: Any) -> None:
        layers: list = []
        for f in filters:
            layers += [MNFConv2D(int(f), 3, padding="SAME", **kwargs), ReLU(),
                       MNFConv2D(int(f), 3, padding="SAME", **kwargs), ReLUMaxPool2D(padding="SAME")]
        layers += [Flatten(), MNFDense(int(n_hidden), **kwargs), ReLU(), MNFDense(int(n_classes), **kwargs), Softmax()]
        super().__init__(layers)
