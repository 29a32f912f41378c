"""MNFFeedForward (reference: tf_mnf/models/mnf_feed_forward.py:11-28)."""
from __future__ import annotations

from typing import Any

from ..layers import MNFDense
from .stock import BatchNormalization, ReLU, Sequential


class MNFFeedForward(Sequential):
    def __init__(self, layer_sizes=(100, 50, 10), activation=ReLU, **kwargs: Any) -> None:
        layers: list = []
        for dim in layer_sizes[:-1]:
            layers.extend([MNFDense(dim, **kwargs), activation(), BatchNormalization()])
        super().__init__([*layers, MNFDense(layer_sizes[-1], **kwargs)])
