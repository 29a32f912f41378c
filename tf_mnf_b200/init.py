"""Host-side parameter initialisers (NumPy), matching what the reference's `build` methods
draw: keras GlorotNormal for layer variables (mnf_dense.py:52, mnf_conv.py:63, planar.py:18)
and Keras Dense defaults (glorot_uniform kernel, zero bias) inside the flows."""
from __future__ import annotations

import numpy as np

_TRUNC_STD = 0.87962566103423978  # std of a unit normal truncated to [-2, 2]


def fans(shape) -> tuple[int, int]:
    shape = tuple(int(s) for s in shape)
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def glorot_normal(rng: np.random.RandomState, shape) -> np.ndarray:
    fi, fo = fans(shape)
    std = np.sqrt(2.0 / (fi + fo)) / _TRUNC_STD
    n = int(np.prod(shape))
    v = rng.standard_normal(n)
    bad = np.abs(v) > 2
    while bad.any():
        v[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(v) > 2
    return (v * std).reshape(shape).astype(np.float32)


def glorot_uniform(rng: np.random.RandomState, shape) -> np.ndarray:
    fi, fo = fans(shape)
    lim = np.sqrt(6.0 / (fi + fo))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


_uid = [0]


def next_uid() -> int:
    """Layer / flow ids: they select Philox sub-streams, so creation order fixes the noise."""
    _uid[0] += 1
    return _uid[0]


def reset_uids() -> None:
    _uid[0] = 0


_seed = [0]


def set_seed(seed: int) -> None:
    """tf.random.set_seed / np.random.seed stand-in (lenet_mnist.py:36-37)."""
    _seed[0] = int(seed)
    reset_uids()


def global_seed() -> int:
    return _seed[0]


def param_rng(uid: int) -> np.random.RandomState:
    return np.random.RandomState((global_seed() * 1000003 + uid * 7919 + 12345) % (2**32))
