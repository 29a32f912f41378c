"""NumPy-backed stand-in for the slice of TensorFlow 2 that tf-mnf's hot path touches.

TEST INFRASTRUCTURE ONLY.  TensorFlow cannot be installed in the build container
(no network), so the reference's *own source files* (imported unmodified from
/root/reference by tests/golden/make_golden.py) are executed on top of this shim in
float64 to produce the golden vectors under tests/golden/*.npz.  The shim supplies
only well-specified primitives (matmul, conv2d, exp, reduce_sum, Keras' build-on-first
-call protocol ...); all MNF control flow, draw order and quirks come from the
reference code itself.

Deliberate deviations from real TF (each needed because the reference at HEAD does not
run on real TF):
  * tf.nn.conv2d defaults strides=1, padding="VALID" (real TF: required args; B-1).
  * keras.layers.Layer.__init__ swallows unknown kwargs such as padding= (B-1).
  * random draws come from an injectable, logged stream (tf.random.normal,
    keras.backend.random_binomial) so that oracle and CUDA tests can replay them.
"""
from __future__ import annotations

import sys
import types

import numpy as np

DTYPE = np.float64
float32 = np.float32
float64 = np.float64


class _NoiseLog:
    """Ordered log of every random draw (kind, shape, values)."""

    def __init__(self) -> None:
        self.rng = np.random.RandomState(1234)
        self.init_rng = np.random.RandomState(4321)
        self.draws: list[tuple[str, tuple, np.ndarray]] = []

    def reset(self, seed: int = 1234, init_seed: int | None = None) -> None:
        self.rng = np.random.RandomState(seed)
        if init_seed is not None:
            self.init_rng = np.random.RandomState(init_seed)
        self.draws = []


NOISE = _NoiseLog()


def _shape(shape) -> tuple:
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


# ----------------------------------------------------------------------------- core
class Variable(np.ndarray):
    def __new__(cls, value, trainable: bool = True, **_):
        obj = np.array(value, dtype=DTYPE).view(cls)
        obj.trainable = trainable
        return obj

    def __array_finalize__(self, obj):
        self.trainable = getattr(obj, "trainable", True)

    def numpy(self):
        return np.asarray(self)


class Module:
    def __init__(self, **kwargs) -> None:
        pass


class TensorShape(tuple):
    pass


Tensor = np.ndarray


def shape(x):
    return np.shape(x)


def zeros(shape_, dtype=None):
    return np.zeros(_shape(shape_), DTYPE)


def ones(shape_, dtype=None):
    return np.ones(_shape(shape_), DTYPE)


def zeros_like(x):
    return np.zeros_like(x)


def tile(x, multiples):
    return np.tile(x, multiples)


def exp(x):
    return np.exp(x)


def sqrt(x):
    return np.sqrt(x)


def square(x):
    return np.square(x)


def abs(x):  # noqa: A001
    return np.abs(x)


def tanh(x):
    return np.tanh(x)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x)))


def transpose(x):
    return np.transpose(x)


def stack(xs, axis=0):
    return np.stack(xs, axis=axis)


def squeeze(x):
    return np.squeeze(x)


def reshape(x, shape_):
    return np.reshape(x, shape_)


def reduce_sum(x, axis=None):
    return np.sum(x, axis=axis)


def reduce_mean(x, axis=None):
    return np.mean(x, axis=axis)


def matmul(a, b):
    return np.matmul(a, b)


def tensordot(a, b, axes):
    return np.tensordot(a, b, axes=axes)


def clip_by_value(x, lo, hi):
    return np.clip(x, lo, hi)


def reverse(x, axis):
    return np.flip(x, axis=tuple(axis))


def split(x, n, axis=-1):
    return np.split(x, n, axis=axis)


def cast(x, dtype):
    if dtype is bool:
        return np.asarray(x).astype(bool)
    return np.asarray(x).astype(DTYPE)


def function(fn):
    return fn


math = types.SimpleNamespace(
    log=np.log,
    softplus=lambda x: np.logaddexp(0.0, x),
)
linalg = types.SimpleNamespace(matvec=lambda a, b: np.matmul(a, b))


def _normal(shape_=()):
    shp = _shape(shape_)
    val = NOISE.rng.standard_normal(shp).astype(DTYPE)
    NOISE.draws.append(("normal", shp, val))
    return val


random = types.SimpleNamespace(normal=_normal, set_seed=lambda s: None)


def _conv2d(input, filters, strides=1, padding="VALID"):  # noqa: A002
    """NHWC x HWIO cross-correlation, stride `strides`, VALID or SAME (TF semantics)."""
    x = np.asarray(input)
    w = np.asarray(filters)
    s = strides if isinstance(strides, int) else strides[1]
    B, H, W, C = x.shape
    kh, kw, C2, F = w.shape
    assert C == C2
    if padding == "SAME":
        Ho, Wo = -(-H // s), -(-W // s)
        ph = max((Ho - 1) * s + kh - H, 0)
        pw = max((Wo - 1) * s + kw - W, 0)
        x = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    else:
        Ho, Wo = (H - kh) // s + 1, (W - kw) // s + 1
    out = np.zeros((B, Ho, Wo, F), DTYPE)
    for i in range(kh):
        for j in range(kw):
            patch = x[:, i : i + (Ho - 1) * s + 1 : s, j : j + (Wo - 1) * s + 1 : s, :]
            out += np.tensordot(patch, w[i, j], axes=([3], [0]))
    return out


nn = types.SimpleNamespace(conv2d=_conv2d)


# ---------------------------------------------------------------------------- keras
def _fans(shp):
    if len(shp) == 1:
        return shp[0], shp[0]
    if len(shp) == 2:
        return shp[0], shp[1]
    rf = int(np.prod(shp[:-2]))
    return shp[-2] * rf, shp[-1] * rf


class GlorotNormal:
    """Truncated normal (+-2 sigma), sigma = sqrt(2/(fan_in+fan_out))/0.8796..."""

    def __call__(self, shape_):
        shp = _shape(shape_)
        fi, fo = _fans(shp)
        std = np.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
        n = int(np.prod(shp))
        vals = NOISE.init_rng.standard_normal(n)
        bad = np.abs(vals) > 2
        while bad.any():
            vals[bad] = NOISE.init_rng.standard_normal(int(bad.sum()))
            bad = np.abs(vals) > 2
        return (vals * std).reshape(shp).astype(DTYPE)


def _glorot_uniform(shp):
    fi, fo = _fans(shp)
    lim = np.sqrt(6.0 / (fi + fo))
    return NOISE.init_rng.uniform(-lim, lim, size=shp).astype(DTYPE)


class Layer:
    def __init__(self, **kwargs) -> None:
        # real Keras rejects unknown kwargs (e.g. padding=); see module docstring
        self._extra_kwargs = kwargs
        self.built = False

    def build(self, input_shape) -> None:
        self.built = True

    def __call__(self, x, *args, **kwargs):
        if not getattr(self, "built", False):
            self.build(TensorShape(np.shape(x)))
            self.built = True
        return self.call(x, *args, **kwargs)


_ACT = {None: lambda v: v, "tanh": np.tanh, "relu": lambda v: np.maximum(v, 0), "linear": lambda v: v}


class Dense(Layer):
    def __init__(self, units, activation=None, **kwargs) -> None:
        super().__init__(**kwargs)
        self.units = units
        self.activation = activation

    def build(self, input_shape) -> None:
        n_in = input_shape[-1]
        self.kernel = Variable(_glorot_uniform((n_in, self.units)))
        self.bias = Variable(np.zeros(self.units))
        self.built = True

    @property
    def weights(self):
        return [self.kernel, self.bias]

    def call(self, x):
        return _ACT[self.activation](np.matmul(x, self.kernel) + self.bias)


class ReLU(Layer):
    def call(self, x):
        return np.maximum(x, 0)


class Flatten(Layer):
    def call(self, x):
        return np.reshape(x, (np.shape(x)[0], -1))


class Softmax(Layer):
    def call(self, x):
        e = np.exp(x - np.max(x, axis=-1, keepdims=True))
        return e / e.sum(axis=-1, keepdims=True)


class MaxPool2D(Layer):
    def __init__(self, pool_size=2, strides=None, padding="VALID", **kwargs) -> None:
        super().__init__(**kwargs)
        self.pool = pool_size
        self.strides = strides or pool_size
        self.padding = padding.upper()

    def call(self, x):
        B, H, W, C = x.shape
        p, s = self.pool, self.strides
        if self.padding == "SAME":
            Ho, Wo = -(-H // s), -(-W // s)
            ph = max((Ho - 1) * s + p - H, 0)
            pw = max((Wo - 1) * s + p - W, 0)
            x = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)),
                       constant_values=-np.inf)
        else:
            Ho, Wo = (H - p) // s + 1, (W - p) // s + 1
        out = np.full((B, Ho, Wo, C), -np.inf)
        for i in range(p):
            for j in range(p):
                out = np.maximum(out, x[:, i : i + (Ho - 1) * s + 1 : s, j : j + (Wo - 1) * s + 1 : s, :])
        return out


class BatchNormalization(Layer):
    """Inference mode (what `model(x)` gives): moving mean 0, var 1, eps 1e-3."""

    def call(self, x):
        return x / np.sqrt(1.0 + 1e-3)


class Conv2D(Layer):
    def __init__(self, *a, **k) -> None:
        super().__init__()


class Sequential(Layer):
    def __init__(self, layers=None, **kwargs) -> None:
        super().__init__(**kwargs)
        self.layers = list(layers or [])

    def call(self, x):
        for lyr in self.layers:
            x = lyr(x)
        return x


class Model(Sequential):
    pass


def _random_binomial(shape_, p=0.5):
    shp = _shape(shape_)
    u = NOISE.rng.uniform(size=shp)
    val = (u <= p).astype(DTYPE)  # keras backend: where(uniform <= p, 1, 0)
    NOISE.draws.append(("bernoulli", shp, val))
    return val


def _mk(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_layers = _mk(
    "tensorflow.keras.layers",
    Layer=Layer, Dense=Dense, ReLU=ReLU, Flatten=Flatten, Softmax=Softmax, MaxPool2D=MaxPool2D,
    BatchNormalization=BatchNormalization, Conv2D=Conv2D,
)
_inits = _mk("tensorflow.keras.initializers", GlorotNormal=GlorotNormal)
_backend = _mk("tensorflow.keras.backend", random_binomial=_random_binomial)
keras = _mk(
    "tensorflow.keras",
    layers=_layers, initializers=_inits, backend=_backend, Sequential=Sequential, Model=Model,
)
