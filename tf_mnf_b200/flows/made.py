"""MADE / MaskedDense host objects (reference: tf_mnf/flows/made.py:14-125).

The masked MLP itself runs inside the fused flow kernel (csrc/flows.cu); these classes own
the parameters and the NumPy mask construction.  Quirk App. B-2: in the reference the masks
are overwritten with ones on the first call, so `mask_mode="ones"` (default) reproduces the
reference as it runs and `"autoregressive"` gives the masks set_masks() intended."""
from __future__ import annotations

import numpy as np

from .. import init
from ..runtime import Context, DeviceArray, default_context


def build_masks(n_in: int, h_sizes, n_outputs: int = 2, seed: int = 0, shuffle: bool = False):
    """made.py:93-125: degrees from RandomState(seed); hidden masks use <=, output mask <."""
    rng = np.random.RandomState(seed)
    deg = {-1: rng.permutation(n_in) if shuffle else np.arange(n_in)}
    for lyr, size in enumerate(h_sizes):
        deg[lyr] = rng.randint(deg[lyr - 1].min(), n_in - 1, size=size)
    L = len(h_sizes)
    masks = [deg[l - 1][:, None] <= deg[l][None, :] for l in range(L)]
    masks.append(deg[L - 1][:, None] < deg[-1][None, :])
    if n_outputs > 1:
        masks[-1] = np.concatenate([masks[-1]] * n_outputs, axis=1)
    return [m.astype(np.float32) for m in masks], deg


class MaskedDense:
    """Dense layer whose kernel is multiplied by a fixed 0/1 mask (made.py:14-30)."""

    def __init__(self, n_out: int) -> None:
        self.units = int(n_out)
        self.kernel: DeviceArray | None = None
        self.bias: DeviceArray | None = None
        self.mask: DeviceArray | None = None  # None == all ones
        self.mask_host: np.ndarray | None = None

    def build(self, n_in: int, rng, ctx: Context) -> None:
        self.kernel = ctx.to_device(init.glorot_uniform(rng, (n_in, self.units)))
        self.bias = ctx.zeros((self.units,))

    def set_mask(self, mask: np.ndarray | None, ctx: Context) -> None:
        self.mask_host = None if mask is None else (np.asarray(mask) != 0).astype(np.float32)
        self.mask = None if mask is None else ctx.to_device(self.mask_host)

    @property
    def weights(self):
        return [self.kernel, self.bias]


class MADE:
    def __init__(self, n_outputs: int = 1, h_sizes=(), num_masks: int = 1, shuffle: bool = False,
                 mask_mode: str = "ones") -> None:
        if num_masks != 1:
            raise NotImplementedError("num_masks > 1 (mask ensembles) is not used by the MNF layers")
        if mask_mode not in ("ones", "autoregressive"):
            raise ValueError(f"mask_mode must be 'ones' or 'autoregressive', got {mask_mode!r}")
        self.n_outputs = int(n_outputs)
        self.h_sizes = tuple(int(h) for h in h_sizes)
        self.shuffle = shuffle
        self.num_masks = num_masks
        self.seed = 0
        self.mask_mode = mask_mode
        self.m: dict[int, np.ndarray] = {}
        self.layers: list[MaskedDense] = []
        self.n_in: int | None = None
        self.built = False

    def build(self, n_in: int, rng=None, ctx: Context | None = None) -> None:
        ctx = ctx or default_context()
        self.n_in = int(n_in)
        rng = rng or init.param_rng(init.next_uid())
        sizes = [self.n_in, *self.h_sizes, self.n_in * self.n_outputs]
        self.layers = [MaskedDense(s) for s in sizes[1:]]
        for lyr, fan_in in zip(self.layers, sizes[:-1]):
            lyr.build(fan_in, rng, ctx)
        self.set_masks(ctx)
        self.built = True

    def set_masks(self, ctx: Context | None = None) -> None:
        ctx = ctx or default_context()
        if self.n_in < 2:
            raise ValueError("MADE needs at least 2 inputs (made.py:108 samples randint(0, n_in - 1))")
        masks, self.m = build_masks(self.n_in, self.h_sizes, self.n_outputs, self.seed, self.shuffle)
        for lyr, mk in zip(self.layers, masks):
            lyr.set_mask(mk if self.mask_mode == "autoregressive" else None, ctx)
