"""Multi-GPU host logic.  One process per GPU.

* Predictive path: rows (MC sample x image) are independent given the parameters, so every
  rank owns a contiguous row range and there is NO data-path collective; Philox counters are
  indexed by the GLOBAL row (layer.row_offset) so results do not depend on the GPU count.
* Training: data parallel.  The likelihood gradient is averaged over the global batch, the KL
  gradient is identical on every rank and must be contributed once: one all-reduce (NCCL over
  NVLink on GPUs, gloo in the CPU tests) over a flat fp32 buffer.

The collective itself is libmnf_b200's mnf_allreduce_grads (tf_mnf_b200/comm.py: NCCL bound inside
the C ABI); this module holds the host-side rules and is free of torch.
"""
from __future__ import annotations

from typing import Any, Sequence


def row_shard(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of n_rows owned by `rank`; sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} not in [0, {world})")
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_model_rows(model: Any, n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Set every MNF layer's global row offset for this rank's shard and return the slice."""
    lo, hi = row_shard(n_rows, rank, world)
    model.set_row_offset(lo)
    return lo, hi


def allreduce_gradients(like_grads: Sequence[Any], kl_grads: Sequence[Any] | None = None, comm: Any = None) -> None:
    """In place: like_grads[i] <- sum_ranks(like_grads[i])  (+ kl_grads[i] once).

    The list form of what `Trainer` does on its flat device buffer, on HOST arrays (NumPy): `like_grads`
    hold each rank's gradient of ITS OWN shard's loss already scaled by 1 / global batch
    (mnf_softmax_xent's inv_global_batch does this), so that the plain SUM over ranks is the gradient
    of the global-batch mean; the KL gradient is identical on every rank and is added once behind the
    sum (or, as Trainer does, pre-scaled by 1 / world and folded into the summed buffer).  The arrays
    are flattened into ONE buffer so a single collective is issued (lenet: 1.70 M parameters =
    6.8 MB).  `comm`: any object with `.world` and `.allreduce_host(flat: np.ndarray)` (in-place sum)."""
    import numpy as np

    tensors = list(like_grads)
    if not tensors:
        return
    flat = np.concatenate([np.asarray(t, np.float32).reshape(-1) for t in tensors])
    if comm is not None and getattr(comm, "world", 1) > 1:
        comm.allreduce_host(flat)
    off = 0
    for i, t in enumerate(tensors):
        n = t.size
        t[...] = flat[off:off + n].reshape(t.shape)
        if kl_grads is not None and kl_grads[i] is not None:
            t += kl_grads[i]
        off += n
