"""Multi-GPU host logic.  One process per GPU.

* Predictive path: rows (MC sample x image) are independent given the parameters, so every
  rank owns a contiguous row range and there is NO data-path collective; Philox counters are
  indexed by the GLOBAL row (layer.row_offset) so results do not depend on the GPU count.
* Training: data parallel.  The likelihood gradient is averaged over the global batch, the KL
  gradient is identical on every rank and must be contributed once: one all-reduce (NCCL over
  NVLink on GPUs, gloo in the CPU tests) over a flat fp32 buffer.

torch.distributed is plumbing here (process group + collective); tensors handed to it are
zero-copy views of DeviceArrays (`torch.as_tensor(arr)` via __cuda_array_interface__).
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


def allreduce_gradients(like_grads: Sequence[Any], kl_grads: Sequence[Any] | None = None, group: Any = None) -> None:
    """In place: like_grads[i] <- sum_ranks(like_grads[i]) / world  (+ kl_grads[i] once).

    `like_grads` hold each rank's gradient of ITS OWN shard's mean loss times (local rows /
    global rows) -- i.e. already scaled so that the plain SUM over ranks is the gradient of the
    global-batch mean (softmax_xent's inv_global_batch does this).  Tensors are torch tensors
    (CPU for gloo, CUDA views for NCCL); they are flattened into ONE buffer so a single
    collective is issued (lenet: 1.70 M parameters = 6.8 MB)."""
    import torch
    import torch.distributed as dist

    tensors = [t for t in like_grads]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for i, t in enumerate(tensors):
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        if kl_grads is not None and kl_grads[i] is not None:
            t.add_(kl_grads[i])
        off += n
