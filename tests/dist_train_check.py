"""Data-parallel training check, run under torchrun (one process per GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/dist_train_check.py

Every rank owns a contiguous slice of the global batch of a small MNFVGG (BASELINE.json
configs[4]); Philox counters use the GLOBAL row.  After one all-reduce (sum) over the flat fp32
gradient buffer every rank must hold the gradient a single process computes on the whole batch
(likelihood gradient averaged over the global batch, KL gradient contributed once), and after the
Adam step all ranks must hold identical parameters.  Prints one line "DIST_TRAIN_OK ..." on rank 0."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(M, MNFVGG):
    M.set_seed(0)  # also resets layer ids: same parameters and same noise keys on every rank / rebuild
    return MNFVGG(filters=(8, 16, 32), n_hidden=64, n_classes=10, flow_h_sizes=[8], std_init=1.0, max_std=1)


def main() -> None:
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFVGG
    from tf_mnf_b200.parallel import row_shard
    from tf_mnf_b200.train import Trainer

    M.runtime.set_default_device(local)
    GB = 32
    rng = np.random.RandomState(0)
    x = rng.uniform(size=(GB, 32, 32, 3)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=GB)]
    lo, hi = row_shard(GB, rank, world)

    model = build(M, MNFVGG)
    model.set_row_offset(lo)
    tr = Trainer(model, loss="xent", kl_weight=1.0 / 500, world=world)
    loss, kl = tr.loss_and_grads(x[lo:hi], y[lo:hi])
    tr.allreduce()
    got = {k: g.numpy().copy() for k, _, g in tr._named_params()}
    tr.apply()
    params = np.concatenate([p.numpy().ravel() for _, p, _ in tr._named_params()])

    ref_model = build(M, MNFVGG)
    ref = Trainer(ref_model, loss="xent", kl_weight=1.0 / 500, world=1)
    ref.loss_and_grads(x, y)
    worst = 0.0
    for k, _, g in ref._named_params():
        a, b = got[k], g.numpy()
        scale = max(np.max(np.abs(b)), 1e-12)
        err = float(np.max(np.abs(a - b)) / scale)
        worst = max(worst, err)
        assert err <= 2e-5, (k, err)  # measured 9e-7: the shards' fp32 sums differ from the single-process order only
    if world > 1:
        t = torch.from_numpy(params).cuda()
        mx, mn = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert torch.equal(mx, mn), "parameters diverged across ranks after the Adam step"
        # the per-rank losses sum to the global-batch mean loss (each is scaled by 1/global batch)
        lt = torch.tensor([float(loss.numpy()[0])], device="cuda")
        dist.all_reduce(lt)
        dist.barrier()
    if rank == 0:
        print(f"DIST_TRAIN_OK world={world} n_params={params.size} worst_rel_grad_err={worst:.2e}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
