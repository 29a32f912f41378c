"""Multi-process host logic on CPU (gloo, world_size 2): row sharding and the data-parallel
gradient rule (likelihood gradient summed over ranks, KL gradient contributed once)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from tf_mnf_b200.parallel import row_shard


def test_row_shard_partitions_exactly():
    for n in (0, 1, 7, 10240, 10241, 12345):
        for world in (1, 2, 3, 4, 8):
            spans = [row_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_shard(10, 2, 2)
    # BASELINE config C3 at 1/2/4/8 GPUs: 10240 rows split evenly
    assert [row_shard(10240, r, 8) for r in (0, 7)] == [(0, 1280), (8960, 10240)]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    from tf_mnf_b200.parallel import allreduce_gradients

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(100 + rank)
        like = [torch.tensor(rng.standard_normal(s).astype(np.float32)) for s in [(5, 3), (7,), (2, 2, 2)]]
        kl = [torch.tensor(np.random.RandomState(7).standard_normal(s).astype(np.float32)) for s in [(5, 3), (7,), (2, 2, 2)]]
        # variant A: KL added after the sum (contributed once)
        a = [t.clone() for t in like]
        allreduce_gradients(a, kl)
        # variant B (what Trainer does): KL pre-scaled by 1/world and folded into the summed buffer
        b = [t + k / world for t, k in zip(like, kl)]
        allreduce_gradients(b, None)
        out.put((rank, [t.numpy() for t in a], [t.numpy() for t in b]))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shapes = [(5, 3), (7,), (2, 2, 2)]
    like = [[np.random.RandomState(100 + r).standard_normal(s).astype(np.float32) for s in shapes][i] for r in range(world) for i in range(0)]
    per_rank = []
    for r in range(world):
        rng = np.random.RandomState(100 + r)
        per_rank.append([rng.standard_normal(s).astype(np.float32) for s in shapes])
    kl = [np.random.RandomState(7).standard_normal(s).astype(np.float32) for s in shapes]
    for i in range(len(shapes)):
        expect = per_rank[0][i] + per_rank[1][i] + kl[i]
        for rank, a, b in res:
            np.testing.assert_allclose(a[i], expect, rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(b[i], expect, rtol=1e-6, atol=1e-6)
