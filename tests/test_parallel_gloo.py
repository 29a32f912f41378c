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


class _GlooComm:
    """Test-side stand-in for tf_mnf_b200.comm.Communicator on CPU: the same protocol over gloo."""

    def __init__(self, world):
        self.world = world

    def allreduce_host(self, flat):
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(flat)  # shares memory: in place
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from tf_mnf_b200.parallel import allreduce_gradients

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = _GlooComm(world)
        rng = np.random.RandomState(100 + rank)
        like = [rng.standard_normal(s).astype(np.float32) for s in [(5, 3), (7,), (2, 2, 2)]]
        kl = [np.random.RandomState(7).standard_normal(s).astype(np.float32) for s in [(5, 3), (7,), (2, 2, 2)]]
        # variant A: KL added after the sum (contributed once)
        a = [t.copy() for t in like]
        allreduce_gradients(a, kl, comm)
        # variant B (what Trainer does): KL pre-scaled by 1/world and folded into the summed buffer
        b = [t + k / world for t, k in zip(like, kl)]
        allreduce_gradients(b, None, comm)
        out.put((rank, a, b))
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


def _rdv_worker(rank, world, port, out):
    from tf_mnf_b200.comm import exchange_id

    made = []

    def make_id():
        made.append(rank)
        return bytes((7 * i + 3) % 256 for i in range(128))

    ident = exchange_id(rank, world, "127.0.0.1", port, make_id, timeout=60.0)
    out.put((rank, ident, made))


@pytest.mark.parametrize("world", [2, 3])
def test_comm_rendezvous_ships_rank0s_id(world):
    """tf_mnf_b200.comm.exchange_id: rank 0 alone creates the 128-byte id; every rank ends up with it
    (late joiners retry until the server is up; a stray connection does not consume a slot)."""
    import multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_rdv_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs[1:]:
        p.start()  # the clients first: they must wait for the server
    with pytest.raises(OSError):
        socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
    procs[0].start()
    res = sorted((out.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes((7 * i + 3) % 256 for i in range(128))
    assert [r[1] for r in res] == [want] * world
    assert [r[2] for r in res] == [[0]] + [[]] * (world - 1)


def test_rendezvous_endpoint_follows_torchrun_env(monkeypatch):
    from tf_mnf_b200.comm import rendezvous_endpoint

    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MASTER_PORT", "29511")
    monkeypatch.delenv("MNF_COMM_PORT", raising=False)
    assert rendezvous_endpoint() == ("127.0.0.1", 29722)
    monkeypatch.setenv("MNF_COMM_PORT", "31000")
    assert rendezvous_endpoint() == ("127.0.0.1", 31000)
    monkeypatch.setenv("MASTER_PORT", "65500")
    monkeypatch.delenv("MNF_COMM_PORT")
    assert 1024 <= rendezvous_endpoint()[1] < 65536
