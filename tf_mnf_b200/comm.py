"""Data-parallel communicator over libmnf_b200's NCCL entry points (include/mnf_b200.h:
mnf_comm_unique_id / mnf_comm_init / mnf_allreduce_grads).  One process per GPU.

The reference has no distributed code; the collective sits where it applies its gradients
(notebooks/lenet_mnist.py:114-115).  The only host-side job is to ship rank 0's 128-byte NCCL id to
the other ranks: a one-shot TCP rendezvous on (MASTER_ADDR, MNF_COMM_PORT | MASTER_PORT + 211) --
the variables torchrun exports -- with no dependency on torch."""
from __future__ import annotations

import ctypes as C
import os
import socket
import struct
import time
from typing import Callable

from . import _lib as L
from .runtime import Context, DeviceArray, default_context

_MAGIC = b"MNFCOMM1"


def _recv_exact(conn: socket.socket, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        part = conn.recv(n - len(buf))
        if not part:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += part
    return buf


def exchange_id(rank: int, world: int, addr: str, port: int, make_id: Callable[[], bytes], timeout: float = 120.0) -> bytes:
    """Rank 0 creates the id (make_id()) and serves it to the world - 1 other ranks; returns the id on
    every rank.  Each client names its rank so that a stray connection cannot consume a slot."""
    if world == 1:
        return make_id()
    deadline = time.monotonic() + timeout
    if rank == 0:
        ident = make_id()
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as srv:
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            served: set[int] = set()
            while len(served) < world - 1:
                srv.settimeout(max(deadline - time.monotonic(), 0.01))
                try:
                    conn, _ = srv.accept()
                except socket.timeout:
                    raise TimeoutError(f"rendezvous: only ranks {sorted(served)} of {world - 1} joined within {timeout}s") from None
                with conn:
                    conn.settimeout(10.0)
                    try:
                        hello = _recv_exact(conn, len(_MAGIC) + 8)
                    except (ConnectionError, socket.timeout):
                        continue
                    if hello[:len(_MAGIC)] != _MAGIC:
                        continue
                    peer, peer_world = struct.unpack("<ii", hello[len(_MAGIC):])
                    if peer_world != world or not (0 < peer < world) or peer in served:
                        continue
                    conn.sendall(ident)
                    served.add(peer)
        return ident
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as conn:
                conn.sendall(_MAGIC + struct.pack("<ii", rank, world))
                return _recv_exact(conn, L.MNF_COMM_ID_BYTES)
        except (ConnectionRefusedError, ConnectionError, socket.timeout, OSError):
            if time.monotonic() > deadline:
                raise TimeoutError(f"rendezvous: rank {rank} could not reach {addr}:{port} within {timeout}s") from None
            time.sleep(0.05)


def rendezvous_endpoint() -> tuple[str, int]:
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = os.environ.get("MNF_COMM_PORT")
    if port is None:
        port = (int(os.environ.get("MASTER_PORT", "29500")) + 211 - 1024) % (65536 - 1024) + 1024
    return addr, int(port)


class Communicator:
    """`allreduce(flat)`: in-place sum of a flat fp32 DeviceArray over the ranks, enqueued on the
    array's context stream (no host synchronisation)."""

    def __init__(self, ctx: Context | None = None, rank: int | None = None, world: int | None = None,
                 id_bytes: bytes | None = None) -> None:
        self.ctx = ctx or default_context()
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else int(world)
        lib = self.ctx.lib

        def make_id() -> bytes:
            buf = (C.c_uint8 * L.MNF_COMM_ID_BYTES)()
            L.check(lib.mnf_comm_unique_id(buf))
            return bytes(buf)

        if id_bytes is None:
            id_bytes = exchange_id(self.rank, self.world, *rendezvous_endpoint(), make_id)
        if len(id_bytes) != L.MNF_COMM_ID_BYTES:
            raise ValueError(f"NCCL unique id must be {L.MNF_COMM_ID_BYTES} bytes, got {len(id_bytes)}")
        h = C.c_void_p()
        buf = (C.c_uint8 * L.MNF_COMM_ID_BYTES).from_buffer_copy(id_bytes)
        L.check(lib.mnf_comm_init(self.ctx.handle, buf, self.rank, self.world, C.byref(h)))
        self.handle = h

    @property
    def nccl_version(self) -> int:
        v = C.c_int()
        L.check(self.ctx.lib.mnf_comm_nccl_version(C.byref(v)))
        return v.value

    def allreduce(self, flat: DeviceArray) -> None:
        if flat.dtype != "f4":
            raise ValueError("allreduce expects a float32 array")
        L.check(flat.ctx.lib.mnf_allreduce_grads(self.handle, flat.ctx.handle, C.c_void_p(flat.ptr), flat.size))

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.ctx.sync()
            self.ctx.lib.mnf_comm_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
