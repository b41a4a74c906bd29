"""Data-parallel exchange over the library's own collective entry points (include/ssdgeom.h, ssdg_comm_*: NCCL
bound at run time).  No torch: a TensorFlow or plain-Python host ships the 128-byte id with whatever channel it has
-- ``tcp_broadcast`` below is a minimal one -- and sums the loss words / mining histograms in place on a stream.

    id = comm.tcp_broadcast(comm.unique_id() if rank == 0 else None, world, rank, addr, port)
    cm = comm.Comm(id, world, rank)             # collective
    cm.allreduce(device_array, stream)          # in place, asynchronous
"""
from __future__ import annotations

import ctypes as C
import socket
import time

import numpy as np

from . import _native as N
from . import device as D

_CODES = {np.dtype(np.float32): N.F32, np.dtype(np.float64): N.F64, np.dtype(np.int32): N.I32,
          np.dtype(np.int64): N.I64}


def available() -> bool:
    return N.lib().ssdg_comm_available(None) == N.OK


def nccl_version() -> int:
    v = C.c_int(0)
    N.check(N.lib().ssdg_comm_available(C.byref(v)), "comm_available")
    return v.value


def unique_id() -> bytes:
    buf = C.create_string_buffer(N.COMM_ID_BYTES)
    N.check(N.lib().ssdg_comm_unique_id(buf), "comm_unique_id")
    return buf.raw


def tcp_broadcast(payload: bytes | None, world: int, rank: int, addr: str = "127.0.0.1", port: int = 29617,
                  nbytes: int = N.COMM_ID_BYTES, timeout: float = 120.0) -> bytes:
    """Rank 0 serves ``payload`` (``nbytes`` long) to the other ``world - 1`` ranks over TCP; returns it everywhere."""
    if world == 1:
        return payload
    if rank == 0:
        if payload is None or len(payload) != nbytes:
            raise ValueError("rank 0 must pass the %d-byte payload" % nbytes)
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as srv:
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            srv.settimeout(timeout)
            for _ in range(world - 1):
                conn, _peer = srv.accept()
                with conn:
                    conn.sendall(payload)
        return payload
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as s:
                chunks, got = [], 0
                while got < nbytes:
                    c = s.recv(nbytes - got)
                    if not c:
                        raise ConnectionError("short read")
                    chunks.append(c)
                    got += len(c)
                return b"".join(chunks)
        except (ConnectionError, OSError):
            if time.monotonic() > deadline:
                raise
            time.sleep(0.05)


class Comm:
    """One rank of an NCCL communicator on the current device (ssdg_set_device first)."""

    def __init__(self, uid: bytes, world: int, rank: int):
        if len(uid) != N.COMM_ID_BYTES:
            raise ValueError("unique id must be %d bytes" % N.COMM_ID_BYTES)
        h = C.c_void_p()
        N.check(N.lib().ssdg_comm_init_rank(C.byref(h), C.c_char_p(uid), int(world), int(rank)), "comm_init_rank")
        self.handle, self.world, self.rank = h.value, int(world), int(rank)

    def allreduce(self, buf, stream=None):
        """Sum ``buf`` (DeviceArray or any __cuda_array_interface__ object: float32/64, int32/64) over the ranks, in
        place, ordered on ``stream``."""
        buf = D.as_device(buf)
        N.check(N.lib().ssdg_comm_allreduce_sum(self.handle, buf.ptr, buf.size, _CODES[buf.dtype],
                                                D.stream_handle(stream)), "comm_allreduce_sum")
        return buf

    def allreduce_multi(self, bufs, stream=None):
        """Several in-place sums as one NCCL group (one launch)."""
        bufs = [D.as_device(b) for b in bufs]
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.ptr for b in bufs])
        counts = (C.c_int64 * n)(*[b.size for b in bufs])
        codes = (C.c_int32 * n)(*[_CODES[b.dtype] for b in bufs])
        N.check(N.lib().ssdg_comm_allreduce_sum_multi(self.handle, n, ptrs, counts, codes, D.stream_handle(stream)),
                "comm_allreduce_sum_multi")
        return bufs

    def close(self):
        if self.handle:
            N.lib().ssdg_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
