"""Data-parallel gradient exchange: one NCCL all-reduce (mean) over the flat gradient arena per step.

Replaces the Lightning ``DDPStrategy`` reducer (``train.py:98-109``, ``modules/utils/fix_ddp.py:5-11``).  The
communicator lives inside ``libsdt_b200.so`` (``sdt_comm_*``); ``torch.distributed`` is only the rendezvous that
carries the 128-byte NCCL unique id from rank 0 to the others.  On CPU-only machines (the gloo tests) the same
``GradExchange`` API reduces through ``torch.distributed`` so that the host logic is testable without a GPU; that
branch never runs on a CUDA tensor.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


class GradExchange:
    def __init__(self, rank: int = 0, world_size: int = 1):
        self.rank, self.world = rank, world_size
        self._native = False

    @classmethod
    def from_torch_distributed(cls, device: torch.device) -> "GradExchange":
        """Bootstrap the native communicator from an initialised ``torch.distributed`` group."""
        if not dist.is_initialized():
            return cls(0, 1)
        self = cls(dist.get_rank(), dist.get_world_size())
        if self.world == 1:
            return self
        if device.type == "cuda":
            lib = _lib.load()
            uid = (ctypes.c_ubyte * 128)()
            if self.rank == 0:
                _lib.check(lib.sdt_comm_unique_id(uid), "sdt_comm_unique_id")
            obj = [bytes(uid)]
            dist.broadcast_object_list(obj, src=0)
            buf = (ctypes.c_ubyte * 128).from_buffer_copy(obj[0])
            torch.cuda.set_device(device)
            _lib.check(lib.sdt_comm_init(buf, self.rank, self.world), "sdt_comm_init")
            self._native = True
        return self

    def all_reduce_mean_(self, flat: torch.Tensor) -> torch.Tensor:
        """In-place mean over ranks (DDP semantics) on the caller's current stream."""
        if self.world == 1:
            return flat
        if flat.is_cuda:
            if not self._native:
                raise _lib.SdtError("GradExchange: native communicator not initialised")
            _lib.check(_lib.load().sdt_allreduce(flat.data_ptr(), flat.numel(), _lib.dtype_code(flat.dtype),
                                                 _lib.stream_ptr()), "sdt_allreduce")
            return flat
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)      # host-logic test path (gloo, CPU tensors only)
        flat.div_(self.world)
        return flat

    def close(self) -> None:
        if self._native:
            _lib.check(_lib.load().sdt_comm_destroy(), "sdt_comm_destroy")
            self._native = False


class OverlappedExchange:
    """Chunked gradient all-reduce overlapped with backward (DDP's bucketing, ``train.py:98-109``, on the flat arena).

    The arena is cut into contiguous chunks of about ``chunk_bytes`` at parameter boundaries.  Every parameter carries a
    post-accumulate-grad hook; when the last parameter of a chunk has received its gradient the chunk's ``ncclAllReduce``
    is issued on a side stream that waits for an event recorded on the compute stream at that moment.  Chunks fill from the
    END of the arena (backward runs the network in reverse), so the first all-reduces are on the wire while most of backward
    is still ahead.  ``finish()`` issues whatever is left (parameters that received no gradient) and makes the compute stream
    wait for the side stream.  Everything is stream-ordered: it can be captured into a CUDA graph."""

    def __init__(self, arena, exchange: GradExchange, chunk_bytes: int = 64 << 20):
        self.arena, self.exchange = arena, exchange
        # CPU arenas (the gloo host-logic tests) reduce each chunk synchronously: same partition, same hooks, no streams
        self.stream = torch.cuda.Stream(device=arena.device) if torch.device(arena.device).type == "cuda" else None
        target = max(1, chunk_bytes // 4)
        self.chunks: list[dict] = []
        cur = None
        for p, off, n in arena.slots:
            if cur is None or cur["end"] - cur["begin"] >= target:
                cur = {"begin": off, "end": off, "params": 0, "pending": 0, "sent": False}
                self.chunks.append(cur)
            cur["end"] = off + n
            cur["params"] += 1
            p.register_post_accumulate_grad_hook(lambda _p, c=cur: self._arrived(c))
        for i, c in enumerate(self.chunks):         # a chunk owns the alignment padding up to the next chunk's start
            c["end"] = self.chunks[i + 1]["begin"] if i + 1 < len(self.chunks) else arena.numel
        self.launched = 0

    def begin(self) -> None:
        for c in self.chunks:
            c["pending"], c["sent"] = c["params"], False
        self.launched = 0

    def _send(self, c: dict) -> None:
        c["sent"] = True
        self.launched += 1
        if self.stream is None:
            self.exchange.all_reduce_mean_(self.arena.grads[c["begin"]:c["end"]])
            return
        cur = torch.cuda.current_stream(self.arena.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            self.exchange.all_reduce_mean_(self.arena.grads[c["begin"]:c["end"]])

    def _arrived(self, c: dict) -> None:
        c["pending"] -= 1
        if c["pending"] == 0 and not c["sent"]:
            self._send(c)

    def finish(self) -> None:
        for c in reversed(self.chunks):
            if not c["sent"]:
                self._send(c)
        if self.stream is not None:
            torch.cuda.current_stream(self.arena.device).wait_stream(self.stream)
