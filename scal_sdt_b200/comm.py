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
