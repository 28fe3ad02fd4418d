"""Fused elementwise neighbours of the LoRA GEMMs: GEGLU after ``ff.net.0.proj``.

torch evaluates ``h, gate = proj.chunk(2, -1); h * gelu(gate)`` as strided, non-vectorised kernels plus a concat of the two
gradient halves in backward; here each direction is ONE 128-bit vectorised pass over the ``[M, 2I]`` projection.
"""
from __future__ import annotations

import torch

from . import _lib


class _GEGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj):
        I = proj.shape[-1] // 2
        p2 = proj.reshape(-1, 2 * I).contiguous()
        out = torch.empty(p2.shape[0], I, dtype=proj.dtype, device=proj.device)
        _lib.check(_lib.load().sdt_geglu(p2.data_ptr(), None, out.data_ptr(), p2.shape[0], I, 0, _lib.dtype_code(proj.dtype),
                                         _lib.stream_ptr()), "sdt_geglu")
        ctx.save_for_backward(p2)
        ctx.lead = proj.shape[:-1]
        return out.view(*ctx.lead, I)

    @staticmethod
    def backward(ctx, dout):
        (p2,) = ctx.saved_tensors
        I = p2.shape[1] // 2
        d2 = dout.reshape(-1, I).contiguous()
        if d2.dtype != p2.dtype:
            d2 = d2.to(p2.dtype)
        dproj = torch.empty_like(p2)
        _lib.check(_lib.load().sdt_geglu(p2.data_ptr(), d2.data_ptr(), dproj.data_ptr(), p2.shape[0], I, 1,
                                         _lib.dtype_code(p2.dtype), _lib.stream_ptr()), "sdt_geglu")
        return dproj.view(*ctx.lead, 2 * I)


def geglu(proj: torch.Tensor) -> torch.Tensor:
    """``h * gelu(gate)`` for ``proj = [h | gate]`` along the last dimension (CUDA tensors, bf16 or fp32)."""
    _lib.require_cuda(proj)
    return _GEGLU.apply(proj)
