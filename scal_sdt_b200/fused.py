"""Fused elementwise neighbours of the LoRA GEMMs: GEGLU after ``ff.net.0.proj``.

torch evaluates ``h, gate = proj.chunk(2, -1); h * gelu(gate)`` as strided, non-vectorised kernels plus a concat of the two
gradient halves in backward; here each direction is ONE 128-bit vectorised pass over the ``[M, 2I]`` projection.
"""
from __future__ import annotations

import contextlib
import logging

import torch

from . import _lib

_log = logging.getLogger("scal_sdt_b200.fused")
_torch_only = False
_noted: set = set()


@contextlib.contextmanager
def torch_only():
    """Route every helper of this file through plain torch ops.  Used by the comparators that time / check the reference's
    own eager torch path on the same GPU (``bench.py`` ``torch_gpu_baseline``, ``tests/test_gpu_sd15_step.py``): none of this
    repo's kernels may run on that arm."""
    global _torch_only
    saved, _torch_only = _torch_only, True
    try:
        yield
    finally:
        _torch_only = saved


def fused_enabled() -> bool:
    return not _torch_only


def _note_torch_path(what: str, x: torch.Tensor) -> None:
    """A CUDA tensor took the torch path of a host-model helper (shape / dtype / layout / trainable affine did not qualify):
    say so once per helper and reason instead of silently changing the kernels that run."""
    if not x.is_cuda or _torch_only:
        return
    key = (what, x.dtype, x.dim())
    if key not in _noted:
        _noted.add(key)
        _log.warning("%s: %s %s tensor does not qualify for the fused kernel; using torch ops for it (host-model code)",
                     what, tuple(x.shape), x.dtype)


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


def geglu_supported(proj: torch.Tensor) -> bool:
    return fused_enabled() and proj.is_cuda and proj.dtype in (torch.bfloat16, torch.float32)


class _GroupNormNHWC(torch.autograd.Function):
    """GroupNorm with frozen affine parameters (+ optional SiLU) on a channels-last bf16 tensor; ``chan_bias`` [B, C] (optional)
    is added to the input on the fly (the time-embedding add of a ResNet block)."""

    @staticmethod
    def forward(ctx, x, chan_bias, gamma, beta, groups, eps, silu):
        B, C, H, W = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        stats = torch.empty(B, groups, 2, dtype=torch.float32, device=x.device)
        cb = chan_bias.contiguous() if chan_bias is not None else None
        lib = _lib.load()
        ws = torch.empty(int(lib.sdt_group_norm_workspace_floats(B, groups)), dtype=torch.float32, device=x.device)   # per-CTA partials
        _lib.check(lib.sdt_group_norm_nhwc(x.data_ptr(), _lib.ptr(cb), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(),
                                           y.data_ptr(), B, H * W, C, groups, eps, int(silu), ws.data_ptr(), ws.numel(),
                                           _lib.stream_ptr()), "sdt_group_norm_nhwc")
        ctx.save_for_backward(x, gamma, beta, stats, *([cb] if cb is not None else []))
        ctx.cfg = (groups, eps, silu)
        ctx.has_bias = cb is not None
        return y

    @staticmethod
    def backward(ctx, dout):
        x, gamma, beta, stats, *rest = ctx.saved_tensors
        cb = rest[0] if ctx.has_bias else None
        groups, eps, silu = ctx.cfg
        B, C, H, W = x.shape
        d = dout.contiguous(memory_format=torch.channels_last)
        if d.dtype != x.dtype:
            d = d.to(x.dtype)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        lib = _lib.load()
        ws = torch.empty(int(lib.sdt_group_norm_workspace_floats(B, groups)), dtype=torch.float32, device=x.device)
        _lib.check(lib.sdt_group_norm_nhwc_bwd(x.data_ptr(), _lib.ptr(cb), d.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                               stats.data_ptr(), ws.data_ptr(), ws.numel(), dx.data_ptr(), B, H * W, C, groups,
                                               eps, int(silu), _lib.stream_ptr()), "sdt_group_norm_nhwc_bwd")
        dcb = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            dcb = dx.sum(dim=(2, 3), dtype=torch.float32).to(cb.dtype)      # the backward of the broadcast add
        return dx, dcb, None, None, None, None, None


def group_norm_nhwc_supported(norm: torch.nn.GroupNorm, x: torch.Tensor) -> bool:
    c, g = norm.num_channels, norm.num_groups
    return (fused_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
            and not norm.weight.requires_grad and not norm.bias.requires_grad and c % 8 == 0 and c // g >= 8 and g <= 128
            and c <= 4096)


def group_norm_act(norm: torch.nn.GroupNorm, x: torch.Tensor, silu: bool, chan_bias: torch.Tensor | None = None) -> torch.Tensor:
    """``silu(norm(x + chan_bias[:, :, None, None]))`` / without SiLU / without the bias: fused channels-last kernels when they
    apply, torch otherwise (host model code)."""
    if group_norm_nhwc_supported(norm, x) and (chan_bias is None or (chan_bias.is_cuda and chan_bias.dtype == x.dtype
                                                                       and chan_bias.shape == x.shape[:2])):
        cache = getattr(norm, "_sdt_affine_f32", None)
        key = (norm.weight.data_ptr(), norm.weight._version, norm.bias._version)
        if cache is None or cache[0] != key:
            cache = (key, norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous())
            norm._sdt_affine_f32 = cache
        return _GroupNormNHWC.apply(x, chan_bias, cache[1], cache[2], norm.num_groups, float(norm.eps), bool(silu))
    _note_torch_path("group_norm_act", x)
    if chan_bias is not None:
        x = x + chan_bias[:, :, None, None]
    y = norm(x)
    return torch.nn.functional.silu(y) if silu else y


class _AddLayerNorm(torch.autograd.Function):
    """``xs = x + res; y = layer_norm(xs)`` (res optional) with frozen affine parameters, one launch per direction."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        C = x.shape[-1]
        x2 = x.reshape(-1, C).contiguous()
        M = x2.shape[0]
        y = torch.empty_like(x2)
        stats = torch.empty(M, 2, dtype=torch.float32, device=x.device)
        if res is not None:
            r2 = res.reshape(-1, C).contiguous()
            xs = torch.empty_like(x2)
        else:
            r2, xs = None, None
        _lib.check(_lib.load().sdt_layer_norm_fwd(x2.data_ptr(), _lib.ptr(r2), gamma.data_ptr(), beta.data_ptr(), _lib.ptr(xs),
                                                  y.data_ptr(), stats.data_ptr(), M, C, eps, _lib.stream_ptr()), "sdt_layer_norm_fwd")
        ctx.save_for_backward(xs if res is not None else x2, gamma, stats)
        ctx.has_res = res is not None
        ctx.shape = x.shape
        if res is not None:
            return xs.view(x.shape), y.view(x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, *grads):
        xs, gamma, stats = ctx.saved_tensors
        M, C = xs.shape
        if ctx.has_res:
            dxs, dy = grads
        else:
            dxs, dy = None, grads[0]
        if dy is None:                               # only the residual stream was used downstream
            return dxs, (dxs if ctx.has_res else None), None, None, None
        dy2 = dy.reshape(M, C).contiguous()
        if dy2.dtype != xs.dtype:
            dy2 = dy2.to(xs.dtype)
        d2 = None
        if dxs is not None:
            d2 = dxs.reshape(M, C).contiguous()
            if d2.dtype != xs.dtype:
                d2 = d2.to(xs.dtype)
        dx = torch.empty_like(xs)
        _lib.check(_lib.load().sdt_layer_norm_bwd(xs.data_ptr(), dy2.data_ptr(), _lib.ptr(d2), gamma.data_ptr(), stats.data_ptr(),
                                                  dx.data_ptr(), M, C, _lib.stream_ptr()), "sdt_layer_norm_bwd")
        dx = dx.view(ctx.shape)
        return dx, (dx if ctx.has_res else None), None, None, None


def layer_norm_supported(norm: torch.nn.LayerNorm, x: torch.Tensor) -> bool:
    c = x.shape[-1]
    return (fused_enabled() and x.is_cuda and x.dtype == torch.bfloat16 and len(norm.normalized_shape) == 1 and norm.normalized_shape[0] == c
            and norm.weight is not None and norm.bias is not None and not norm.weight.requires_grad
            and not norm.bias.requires_grad and c % 8 == 0 and c <= 2048 and x.numel() > 0)


def _ln_affine(norm):
    cache = getattr(norm, "_sdt_affine_f32", None)
    key = (norm.weight.data_ptr(), norm.weight._version, norm.bias._version)
    if cache is None or cache[0] != key:
        cache = (key, norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous())
        norm._sdt_affine_f32 = cache
    return cache[1], cache[2]


def add_layer_norm(norm: torch.nn.LayerNorm, x: torch.Tensor, res: torch.Tensor | None = None):
    """``res is None``: ``norm(x)``.  Otherwise ``(x + res, norm(x + res))`` -- the residual add and the LayerNorm that follows
    it in the transformer block as one pass (torch otherwise: host-model code for trainable norms / CPU / fp32)."""
    if layer_norm_supported(norm, x) and (res is None or (res.shape == x.shape and res.dtype == x.dtype and res.is_cuda)):
        g, b = _ln_affine(norm)
        return _AddLayerNorm.apply(x, res, g, b, float(norm.eps))
    _note_torch_path("add_layer_norm", x)
    if res is None:
        return norm(x)
    xs = x + res
    return xs, norm(xs)


class _ResidualBiasAdd(torch.autograd.Function):
    """``a + b + bias[None, :, None, None]`` on channels-last bf16 tensors, frozen bias (f32 [C]); one vectorised pass."""

    @staticmethod
    def forward(ctx, a, b, bias_f32):
        B, C, H, W = a.shape
        out = torch.empty_like(a, memory_format=torch.channels_last)
        _lib.check(_lib.load().sdt_residual_bias_add(a.data_ptr(), b.data_ptr(), bias_f32.data_ptr(), out.data_ptr(), B * H * W, C,
                                                     _lib.stream_ptr()), "sdt_residual_bias_add")
        return out

    @staticmethod
    def backward(ctx, dout):
        return dout, dout, None


def residual_bias_add(a: torch.Tensor, b: torch.Tensor, bias_f32: torch.Tensor) -> torch.Tensor:
    """``a + b + bias`` (per channel) for two channels-last bf16 NCHW tensors; torch otherwise (host-model code)."""
    cl = torch.channels_last
    if (fused_enabled() and a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape == b.shape and a.dim() == 4
            and a.shape[1] % 8 == 0 and a.is_contiguous(memory_format=cl) and b.is_contiguous(memory_format=cl) and a.numel() > 0):
        return _ResidualBiasAdd.apply(a, b, bias_f32)
    _note_torch_path("residual_bias_add", a)
    return a + b + bias_f32.to(a.dtype)[None, :, None, None]
