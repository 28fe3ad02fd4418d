"""Flat HBM arenas for the trainable parameters of the path.

The reference hands Lightning one optimizer param-group per injected module (``modules/model.py:152-155``) and lets
DDP bucket the gradients (``train.py:98-109``).  Here every trainable tensor is a view into ONE contiguous fp32
parameter arena with a matching gradient arena, so that the per-step work is a handful of launches regardless of
the number of sites: one NCCL all-reduce over the gradient arena, one fused AdamW(+EMA) pass per hyper-parameter
group, one multi-tensor repack of the bf16 tensor-core operands.  ``Parameter`` objects keep their identity, names
and shapes, so ``state_dict`` / checkpoint layout are unchanged.
"""
from __future__ import annotations

import weakref
from typing import Iterable, Optional

import torch
from torch import nn

from . import _lib
from .lora import PackedOperands, _LoRABase, _pack_sites, lora_modules

ALIGN = 4  # elements (16 bytes of fp32): every view stays 128-bit aligned


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class ParamArena:
    """Contiguous fp32 storage for a list of parameters, grouped by optimizer hyper-parameters.

    ``param_groups`` follows torch's convention (list of dicts with ``params`` and optional overrides) -- exactly
    what ``config_module`` returns.  Groups with equal overrides are merged into one contiguous range.
    """

    def __init__(self, param_groups: list[dict], device=None):
        merged: dict[tuple, list[nn.Parameter]] = {}
        order: list[tuple] = []
        for g in param_groups:
            key = tuple(sorted((k, v) for k, v in g.items() if k != "params"))
            if key not in merged:
                merged[key] = []
                order.append(key)
            merged[key].extend(g["params"])
        params = [p for k in order for p in merged[k]]
        if not params:
            raise ValueError("ParamArena needs at least one parameter")
        seen = set()
        for p in params:
            if id(p) in seen:
                raise ValueError("a parameter appears in more than one group")
            seen.add(id(p))
            if p.dtype != torch.float32:
                raise _lib.SdtError("trainable parameters must be fp32 masters")
        device = device or params[0].device
        self.device = device
        self.ranges: list[tuple[dict, int, int]] = []     # (overrides, begin, end) per merged group
        self.slots: list[tuple[nn.Parameter, int, int]] = []  # (param, offset, numel)
        off = 0
        for key in order:
            begin = off
            for p in merged[key]:
                self.slots.append((p, off, p.numel()))
                off = _round_up(off + p.numel(), ALIGN)
            self.ranges.append((dict(key), begin, off))
        self.numel = off
        self.params = torch.zeros(off, dtype=torch.float32, device=device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, o, n in self.slots:
                self.params[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.params[o:o + n].view(p.shape)
                p.grad = self.grads[o:o + n].view(p.shape)

    def zero_grad(self) -> None:
        self.grads.zero_()
        for p, o, n in self.slots:          # survive optimizers that set grads to None
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * o:
                p.grad = self.grads[o:o + n].view(p.shape)

    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        for q, o, n in self.slots:
            if q is p:
                return self.grads[o:o + n].view(p.shape)
        raise KeyError("parameter is not in this arena")


class LoraArena(ParamArena):
    """``ParamArena`` + the bf16 operand arena of every LoRA site of ``module``.

    After construction the sites' backward kernels accumulate dA / dB straight into the gradient arena and read
    their tensor-core operands from one bf16 arena that ``pack()`` refreshes with a single launch.
    """

    def __init__(self, module: nn.Module, param_groups: list[dict], compute_dtype: Optional[torch.dtype] = None):
        super().__init__(param_groups)
        self.sites: list[tuple[str, _LoRABase]] = [(n, m) for n, m in lora_modules(module) if m.lora_A.requires_grad]
        in_arena = {id(p) for p, _, _ in self.slots}
        self.sites = [(n, m) for n, m in self.sites if id(m.lora_A) in in_arena and id(m.lora_B) in in_arena]
        if compute_dtype is None:
            # the 16-bit format of the frozen base decides (unet.to(float16) -> fp16, the reference's `precision: 16`); fp32
            # bases are driven through autocast and switch the arena on their first launch (set_compute_dtype)
            wd = self.sites[0][1].weight.dtype if self.sites else torch.bfloat16
            compute_dtype = wd if wd in (torch.bfloat16, torch.float16) else torch.bfloat16
        self._build_operands(compute_dtype)

    def _build_operands(self, dtype: torch.dtype) -> None:
        self.compute_dtype = dtype
        total = sum(PackedOperands.numel(m.in_features, m.out_features, m.r) for _, m in self.sites)
        self.packed = torch.zeros(max(total, 8), dtype=dtype, device=self.device)
        off = 0
        recs = []
        self._max_elems = 1
        for _, m in self.sites:
            n = PackedOperands.numel(m.in_features, m.out_features, m.r)
            ops = PackedOperands(m.in_features, m.out_features, m.r, self.device, self.packed[off:off + n])
            off += n
            m._ops, m._ops_external = ops, True
            m._arena_ref = weakref.ref(self)
            m._grad_A, m._grad_B = self.grad_view(m.lora_A), self.grad_view(m.lora_B)
            recs.append(ops.site(m.lora_A.detach(), m.lora_B.detach()))
            self._max_elems = max(self._max_elems, ops.R * (m.in_features + m.out_features))
        self._recs = recs
        self._sites_dev: Optional[torch.Tensor] = None
        if recs:
            self.pack()

    def set_compute_dtype(self, dtype: torch.dtype) -> None:
        """Switch the packed operands between bf16 and fp16 (re-packs every site; not meant for the inner loop)."""
        if dtype not in (torch.bfloat16, torch.float16):
            raise _lib.SdtError(f"tensor-core operands are bfloat16 or float16, not {dtype}")
        if dtype != self.compute_dtype:
            self._build_operands(dtype)

    def pack(self) -> None:
        """fp32 masters -> bf16 A_p / At_p / B_p / Bt_p of every site, one launch (call after optimizer.step)."""
        if not self._recs:
            return
        if self._sites_dev is None:
            self._sites_dev = _pack_sites(self._recs, self._max_elems, self.device, self.compute_dtype)
        else:
            _lib.check(_lib.load().sdt_lora_pack(self._sites_dev.data_ptr(), len(self._recs), self._max_elems,
                                                 _lib.dtype_code(self.compute_dtype), _lib.stream_ptr()), "sdt_lora_pack")
