"""Fused AdamW over the flat parameter arena (``modules/model.py:33-64``, ``configs/lora.yaml:66-73``).

Same update rule as ``torch.optim.AdamW`` (decoupled weight decay, bias correction, eps outside the sqrt of the
bias-corrected second moment); one launch per hyper-parameter group instead of one torch multi-tensor pass per
reference param group.  Optionally folds the gradient unscale (1/world after a sum all-reduce) and the EMA lerp of
``modules/ema.py`` into the same pass.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from .arena import ParamArena


class FlatAdamW:
    def __init__(self, arena: ParamArena, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.arena = arena
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.exp_avg = torch.zeros_like(arena.params)
        self.exp_avg_sq = torch.zeros_like(arena.params)
        self.step_count = 0
        # torch-style view of the groups so LR schedulers / lr_scale code can read and write `lr`
        self.param_groups = []
        for overrides, begin, end in arena.ranges:
            g = dict(self.defaults)
            g.update(overrides)
            if "beta1" in g or "beta2" in g:
                g["betas"] = (g.pop("beta1", g["betas"][0]), g.pop("beta2", g["betas"][1]))
            g["_range"] = (begin, end)
            self.param_groups.append(g)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, ema_shadow: Optional[torch.Tensor] = None,
             ema_one_minus_decay: float = 0.0) -> None:
        lib = _lib.load()
        self.step_count += 1
        t = self.step_count
        a = self.arena
        for g in self.param_groups:
            begin, end = g["_range"]
            if end <= begin:
                continue
            b1, b2 = g["betas"]
            hyper = (ctypes.c_float * 7)(g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** t, 1.0 - b2 ** t)
            sh = 0 if ema_shadow is None else ema_shadow.data_ptr() + 4 * begin
            _lib.check(lib.sdt_adamw_flat(a.params.data_ptr() + 4 * begin, a.grads.data_ptr() + 4 * begin,
                                          self.exp_avg.data_ptr() + 4 * begin, self.exp_avg_sq.data_ptr() + 4 * begin,
                                          end - begin, hyper, None, grad_scale, sh, ema_one_minus_decay, None,
                                          _lib.stream_ptr()), "sdt_adamw_flat")
        # the masters changed in place: LoraArena.pack() (explicit, one launch) refreshes the bf16 operands

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "_range"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
