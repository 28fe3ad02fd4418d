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

        # device copy of the per-group hyper-parameters (lr, beta1, beta2, eps, wd, bias_corr1, bias_corr2): lets a captured
        # CUDA graph follow the step count / LR schedule without being re-captured
        self._hyper_host = None
        self._hyper_dev = None

    def refresh_device_hyper(self, step: int) -> None:
        """Write the hyper-parameters of optimizer step number `step` (1-based) to the device table (async copy)."""
        n = len(self.param_groups)
        if self._hyper_host is None:
            self._hyper_dev = torch.zeros(n, 8, dtype=torch.float32, device=self.arena.params.device)
            self._hyper_host = _lib.PinnedRing(self._hyper_dev)      # event-guarded slots: the host may run steps ahead

        def fill(host):
            for i, g in enumerate(self.param_groups):
                b1, b2 = g["betas"]
                host[i, :7] = torch.tensor([g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** step, 1.0 - b2 ** step])
        self._hyper_host.push(fill)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, ema_shadow: Optional[torch.Tensor] = None,
             ema_one_minus_decay: float = 0.0, use_device_hyper: bool = False,
             ema_one_minus_decay_dev: Optional[torch.Tensor] = None) -> None:
        """One AdamW step.  With ``use_device_hyper`` the kernels read the table written by ``refresh_device_hyper`` (the
        caller advances ``step_count``) -- the form that can be captured in a CUDA graph."""
        lib = _lib.load()
        if not use_device_hyper:
            self.step_count += 1
        t = max(self.step_count, 1)
        a = self.arena
        for gi, g in enumerate(self.param_groups):
            begin, end = g["_range"]
            if end <= begin:
                continue
            b1, b2 = g["betas"]
            hyper = (ctypes.c_float * 7)(g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** t, 1.0 - b2 ** t)
            sh = 0 if ema_shadow is None else ema_shadow.data_ptr() + 4 * begin
            hdev = self._hyper_dev.data_ptr() + 32 * gi if use_device_hyper else None
            _lib.check(lib.sdt_adamw_flat(a.params.data_ptr() + 4 * begin, a.grads.data_ptr() + 4 * begin,
                                          self.exp_avg.data_ptr() + 4 * begin, self.exp_avg_sq.data_ptr() + 4 * begin,
                                          end - begin, hyper, hdev, grad_scale, sh, ema_one_minus_decay,
                                          _lib.ptr(ema_one_minus_decay_dev), _lib.stream_ptr()), "sdt_adamw_flat")
        # the masters changed in place: LoraArena.pack() (explicit, one launch) refreshes the bf16 operands

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "_range"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
