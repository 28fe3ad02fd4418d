"""Oracle (test infrastructure): the reference's EMA, ``modules/ema.py:9-140``.

``RefEMA`` restates the reference class.  One deliberate difference, marked
below: ``update`` / ``apply`` / ``average_parameters`` walk the *shadow keys*
instead of every ``named_parameters()`` entry.  The reference stores shadows
only for ``requires_grad`` parameters (``ema.py:33-37``) but walks all
parameters (``ema.py:56-57``), so it raises ``KeyError`` on any partially
frozen module -- i.e. always under LoRA.  For all-trainable
modules both walks are identical, and there this class is pinned against the
reference's own ``modules/ema.py`` (``tests/golden/ema_*.pt``).
"""
from __future__ import annotations

import contextlib
import copy

import torch
from torch import nn


class RefEMA:
    def __init__(self, module: nn.Module, decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.module = module
        self.shadow_params = {name: p.clone().detach()
                              for name, p in module.named_parameters() if p.requires_grad}

    def _tracked(self):
        # DEVIATION from ema.py:56 (see module docstring): restrict to shadow keys.
        for name, p in self.module.named_parameters():
            if name in self.shadow_params:
                yield name, p

    @staticmethod
    def decay_at(decay, num_updates):
        """``ema.py:47-54``: warm-up ``min(decay, (1+n)/(10+n))`` after ``n += 1``."""
        return min(decay, (1 + num_updates) / (10 + num_updates))

    @torch.no_grad()
    def update(self):
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = self.decay_at(decay, self.num_updates)
        one_minus_decay = 1.0 - decay
        for name, p in self._tracked():
            s = self.shadow_params[name]
            tmp = s - p
            tmp.mul_(one_minus_decay)
            s.sub_(tmp)

    def apply(self):
        for name, p in self._tracked():
            p.data.copy_(self.shadow_params[name].data)

    @contextlib.contextmanager
    def average_parameters(self):
        saved = {name: p.clone() for name, p in self._tracked()}
        self.apply()
        try:
            yield
        finally:
            for name, p in self._tracked():
                p.data.copy_(saved[name].data)

    def to(self, device=None, dtype=None):
        self.shadow_params = {
            n: (p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device))
            for n, p in self.shadow_params.items()}

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params}

    def load_state_dict(self, state_dict):
        state_dict = copy.deepcopy(state_dict)
        decay = state_dict["decay"]
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        num_updates = state_dict["num_updates"]
        assert num_updates is None or isinstance(num_updates, int), "Invalid num_updates"
        shadow = state_dict["shadow_params"]
        assert isinstance(shadow, dict), "shadow_params must be a dict"
        assert all(isinstance(p, torch.Tensor) for p in shadow.values()), "shadow_params must all be Tensors"
        self.decay, self.num_updates, self.shadow_params = decay, num_updates, shadow
