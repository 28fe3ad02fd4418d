"""Oracle (test infrastructure): the reference's whole training step on the CPU, in torch fp32.

This is what the reference executes per step -- ``training_step`` + backward + ``optimizer.step`` +
``on_train_batch_end`` (``modules/model.py:318-348,399-412``) -- with the pieces that live in absent third-party
packages restated: ``oracle.lora_ref`` for loralib, ``oracle.diffusion_ref`` for the DDIM scheduler arithmetic,
``torch.optim.AdamW`` (the optimizer ``configs/lora.yaml:66-73`` names) and ``oracle.ema_ref`` for ``modules/ema.py``.
The host UNet is the torch skeleton shared with the product (``scal_sdt_b200/unet.py``: it contains no hot-path code).
Used by the parity tests (small widths) and timed by ``bench.py`` as the CPU baseline / ``--impl reference`` arm.
"""
from __future__ import annotations

from typing import Any

import torch
from torch import nn

from oracle import diffusion_ref, ema_ref, lora_ref
from scal_sdt_b200.module_config import apply_module_config, set_submodule   # pure-python walker (pinned by walker.json)


def ref_config_module(module: nn.Module, module_configs: list) -> list[dict[str, Any]]:
    """``modules/model.py:136-164`` with the restated loralib modules."""
    module.requires_grad_(False)
    groups: list[dict[str, Any]] = []

    def innermost(sub, conf, path):
        if (lc := conf.get("lora")) is not None:
            sub = lora_ref.ref_get_lora(sub, **lc)
            set_submodule(module, path, sub)
            params = [sub.lora_A, sub.lora_B]
        else:
            params = list(sub.parameters())
        for p in params:
            p.requires_grad = True
        groups.append({"params": params, **(conf.get("optimizer") or {})})

    apply_module_config(module, module_configs, innermost)
    return groups


class RefTrainer:
    def __init__(self, unet: nn.Module, targets: list, prediction_type="epsilon", optimizer_params=None,
                 prior_preservation=False, prior_loss_weight=1.0, ema_decay=None):
        self.unet = unet
        self.groups = ref_config_module(unet, targets)
        op = dict(optimizer_params or {"lr": 5e-4, "betas": (0.9, 0.999), "weight_decay": 2e-2, "eps": 1e-7})
        self.optimizer = torch.optim.AdamW(self.groups, **op)
        self.alphas_cumprod = diffusion_ref.ref_alphas_cumprod()
        self.prediction_type = prediction_type
        self.prior_preservation, self.prior_loss_weight = prior_preservation, prior_loss_weight
        self.ema = ema_ref.RefEMA(unet, ema_decay) if ema_decay is not None else None

    def training_step(self, batch, noise, timesteps):
        return diffusion_ref.ref_training_step(lambda x, t, c: self.unet(x, t, c).sample, self.alphas_cumprod,
                                               self.prediction_type, batch, noise, timesteps, self.prior_preservation,
                                               self.prior_loss_weight)

    def step(self, batch, noise, timesteps):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.training_step(batch, noise, timesteps)
        loss.backward()
        self.optimizer.step()
        if self.ema is not None:
            self.ema.update()
        return loss.detach()
