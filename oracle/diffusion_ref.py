"""Oracle (test infrastructure): DDPM noising, prediction target and MSE loss.

Restates ``modules/model.py:289-316`` (``_denoise_loss``) and ``:318-348``
(``training_step``).  ``scheduler.add_noise`` / ``scheduler.get_velocity``
(called at ``model.py:302,312``) live in an unpinned diffusers fork
(``requirements.txt:13``) that is absent here; the published
``DDIMScheduler`` algorithm is restated:

* ``betas = linspace(sqrt(b0), sqrt(b1), T, float32) ** 2`` ("scaled_linear",
  schedule type per ``modules/convert/sd_to_diffusers.py:236-243``; SD1.x
  constants b0=0.00085, b1=0.012, T=1000 per
  ``lab/diffusers_sampler_experiment.py:60-65``);
  ``alphas_cumprod = cumprod(1 - betas)``.
* ``add_noise``: ``a = alphas_cumprod.to(sample.dtype)[t] ** 0.5``,
  ``b = (1 - alphas_cumprod.to(sample.dtype)[t]) ** 0.5``, broadcast over
  C,H,W; ``noisy = a * x0 + b * eps``.
* ``get_velocity``: ``v = a * eps - b * x0``.

PARITY UNPINNED by the reference (no tests, package absent); cross-checked by
analytic properties in ``tests/test_oracle.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SD1_BETA_START, SD1_BETA_END, SD1_TRAIN_STEPS = 0.00085, 0.012, 1000


def ref_alphas_cumprod(beta_start=SD1_BETA_START, beta_end=SD1_BETA_END, num_train_timesteps=SD1_TRAIN_STEPS):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def _coeffs(alphas_cumprod, like, timesteps):
    ac = alphas_cumprod.to(device=like.device, dtype=like.dtype)
    a = ac[timesteps] ** 0.5
    b = (1 - ac[timesteps]) ** 0.5
    a, b = a.flatten(), b.flatten()
    while a.dim() < like.dim():
        a, b = a.unsqueeze(-1), b.unsqueeze(-1)
    return a, b


def ref_add_noise(alphas_cumprod, x0, eps, timesteps):
    a, b = _coeffs(alphas_cumprod, x0, timesteps)
    return a * x0 + b * eps


def ref_get_velocity(alphas_cumprod, x0, eps, timesteps):
    a, b = _coeffs(alphas_cumprod, x0, timesteps)
    return a * eps - b * x0


def ref_target(prediction_type, alphas_cumprod, x0, eps, timesteps):
    """``modules/model.py:306-314`` -- note the reference spells it ``"v"``."""
    if prediction_type == "epsilon":
        return eps
    if prediction_type == "sample":
        return x0
    if prediction_type == "v":
        return ref_get_velocity(alphas_cumprod, x0, eps, timesteps)
    raise Exception("Unknown prediction type")


def ref_elementwise_loss(pred, target):
    """``modules/model.py:316``; autocast promotes mse_loss to fp32."""
    return F.mse_loss(pred.float(), target.float(), reduction="none")


def ref_reduce_loss(loss, prior_preservation=False, prior_loss_weight=1.0):
    """``modules/model.py:338-342``; halves are instance then class
    (``modules/dataset/__init__.py:77-86``)."""
    if prior_preservation:
        loss, prior_loss = torch.chunk(loss, 2, dim=0)
        return loss.mean() + prior_loss_weight * prior_loss.mean()
    return loss.mean()


def ref_denoise_loss(unet, alphas_cumprod, prediction_type, latents, conds, noise, timesteps):
    """``_denoise_loss`` with the random draws (``model.py:294,297-298``) passed in."""
    noisy = ref_add_noise(alphas_cumprod, latents, noise, timesteps)
    pred = unet(noisy, timesteps, conds)
    target = ref_target(prediction_type, alphas_cumprod, latents, noise, timesteps)
    return ref_elementwise_loss(pred, target)


def ref_training_step(unet, alphas_cumprod, prediction_type, batch, noise, timesteps,
                      prior_preservation=False, prior_loss_weight=1.0):
    """``training_step`` for cached batches (keys ``latents``/``conds``), NaN guards included."""
    latents, conds = batch["latents"], batch["conds"]
    for x, name in ((latents, "VAE output"), (conds, "text encoder output")):
        if torch.any(torch.isnan(x)):
            raise Exception(f"NaN element discovered in {name}")
    loss = ref_denoise_loss(unet, alphas_cumprod, prediction_type, latents, conds, noise, timesteps)
    if torch.any(torch.isnan(loss)):
        raise Exception("NaN element discovered in loss")
    return ref_reduce_loss(loss, prior_preservation, prior_loss_weight)


class RefDDIMScheduler:
    """The slice of diffusers' ``DDIMScheduler`` the training path touches (``modules/model.py:297,302,306,312``):
    ``config.num_train_timesteps``, ``config.prediction_type``, ``add_noise``, ``get_velocity`` -- restated, [ext]-unpinned."""

    def __init__(self, prediction_type="epsilon", beta_start=SD1_BETA_START, beta_end=SD1_BETA_END,
                 num_train_timesteps=SD1_TRAIN_STEPS):
        from types import SimpleNamespace
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                                      beta_start=beta_start, beta_end=beta_end, beta_schedule="scaled_linear")
        self.alphas_cumprod = ref_alphas_cumprod(beta_start, beta_end, num_train_timesteps)

    def add_noise(self, original_samples, noise, timesteps):
        return ref_add_noise(self.alphas_cumprod, original_samples, noise, timesteps)

    def get_velocity(self, sample, noise, timesteps):
        return ref_get_velocity(self.alphas_cumprod, sample, noise, timesteps)
