"""DDPM noising, prediction target and MSE loss -- the arithmetic under the reference's
``LatentDiffusionModel._denoise_loss`` / ``training_step`` (``modules/model.py:289-348``).

``NoiseScheduler`` stands in for the part of ``diffusers.DDIMScheduler`` the training path touches
(``scheduler.add_noise``, ``scheduler.get_velocity``, ``scheduler.config.num_train_timesteps``,
``scheduler.config.prediction_type``).  Both methods are ONE launch of ``sdt_noise_target``; the loss, its two-segment
(prior-preservation) mean, the gradient w.r.t. the prediction and the NaN guard are ONE launch of ``sdt_mse_loss``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import _lib

_MODES = {"epsilon": _lib.TARGET_EPSILON, "sample": _lib.TARGET_SAMPLE, "v": _lib.TARGET_V}


def scaled_linear_alphas_cumprod(beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000) -> torch.Tensor:
    """``scaled_linear`` schedule (``modules/convert/sd_to_diffusers.py:236-243``; SD1.x constants as in
    ``lab/diffusers_sampler_experiment.py:60-65``): betas = linspace(sqrt b0, sqrt b1, T)^2, abar = cumprod(1-betas)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class NoiseScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, prediction_type="epsilon",
                 alphas_cumprod: Optional[torch.Tensor] = None):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule="scaled_linear", prediction_type=prediction_type)
        self.alphas_cumprod = (alphas_cumprod.float().clone() if alphas_cumprod is not None
                               else scaled_linear_alphas_cumprod(beta_start, beta_end, num_train_timesteps))
        self._dev_cache = {}

    def _abar(self, device) -> torch.Tensor:
        t = self._dev_cache.get(device)
        if t is None:
            t = self.alphas_cumprod.to(device)
            self._dev_cache[device] = t
        return t

    def noise_and_target(self, latents: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor,
                         prediction_type: Optional[str] = None, check_range: bool = False):
        """(noisy_latents, target) in one launch.  For ``epsilon`` / ``sample`` the target aliases ``noise`` /
        ``latents`` exactly as ``modules/model.py:306-310`` does."""
        ptype = prediction_type or self.config.prediction_type
        if ptype not in _MODES:
            raise Exception("Unknown prediction type")          # modules/model.py:313-314
        _lib.require_cuda(latents, noise, timesteps)
        _lib.device_check()
        if latents.shape != noise.shape or latents.dtype != noise.dtype:
            raise _lib.SdtError("latents and noise must have the same shape and dtype")
        if timesteps.dtype != torch.int64 or timesteps.numel() != latents.shape[0]:
            raise _lib.SdtError("timesteps must be int64 of shape [batch]")
        x0, eps, t = latents.contiguous(), noise.contiguous(), timesteps.contiguous()
        noisy = torch.empty_like(x0)
        mode = _MODES[ptype]
        target = torch.empty_like(x0) if mode == _lib.TARGET_V else None
        flag = torch.zeros(1, dtype=torch.int32, device=x0.device) if check_range else None
        B = x0.shape[0]
        _lib.check(_lib.load().sdt_noise_target(x0.data_ptr(), eps.data_ptr(), t.data_ptr(), self._abar(x0.device).data_ptr(),
                                                self.config.num_train_timesteps, noisy.data_ptr(), _lib.ptr(target), mode, B,
                                                x0.numel() // B, _lib.dtype_code(x0.dtype), _lib.ptr(flag),
                                                _lib.stream_ptr()), "sdt_noise_target")
        if check_range and int(flag.item()) != 0:
            raise IndexError("timestep out of range")
        if mode == _lib.TARGET_EPSILON:
            target = noise
        elif mode == _lib.TARGET_SAMPLE:
            target = latents
        return noisy, target

    # diffusers-compatible surface used by modules/model.py:302,312
    def add_noise(self, original_samples, noise, timesteps):
        return self.noise_and_target(original_samples, noise, timesteps, "epsilon")[0]

    def get_velocity(self, sample, noise, timesteps):
        return self.noise_and_target(sample, noise, timesteps, "v")[1]


class _MSELoss(torch.autograd.Function):
    """loss = mean((pred-target)^2) or mean(first half) + w * mean(second half); dPred comes from the same launch."""

    @staticmethod
    def forward(ctx, pred, target, split, w_prior, workspace, nan_flag, want_elem):
        lib = _lib.load()
        B = pred.shape[0]
        chw = pred.numel() // B
        p, t = pred.contiguous(), target.contiguous()
        out = torch.empty(3, dtype=torch.float32, device=pred.device)
        dpred = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        elem = torch.empty(pred.shape, dtype=torch.float32, device=pred.device) if want_elem else None
        _lib.check(lib.sdt_mse_loss(p.data_ptr(), _lib.dtype_code(p.dtype), t.data_ptr(), _lib.dtype_code(t.dtype),
                                    out.data_ptr(), _lib.ptr(dpred), _lib.ptr(elem), _lib.ptr(nan_flag), B, chw, split,
                                    w_prior, 1.0, workspace.data_ptr(), _lib.stream_ptr()), "sdt_mse_loss")
        ctx.save_for_backward(dpred)
        ctx.mark_non_differentiable(out)
        loss = out[0].clone()
        if want_elem:
            ctx.mark_non_differentiable(elem)
            return loss, out, elem
        return loss, out, None

    @staticmethod
    def backward(ctx, g_loss, _g_out, _g_elem):
        (dpred,) = ctx.saved_tensors
        if dpred is None:
            return None, None, None, None, None, None, None
        return dpred * g_loss.to(dpred.dtype), None, None, None, None, None, None


class DenoiseLoss:
    """Holds the reduction workspace + NaN flag and mirrors the loss part of ``training_step``
    (``modules/model.py:316,336-342``)."""

    def __init__(self, device, prior_preservation: bool = False, prior_loss_weight: float = 1.0):
        lib = _lib.load()
        self.prior_preservation = prior_preservation
        self.prior_loss_weight = float(prior_loss_weight)
        self.workspace = torch.zeros(lib.sdt_mse_loss_workspace_bytes(), dtype=torch.uint8, device=device)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.last_parts: Optional[torch.Tensor] = None      # [loss, mean(first), mean(second)]

    def __call__(self, pred: torch.Tensor, target: torch.Tensor, want_elementwise: bool = False):
        _lib.require_cuda(pred, target)
        _lib.device_check()
        if pred.shape != target.shape:
            raise _lib.SdtError("pred and target must have the same shape")
        B = pred.shape[0]
        if self.prior_preservation:
            if B % 2 != 0:
                raise _lib.SdtError("prior preservation needs an even batch (instance | class halves)")
            split, w = B // 2, self.prior_loss_weight
        else:
            split, w = B, 0.0
        loss, parts, elem = _MSELoss.apply(pred, target.detach(), split, w, self.workspace, self.nan_flag, want_elementwise)
        self.last_parts = parts
        return (loss, elem) if want_elementwise else loss

    def raise_if_nan(self, name="loss") -> None:
        """Deferred form of ``raise_if_nan`` (``modules/utils/torch/__init__.py:4-8``): one sync, when asked."""
        if int(self.nan_flag.item()) != 0:
            self.nan_flag.zero_()
            raise Exception(f"NaN element discovered in {name}")
