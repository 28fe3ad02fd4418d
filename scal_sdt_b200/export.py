"""Checkpoint export to the kohya / AddNet LoRA key layout -- ``ckpt_tool.py:156-234`` (``extract_lora``), host-side key
renaming with no arithmetic.

Input: the trainable-only state dict the trainer checkpoints (``unet.<path>.lora_A`` / ``.lora_B``,
``condition_model.encoder.<path>.lora_A/B``; ``modules/model.py:378-391``).  Output keys:
``lora_unet_<path with '.' -> '_'>.lora_down.weight`` (= lora_A), ``.lora_up.weight`` (= lora_B), ``.alpha`` (int32);
text-encoder prefix ``lora_te``.  ``lora_alpha`` is a buffer and is not in the checkpoint, so alpha comes from the run's
optim_target config exactly as the reference recovers it (``ckpt_tool.py:169-179``): the first ``lora`` dict found.
"""
from __future__ import annotations

from typing import Any, Iterator, Mapping, Optional

import torch

_KEY_MAP = {"lora_A": "lora_down.weight", "lora_B": "lora_up.weight", "lora_alpha": "alpha"}


def search_key(config: Any, key: str) -> Iterator[Any]:
    """Depth-first search for ``key`` in nested dicts / lists (``modules/utils/config.py``)."""
    if isinstance(config, Mapping):
        for k, v in config.items():
            if k == key:
                yield v
            else:
                yield from search_key(v, key)
    elif isinstance(config, (list, tuple)):
        for v in config:
            yield from search_key(v, key)


def alpha_from_optim_target(optim_target: Any) -> Optional[int]:
    lora_cfg = next(search_key(optim_target, "lora"), None)
    return None if lora_cfg is None else lora_cfg.get("alpha")


def _strip_prefix(state: Mapping[str, torch.Tensor], prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}


def _to_kohya(state: dict, prefix: str, alpha: Optional[int], dtype: torch.dtype) -> dict:
    modules = sorted({k.rsplit(".", 1)[0] for k in state if k.rsplit(".", 1)[-1] in ("lora_A", "lora_B")})
    out = {}
    for mod in modules:
        alpha_key = f"{mod}.lora_alpha"
        if alpha_key not in state and alpha is not None:
            state[alpha_key] = torch.tensor(alpha, dtype=torch.int32)
        for name, mapped in _KEY_MAP.items():
            v = state.get(f"{mod}.{name}")
            if v is None:
                continue
            if v.dtype.is_floating_point:
                v = v.to(dtype)
            out["_".join([prefix, *mod.split(".")]) + f".{mapped}"] = v
    return out


def to_kohya_state_dict(ssdt_state: Mapping[str, Any], optim_target: Any = None, alpha: Optional[int] = None,
                        dtype: torch.dtype = torch.float16) -> dict:
    """SCAL-SDT trainable-only state dict -> AddNet-compatible LoRA state dict."""
    if alpha is None and optim_target is not None:
        alpha = alpha_from_optim_target(optim_target)
    tensors = {k: v for k, v in ssdt_state.items() if isinstance(v, torch.Tensor)}
    out = {}
    out.update(_to_kohya(_strip_prefix(tensors, "unet."), "lora_unet", alpha, dtype))
    out.update(_to_kohya(_strip_prefix(tensors, "condition_model.encoder."), "lora_te", alpha, dtype))
    return out
