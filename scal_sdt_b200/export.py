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


# ---- SVD extraction (``extract_lora.py``) --------------------------------------------------------------------------
def lora_approx(delta_w: torch.Tensor, rank: int) -> tuple[torch.Tensor, torch.Tensor]:
    """``extract_lora.py:23-39``: rank-``rank`` factors ``(lora_down [r,in], lora_up [out,r])`` of a weight difference,
    ``up @ down`` being its best rank-r approximation.  Runs where ``delta_w`` lives (cuSOLVER on a GPU tensor -- the
    reference quotes ~15x over the CPU for this offline tool, ``extract_lora.py:79``)."""
    u, s, v_t = torch.linalg.svd(delta_w.float(), full_matrices=False)
    return v_t[:rank, :], u[:, :rank] * s[:rank]


def extract_lora_state_dict(tuned: torch.nn.Module, base: torch.nn.Module, targets: list, prefix: str = "lora_unet",
                            dtype: torch.dtype = torch.float16, device: Optional[torch.device] = None) -> dict:
    """``extract_lora.py:104-154`` for one module tree: walk ``targets`` (an optim_target ``targets`` list) over the tuned and
    the base model, SVD-approximate ``W_tuned - W_base`` of every site that carries a ``lora`` config, and emit kohya keys
    ``<prefix>_<path '.'->'_'>.lora_down.weight / .lora_up.weight / .alpha``; both factors are scaled by ``sqrt(rank / alpha)``
    so that ``(alpha / rank) * up @ down`` reproduces the difference.  1x1-conv weights are emitted 2-D."""
    import math

    from .module_config import apply_module_config
    sites: dict[str, list] = {}
    apply_module_config(tuned, targets, lambda m, c, p: sites.__setitem__(p, [c, m]))
    apply_module_config(base, targets, lambda m, c, p: sites[p].append(m))
    state = {}
    for path, (cfg, mod, mod_base) in sites.items():
        lora_cfg = cfg.get("lora") if isinstance(cfg, Mapping) else None
        if lora_cfg is None:
            continue
        rank, alpha = lora_cfg.get("rank", 4), lora_cfg.get("alpha", 1)
        if isinstance(mod, torch.nn.Linear):
            delta = mod.weight.detach() - mod_base.weight.detach()
        elif isinstance(mod, torch.nn.Conv2d) and tuple(mod.kernel_size) == (1, 1):
            delta = (mod.weight.detach() - mod_base.weight.detach()).squeeze()
        else:
            raise Exception("Only Linear and Conv2d(kernel_size=(1,1)) supports LoRA.")
        if device is not None:
            delta = delta.to(device)
        down, up = lora_approx(delta, rank)
        scale = math.sqrt(rank / alpha)              # X @ c Vt @ c U == c^2 X @ Vt @ U
        key = f"{prefix}_{path.replace('.', '_')}"
        state[f"{key}.lora_down.weight"] = (down * scale).to(dtype).cpu()
        state[f"{key}.lora_up.weight"] = (up * scale).to(dtype).cpu()
        state[f"{key}.alpha"] = torch.tensor(alpha, dtype=torch.int32)
    return state
