"""Target-selection DSL -- ``modules/utils/torch/module.py:8-69`` and ``config_module``
(``modules/model.py:136-164``) over plain dicts/lists (PyYAML), so that ``configs/optim_targets/*.yaml`` are consumed
unmodified.  ``merge_config`` stands in for ``OmegaConf.merge`` (recursive dict merge, later wins, lists replaced).
"""
from __future__ import annotations

import warnings
from types import MethodType
from typing import Any, Callable, Optional

from torch import nn

from .lora import get_lora


def merge_config(*configs) -> dict:
    out: dict = {}
    for cfg in configs:
        if cfg is None:
            continue
        for k, v in dict(cfg).items():
            if isinstance(v, dict) and isinstance(out.get(k), dict):
                out[k] = merge_config(out[k], v)
            elif isinstance(v, dict):
                out[k] = merge_config(v)
            else:
                out[k] = v
    return out


def set_submodule(module: nn.Module, name: str, sub: nn.Module):
    segments = name.split(".")
    module = module.get_submodule(".".join(segments[:-1]))
    module.__setattr__(segments[-1], sub)


def apply_module_config(module: nn.Module, module_configs: list, fn: Callable[[nn.Module, dict, str], None],
                        recursive=True, path="", recurse_config: Optional[dict] = None):
    """Apply ``fn`` to each submodule selected by the nested ``{index, targets, recurse_conf}`` list.
    Semantics follow ``module.py:30-63`` exactly, including that ``recurse_conf`` keeps accumulating across sibling
    entries of one list (``:35-39``) and that an entry without ``index`` visits every ``named_children()``."""
    for module_config in module_configs:
        index = module_config.get("index")
        targets = module_config.get("targets")

        current_depth = module_config.get("recurse_conf")
        if recurse_config is None:
            recurse_config = current_depth
        elif current_depth is not None:
            recurse_config = merge_config(recurse_config, current_depth)

        def invoke_on_submodule(_submodule: nn.Module, _module_path: str):
            _path = _module_path if path == "" else f"{path}.{_module_path}"
            if recursive and targets is not None:
                apply_module_config(_submodule, targets, fn, path=_path, recurse_config=recurse_config)
            else:
                config = module_config if recurse_config is None else merge_config(module_config, recurse_config)
                fn(_submodule, config, _path)

        if index is None:
            for name, submodule in module.named_children():
                if submodule == module:
                    continue
                invoke_on_submodule(submodule, name)
        else:
            for module_path in index:
                submodule = module.get_submodule(module_path)
                invoke_on_submodule(submodule, module_path)


def freeze_permanently(module: nn.Module):
    module.requires_grad_(False)
    module.eval()
    module.train = MethodType(lambda self, mode: self, module)


def config_module(module: nn.Module, module_configs: list) -> list[dict[str, Any]]:
    """``modules/model.py:136-164``: freeze everything, inject LoRA where the entry carries a ``lora`` dict, make the
    selected parameters trainable and return one optimizer param group per selected module."""
    module.requires_grad_(False)
    param_groups: list[dict[str, Any]] = []

    def apply_innermost(submodule: nn.Module, submodule_config: dict, module_path: str):
        if (lora_config := submodule_config.get("lora")) is not None:
            assert isinstance(submodule, nn.Linear) or isinstance(submodule, nn.Conv2d)
            submodule = get_lora(submodule, **lora_config)
            set_submodule(module, module_path, submodule)
            params = [submodule.lora_A, submodule.lora_B]
        else:
            params = list(submodule.parameters())

        for param in params:
            param.requires_grad = True

        param_groups.append({"params": params, **(submodule_config.get("optimizer") or {})})

    apply_module_config(module, module_configs, apply_innermost)

    if len(list(module.parameters())) != len([p for g in param_groups for p in g["params"]]):
        warnings.filterwarnings("ignore", message="None of the inputs have requires_grad=True. Gradients will be None")

    return param_groups
