"""Target-selection DSL of the reference (``configs/optim_targets/*.yaml``) over plain dicts / lists (PyYAML).

The reference walks the nested ``{index, targets, recurse_conf}`` list and calls a callback while it walks
(``modules/utils/torch/module.py:30-63``), and ``config_module`` (``modules/model.py:136-164``) injects LoRA from inside that
callback.  Here the two concerns are separate stages:

1. ``plan_module_config`` expands the nested list against the module tree into a flat, ordered work list of
   ``Selection(path, config)`` -- which module, and the entry's own keys merged with every ``recurse_conf`` in force;
2. ``config_module`` / ``apply_module_config`` run over that list.

The work list is what ``tests/golden/walker.json`` pins: it was recorded from the reference's own walker on all five stock
optim_target files, visit for visit (order, dotted path, merged config).  Quirks of the reference that the plan keeps
because they change which config a module receives: a ``recurse_conf`` keeps applying to the *later siblings* of the
entry that introduced it (it is folded into a running value while the list is scanned), and an entry without ``index``
selects every direct child.

``merge_config`` stands in for ``OmegaConf.merge``: recursive dict merge, later argument wins, lists are replaced.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from types import MethodType
from typing import Any, Callable, Iterator, Optional

from torch import nn

from .lora import get_lora


def merge_config(*configs) -> dict:
    merged: dict = {}
    for layer in configs:
        if layer is None:
            continue
        for key, value in dict(layer).items():
            if isinstance(value, dict):
                below = merged.get(key)
                merged[key] = merge_config(below if isinstance(below, dict) else None, value)
            else:
                merged[key] = value
    return merged


@dataclass(frozen=True)
class Selection:
    """One visit of the walker: ``path`` is dotted from the walk's root (prefixed by the caller's ``path`` argument),
    ``relative`` is the same without that prefix (what ``root.get_submodule`` resolves), ``config`` the merged entry."""
    path: str
    relative: str
    config: dict


def _join(prefix: str, name: str) -> str:
    return f"{prefix}.{name}" if prefix else name


def _expand(scope: nn.Module, entries: list, here: str, inherited: Optional[dict], descend: bool) -> Iterator[tuple[str, dict]]:
    """Yield ``(path relative to the walk root, merged config)`` for the entries of one nesting level.  ``scope`` is the
    module the entries' ``index`` paths are relative to and ``here`` its own path."""
    running = inherited
    for entry in entries:
        own = entry.get("recurse_conf")
        if own is not None:
            running = own if running is None else merge_config(running, own)
        selected = entry.get("index")
        if selected is None:
            names = [name for name, child in scope.named_children() if child is not scope]
        else:
            names = list(selected)
        nested = entry.get("targets") if descend else None
        for name in names:
            where = _join(here, name)
            if nested is not None:
                yield from _expand(scope.get_submodule(name), nested, where, running, True)
            else:
                yield where, (entry if running is None else merge_config(entry, running))


def plan_module_config(module: nn.Module, module_configs: list, recursive: bool = True, path: str = "",
                       recurse_config: Optional[dict] = None) -> list[Selection]:
    """The ordered work list the reference's walker would visit on ``module`` (nothing is modified)."""
    return [Selection(_join(path, rel), rel, conf) for rel, conf in _expand(module, module_configs, "", recurse_config, recursive)]


def apply_module_config(module: nn.Module, module_configs: list, fn: Callable[[nn.Module, dict, str], None],
                        recursive=True, path="", recurse_config: Optional[dict] = None):
    """Same signature and visit order as ``modules/utils/torch/module.py:30``: ``fn(submodule, merged_config, dotted_path)``
    once per selected module.  Each module is looked up when its turn comes, so a callback that has replaced an earlier
    selection (LoRA injection) is seen by a later selection of the same path, as in the reference."""
    for sel in plan_module_config(module, module_configs, recursive, path, recurse_config):
        fn(module.get_submodule(sel.relative), sel.config, sel.path)


def set_submodule(module: nn.Module, name: str, sub: nn.Module):
    parent_path, _, leaf = name.rpartition(".")
    setattr(module.get_submodule(parent_path), leaf, sub)


def freeze_permanently(module: nn.Module):
    """Frozen for good: no gradients, eval mode, and ``train()`` calls from a parent no longer flip it back."""
    module.requires_grad_(False)
    module.eval()
    module.train = MethodType(lambda self, mode=True: self, module)


def config_module(module: nn.Module, module_configs: list) -> list[dict[str, Any]]:
    """``modules/model.py:136-164``: freeze everything, then for every selected module either inject LoRA (the entry carries
    a ``lora`` dict -> ``get_lora(module, **lora)`` replaces it in the tree and its two factors become trainable) or make all
    of its parameters trainable; one optimizer param group per selection, carrying the entry's ``optimizer`` overrides."""
    module.requires_grad_(False)
    param_groups: list[dict[str, Any]] = []
    for sel in plan_module_config(module, module_configs):
        target = module.get_submodule(sel.relative)
        lora_kwargs = sel.config.get("lora")
        if lora_kwargs is None:
            trainable = list(target.parameters())
        else:
            if not isinstance(target, (nn.Linear, nn.Conv2d)):
                raise AssertionError(f"lora target {sel.path} is a {type(target).__name__}, not nn.Linear / nn.Conv2d")
            injected = get_lora(target, **lora_kwargs)
            set_submodule(module, sel.relative, injected)
            trainable = [injected.lora_A, injected.lora_B]
        for p in trainable:
            p.requires_grad = True
        param_groups.append({"params": trainable, **(sel.config.get("optimizer") or {})})

    selected = sum(len(g["params"]) for g in param_groups)
    if selected != sum(1 for _ in module.parameters()):
        # part of the network stays frozen: autograd's "inputs have requires_grad=False" notice is expected (model.py:158-162)
        warnings.filterwarnings("ignore", message="None of the inputs have requires_grad=True. Gradients will be None")
    return param_groups
