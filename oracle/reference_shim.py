"""Oracle (test infrastructure): execute the reference's OWN source for the
three pieces of the path that are importable without its third-party stack.

Only usable where ``/root/reference`` exists (the build container).  Nothing is
copied: the source text is read from ``/root/reference`` at run time and
executed with the two un-importable imports replaced by stand-ins:

* ``modules/ema.py``                     -- imported as is (torch only).
* ``modules/dataset/bucket.py``          -- ``from . import Size`` and the
  rank-zero logger import (needs ``lightning_utilities``) are replaced.
* ``modules/utils/torch/module.py``      -- ``omegaconf`` is replaced by a
  stand-in whose ``OmegaConf.merge`` is a recursive dict merge (later wins,
  lists replaced), ``DictConfig = dict``, ``ListConfig = list``.
* ``modules/dataset/samplers.py``        -- only ``scale_bucket_params``,
  ``get_gen_bucket_params`` and the two Aspect samplers are extracted by AST
  (the module imports PIL datasets we do not need).

Used by ``oracle/make_golden.py`` to write ``tests/golden/`` and by
``tests/test_reference_live.py`` (skipped when ``/root/reference`` is absent).
"""
from __future__ import annotations

import ast
import importlib.util
import logging
import sys
import types
from dataclasses import dataclass
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "modules" / "ema.py").is_file()


def deep_merge(*configs):
    """Stand-in for ``OmegaConf.merge`` on plain dicts."""
    out = {}
    for cfg in configs:
        for k, v in dict(cfg).items():
            if isinstance(v, dict) and isinstance(out.get(k), dict):
                out[k] = deep_merge(out[k], v)
            elif isinstance(v, dict):
                out[k] = deep_merge(v)
            else:
                out[k] = v
    return out


def _fake_omegaconf():
    mod = types.ModuleType("omegaconf")

    class OmegaConf:
        merge = staticmethod(deep_merge)

    mod.OmegaConf, mod.DictConfig, mod.ListConfig = OmegaConf, dict, list
    return mod


def load_reference_ema():
    spec = importlib.util.spec_from_file_location("_ref_ema", REFERENCE_ROOT / "modules" / "ema.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_bucket():
    src = (REFERENCE_ROOT / "modules" / "dataset" / "bucket.py").read_text()
    src = src.replace("from . import Size\n", "Size = tuple\n")
    src = src.replace("from ..utils.logging import rank_zero_logger\n",
                      "import logging as _l\n\n\ndef rank_zero_logger(name):\n    return _Quiet()\n")
    mod = types.ModuleType("_ref_bucket")

    class _Quiet:  # the reference logs with str.format-style args; swallow them
        def debug(self, *a, **k):
            pass

    mod.__dict__["_Quiet"] = _Quiet
    sys.modules["_ref_bucket"] = mod  # dataclasses needs the module registered
    exec(compile(src, "reference:modules/dataset/bucket.py", "exec"), mod.__dict__)
    return mod


def load_reference_module_walker():
    saved = sys.modules.get("omegaconf")
    sys.modules["omegaconf"] = _fake_omegaconf()
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_module", REFERENCE_ROOT / "modules" / "utils" / "torch" / "module.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["omegaconf"]
        else:
            sys.modules["omegaconf"] = saved
    return mod


@dataclass
class Index:  # stand-in for modules/dataset/datasets.py:45-48
    value: int
    size: tuple


class _AttrDict(dict):
    """dict with attribute access, standing in for DictConfig in samplers.py."""
    __getattr__ = dict.get


def load_reference_samplers():
    import copy
    import random

    class Sampler:  # torch >= 2.x Sampler.__init__ no longer accepts data_source; the reference passes it
        def __init__(self, *args, **kwargs):
            pass

    bucket = load_reference_bucket()
    tree = ast.parse((REFERENCE_ROOT / "modules" / "dataset" / "samplers.py").read_text())
    wanted = {"scale_bucket_params", "get_gen_bucket_params", "AspectSampler", "AspectSamplerDB"}
    body = [n for n in tree.body if getattr(n, "name", None) in wanted]
    ns = {
        "copy": copy, "random": random, "Sampler": Sampler, "Size": tuple, "Index": Index,
        "BucketManager": bucket.BucketManager, "DictConfig": dict, "AspectDataset": object, "DBDataset": object,
        "OmegaConf": _fake_omegaconf().OmegaConf,
    }
    exec(compile(ast.Module(body=body, type_ignores=[]), "reference:modules/dataset/samplers.py", "exec"), ns)
    mod = types.SimpleNamespace(**{k: ns[k] for k in wanted})
    mod.bucket = bucket
    mod.AttrDict = _AttrDict
    return mod


logging.getLogger("arb").addHandler(logging.NullHandler())


# ----------------------------------------------------------------------------------------------------------------
# modules/lora.py and the hot-path methods of modules/model.py: the reference's OWN glue, executed with stand-ins
# for the third-party packages that are absent here (loralib, diffusers, pytorch_lightning, omegaconf).
# What these pins cover: the isinstance switch / attribute aliasing / buffer juggling of ``get_lora``; the target
# switch (``"v"`` spelling), the order of the random draws, the NaN guards, ``chunk`` / ``mean`` prior-preservation
# reduction of ``_denoise_loss`` / ``training_step``; ``config_module``'s param groups; ``get_optimizer``'s LR scaling;
# ``on_save_checkpoint``'s key filter.  What stays [ext]-unpinned: the arithmetic INSIDE loralib's forward and inside
# DDIMScheduler.add_noise / get_velocity (restated from the published algorithms in oracle/lora_ref.py, diffusion_ref.py).
# ----------------------------------------------------------------------------------------------------------------

class AttrDict(dict):
    """Nested attribute access over plain dicts (stand-in for DictConfig on the read-only paths the methods use)."""

    def __getattr__(self, key):
        try:
            v = self[key]
        except KeyError as e:
            raise AttributeError(key) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v


def _fake_loralib():
    """``loralib`` stand-in: the restated loralib-0.1 layers (oracle/lora_ref.py) with the positional constructor signature
    ``modules/lora.py:14,16`` uses and the ``lora_alpha`` attribute loralib's ``LoRALayer`` sets (``get_lora`` deletes it)."""
    from oracle import lora_ref
    mod = types.ModuleType("loralib")

    class Linear(lora_ref.RefLoRALinear):
        def __init__(self, in_features, out_features, r=0, lora_alpha=1, lora_dropout=0.0):
            super().__init__(in_features, out_features, r, lora_alpha, lora_dropout)
            self.lora_alpha = lora_alpha

    class Conv2d(lora_ref.RefLoRAConv2d):
        def __init__(self, in_channels, out_channels, kernel_size, r=0, lora_alpha=1, lora_dropout=0.0):
            super().__init__(in_channels, out_channels, kernel_size, r, lora_alpha, lora_dropout)
            self.lora_alpha = lora_alpha

    mod.Linear, mod.Conv2d = Linear, Conv2d
    return mod


def load_reference_lora():
    """Execute ``/root/reference/modules/lora.py`` as is, ``import loralib`` resolving to the stand-in."""
    saved = sys.modules.get("loralib")
    sys.modules["loralib"] = _fake_loralib()
    try:
        spec = importlib.util.spec_from_file_location("_ref_lora", REFERENCE_ROOT / "modules" / "lora.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["loralib"]
        else:
            sys.modules["loralib"] = saved
    return mod


def load_reference_model():
    """The hot-path pieces of ``/root/reference/modules/model.py`` -- ``get_optimizer``, ``config_module`` and the class
    ``LatentDiffusionModel`` -- compiled from the reference's own source text (AST extraction: the module's import block pulls
    in diffusers / Lightning / PIL datasets).  Names they use resolve to the reference's own importable helpers
    (``get_lora`` via ``load_reference_lora``, ``apply_module_config`` / ``set_submodule`` / ``freeze_permanently``,
    ``raise_if_nan``, ``get_class``, ``ExponentialMovingAverage``) and ``pl.LightningModule`` to ``torch.nn.Module``."""
    import __future__
    import math
    import warnings
    from typing import Any, Mapping, Optional, Sequence

    import torch
    import torch.nn.functional as F
    from torch import nn

    def _load(rel, name):
        spec = importlib.util.spec_from_file_location(name, REFERENCE_ROOT / rel)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    walker = load_reference_module_walker()
    tree = ast.parse((REFERENCE_ROOT / "modules" / "model.py").read_text())
    wanted = {"get_optimizer", "config_module", "LatentDiffusionModel"}
    body = [n for n in tree.body if getattr(n, "name", None) in wanted]
    pl = types.SimpleNamespace(LightningModule=nn.Module, Trainer=object)
    ns = {
        "math": math, "warnings": warnings, "torch": torch, "F": F, "nn": nn, "pl": pl,
        "Any": Any, "Optional": Optional, "Sequence": Sequence, "Mapping": Mapping,
        "DictConfig": dict, "ListConfig": list, "OmegaConf": _fake_omegaconf().OmegaConf,
        "get_lora": load_reference_lora().get_lora,
        "set_submodule": walker.set_submodule, "apply_module_config": walker.apply_module_config,
        "freeze_permanently": walker.freeze_permanently,
        "raise_if_nan": _load("modules/utils/torch/__init__.py", "_ref_torch_utils").raise_if_nan,
        "get_class": _load("modules/utils/activator.py", "_ref_activator").get_class,
        "ExponentialMovingAverage": load_reference_ema().ExponentialMovingAverage,
    }
    code = compile(ast.Module(body=body, type_ignores=[]), "reference:modules/model.py", "exec",
                   flags=__future__.annotations.compiler_flag)
    exec(code, ns)
    return types.SimpleNamespace(get_optimizer=ns["get_optimizer"], config_module=ns["config_module"],
                                 LatentDiffusionModel=ns["LatentDiffusionModel"])


def bare_lightning_module(model_ns, **attrs):
    """A ``LatentDiffusionModel`` instance without running its constructor (which builds a diffusers pipeline): the hot-path
    methods only read ``self.unet / scheduler / config`` and call ``self.log_dict`` / ``self.lr_schedulers()``."""
    import torch
    cls = model_ns.LatentDiffusionModel
    obj = cls.__new__(cls)
    torch.nn.Module.__init__(obj)
    obj.logged = []
    obj.log_dict = lambda d: obj.logged.append(dict(d))
    obj.lr_schedulers = lambda: types.SimpleNamespace(get_lr=lambda: [0.0])
    for k, v in attrs.items():
        setattr(obj, k, v)
    return obj
