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
