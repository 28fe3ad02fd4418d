"""YAML loading for the reference's config files without OmegaConf.

OmegaConf (YAML 1.2 scalars) reads ``lr: 5e-4`` as a float; PyYAML's YAML 1.1 resolver needs a dot in the mantissa
and would hand back the string ``'5e-4'``.  ``load_yaml`` adds the 1.2 float form, so ``configs/*.yaml`` and
``configs/optim_targets/*.yaml`` (anchors / aliases included) load with the values the reference sees.
"""
from __future__ import annotations

import re
from pathlib import Path

import yaml


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)
                    |\.(?:nan|NaN|NAN))$""", re.X),
    list("-+0123456789."))


def load_yaml_text(text: str):
    return yaml.load(text, Loader=_Loader)


def load_yaml(path):
    return load_yaml_text(Path(path).read_text())
