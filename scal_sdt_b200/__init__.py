"""scal_sdt_b200 -- B200 (sm_100a) implementation of SCAL-SDT's LoRA training-step hot path.

Scope: LoRA-injected projections (forward + dX/dA/dB), DDPM noising / prediction target / MSE
loss, EMA update, flat-arena AdamW and the data-parallel gradient exchange -- behind the reference's own
module-injection / loss / EMA interfaces.  All arithmetic runs in ``libsdt_b200.so`` (hand-written CUDA, C ABI in
``include/sdt_b200.h``); importing the package never builds or loads it, calling an op without it raises.
"""
from ._lib import SdtError, library_path  # noqa: F401
from .lora import LoRAConv2d, LoRALinear, get_lora, get_linears, lora_modules  # noqa: F401
from .module_config import (Selection, apply_module_config, config_module, freeze_permanently,  # noqa: F401
                            merge_config, plan_module_config, set_submodule)
from .ema import ExponentialMovingAverage  # noqa: F401
from .diffusion import DenoiseLoss, NoiseScheduler, scaled_linear_alphas_cumprod  # noqa: F401
from .arena import LoraArena, ParamArena  # noqa: F401
from .optim import FlatAdamW  # noqa: F401
from .comm import GradExchange  # noqa: F401

__version__ = "0.1.0"
