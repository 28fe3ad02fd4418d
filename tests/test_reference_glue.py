"""The reference's OWN glue around the third-party arithmetic, as recorded by ``oracle/make_golden.py`` from executing
``/root/reference/modules/lora.py`` and ``modules/model.py`` (``get_lora``; ``_denoise_loss`` / ``training_step``;
``config_module``; ``get_optimizer``; ``on_save_checkpoint``) with stand-ins for loralib / diffusers / Lightning / omegaconf
(``oracle/reference_shim.py``).

CPU tier: the oracle restatements (``oracle/lora_ref.py``, ``oracle/diffusion_ref.py``) and the product's host code are
held against those recordings.  GPU tier: the CUDA path (C ABI) against the same recordings.

Residue that stays unpinned by the reference: the arithmetic INSIDE loralib-0.1's forward and inside
``DDIMScheduler.add_noise / get_velocity`` -- both stood in for by the published-algorithm restatements.
"""
import json
from pathlib import Path

import pytest
import torch
from torch import nn

from oracle import diffusion_ref, lora_ref

GOLDEN = Path(__file__).parent / "golden"
SITE_CASES = ("linear_bias_r2_a4", "linear_nobias_r4_a1", "conv1x1_r2_a2")


def _glue():
    return torch.load(GOLDEN / "lora_glue.pt")


def _steps():
    return torch.load(GOLDEN / "denoise_steps.pt")


def _base_for(name, rec):
    base = {"linear_bias_r2_a4": lambda: nn.Linear(6, 5), "linear_nobias_r4_a1": lambda: nn.Linear(8, 3, bias=False),
            "conv1x1_r2_a2": lambda: nn.Conv2d(4, 6, 1)}[name]()
    base.load_state_dict(rec["base_state"])
    return base.requires_grad_(False)


def _contract(lora, base):
    return {"state_keys": sorted(lora.state_dict().keys()),
            "state_dtypes": {k: str(v.dtype) for k, v in lora.state_dict().items()},
            "buffers": sorted(n for n, _ in lora.named_buffers()),
            "requires_grad": {n: bool(p.requires_grad) for n, p in lora.named_parameters()},
            "weight_is_aliased": lora.weight is base.weight, "bias_is_aliased": lora.bias is base.bias,
            "scaling": float(lora.scaling)}


# ---- CPU tier -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SITE_CASES)
def test_get_lora_contract_oracle_and_product_vs_reference_glue(name):
    """``modules/lora.py:12-27`` executed by the reference's own source: keys, dtypes, buffers, aliasing, int32 alpha."""
    from scal_sdt_b200 import get_lora
    rec = _glue()[name]
    for make in (lora_ref.ref_get_lora, get_lora):
        base = _base_for(name, rec)
        lora = make(base, rec["rank"], rec["alpha"])
        got = _contract(lora, base)
        for k, v in got.items():
            assert v == rec[k], (make.__module__, name, k, v, rec[k])
        assert torch.equal(lora.lora_alpha, rec["lora_alpha"]) and lora.lora_alpha.dtype == torch.int32
        assert lora.lora_A.shape == rec["lora_A"].shape and lora.lora_B.shape == rec["lora_B"].shape
        assert "lora_alpha" not in lora.__dict__          # the python attribute is gone, only the buffer remains (lora.py:24-25)
    assert rec["has_python_lora_alpha_attr"] is False


def test_get_lora_init_and_error_vs_reference_glue():
    from scal_sdt_b200 import get_lora
    g = _glue()
    assert g["other_module_error"] == "Unexpected module type"
    for make in (lora_ref.ref_get_lora, get_lora):
        with pytest.raises(Exception, match="Unexpected module type"):
            make(nn.LayerNorm(8))
        torch.manual_seed(0)
        fresh = make(nn.Linear(64, 32), 4, 1)
        assert bool(torch.count_nonzero(fresh.lora_B) == 0) == g["init"]["lora_B_all_zero"]
        assert float(fresh.lora_A.detach().abs().max()) <= g["init"]["bound"] + 1e-7
    assert g["init"]["lora_A_absmax"] <= g["init"]["bound"] + 1e-7


@pytest.mark.parametrize("name", SITE_CASES)
def test_oracle_site_numerics_equal_reference_glue(name):
    """Same restated loralib arithmetic on both sides, reached through different glue: must agree to the bit."""
    rec = _glue()[name]
    lora = lora_ref.ref_get_lora(_base_for(name, rec), rec["rank"], rec["alpha"])
    with torch.no_grad():
        lora.lora_A.copy_(rec["lora_A"]); lora.lora_B.copy_(rec["lora_B"])
    x = rec["x"].clone().requires_grad_(True)
    y = lora(x)
    y.backward(rec["dy"])
    assert torch.equal(y, rec["y"]) and torch.equal(x.grad, rec["dx"])
    assert torch.equal(lora.lora_A.grad, rec["dA"]) and torch.equal(lora.lora_B.grad, rec["dB"])
    assert lora.weight.grad is None and rec["frozen_grads_none"]


def _step_cases():
    return [k for k in _steps() if k != "errors"]


@pytest.mark.parametrize("key", ["epsilon_priorNone", "epsilon_prior0.6", "sample_priorNone", "sample_prior0.6", "v_priorNone",
                                 "v_prior0.6"])
def test_oracle_denoise_step_equals_reference_training_step(key):
    """``oracle/diffusion_ref.py`` against what the reference's own ``_denoise_loss`` / ``training_step`` produced."""
    rec = _steps()[key]
    ac = diffusion_ref.ref_alphas_cumprod()
    fed = {}

    def unet(noisy, t, conds):
        fed["noisy"], fed["t"] = noisy, t
        return rec["pred"]

    loss_elem = diffusion_ref.ref_denoise_loss(unet, ac, rec["prediction_type"], rec["latents"], rec["conds"], rec["noise"],
                                               rec["timesteps"])
    assert torch.equal(fed["noisy"], rec["noisy"]) and torch.equal(fed["t"], rec["timesteps"])
    assert torch.equal(loss_elem, rec["loss_elem"])
    prior = rec["prior_loss_weight"]
    loss = diffusion_ref.ref_training_step(unet, ac, rec["prediction_type"], {"latents": rec["latents"], "conds": rec["conds"]},
                                           rec["noise"], rec["timesteps"], prior is not None, prior or 1.0)
    assert torch.equal(loss, rec["loss"]) and float(loss) == rec["logged"]


def test_oracle_guards_equal_reference_messages():
    errs = _steps()["errors"]
    assert errs == {"unknown_type": "Unknown prediction type", "nan_latents": "NaN element discovered in VAE output",
                    "nan_conds": "NaN element discovered in text encoder output"}
    ac = diffusion_ref.ref_alphas_cumprod()
    z = {"latents": torch.zeros(2, 4, 2, 2), "conds": torch.zeros(2, 3, 12)}
    with pytest.raises(Exception, match=errs["unknown_type"]):
        diffusion_ref.ref_training_step(lambda a, b, c: a, ac, "v_prediction", z, torch.zeros(2, 4, 2, 2), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(Exception, match=errs["nan_latents"]):
        diffusion_ref.ref_training_step(lambda a, b, c: a, ac, "epsilon", dict(z, latents=torch.full((2, 4, 2, 2), float("nan"))),
                                        torch.zeros(2, 4, 2, 2), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(Exception, match=errs["nan_conds"]):
        diffusion_ref.ref_training_step(lambda a, b, c: a, ac, "epsilon", dict(z, conds=torch.full((2, 3, 12), float("nan"))),
                                        torch.zeros(2, 4, 2, 2), torch.zeros(2, dtype=torch.int64))


def test_config_module_param_groups_equal_reference():
    """The reference's own ``config_module`` on the UNet skeleton for every stock optim_target: param groups (names and
    optimizer overrides, in order), trainable set, injected modules, ``on_save_checkpoint`` keys."""
    from scal_sdt_b200 import config_module
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    gold = json.loads((GOLDEN / "config_module.json").read_text())["targets"]
    cfgs = json.loads((GOLDEN / "walker.json").read_text())
    assert set(gold) >= {"lora", "lora_no-te", "full_unet"}
    for stem, rec in gold.items():
        torch.manual_seed(0)
        unet = UNet2DConditionModel(UNetConfig.tiny())
        groups = config_module(unet, cfgs[stem]["config"]["unet"]["targets"])
        names = {id(p): n for n, p in unet.named_parameters()}
        got = [{"params": [names[id(p)] for p in g["params"]], "overrides": {k: v for k, v in g.items() if k != "params"}}
               for g in groups]
        assert got == rec["groups"], stem
        assert sorted(n for n, p in unet.named_parameters() if p.requires_grad) == rec["trainable"], stem
        assert sorted(n for n, m in unet.named_modules() if hasattr(m, "lora_A")) == rec["injected"], stem
        ckpt_keys = sorted(f"unet.{n}" for n, p in unet.named_parameters() if p.requires_grad)
        assert ckpt_keys == rec["checkpoint_keys"], stem        # model.py:378-391 (LightningModule attribute name 'unet')


def test_scale_lr_equals_reference_get_optimizer():
    """``get_optimizer`` (``modules/model.py:33-64``) executed by the reference's own source: lr *= c, weight_decay /= c."""
    from scal_sdt_b200.trainer import scale_lr
    gold = json.loads((GOLDEN / "config_module.json").read_text())["get_optimizer"]
    for method in ("sqrt", "linear"):
        groups = [{"lr": 1e-4, "weight_decay": 1e-2}, {"lr": 5e-4, "weight_decay": 2e-2}]
        scale_lr(groups, {}, gold["batch_size"], gold["devices"], gold["nodes"], gold["accumulate"], method)
        for g, r in zip(groups, gold["result"][method]):
            assert g["lr"] == r["lr"] and g["weight_decay"] == r["weight_decay"], (method, g, r)
        assert gold["result"][method + "_class"] == "AdamW"
        assert gold["result"][method][0]["betas"] == [0.9, 0.999] and gold["result"][method][0]["eps"] == 1e-7


# ---- GPU tier: the CUDA path against the same recordings ---------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("key", ["epsilon_priorNone", "epsilon_prior0.6", "sample_priorNone", "sample_prior0.6", "v_priorNone",
                                 "v_prior0.6"])
def test_cuda_noising_target_loss_equal_reference_training_step(sdt_lib, key):
    from scal_sdt_b200 import DenoiseLoss, NoiseScheduler
    dev = torch.device("cuda:0")
    rec = _steps()[key]
    sched = NoiseScheduler(prediction_type=rec["prediction_type"])
    noisy, target = sched.noise_and_target(rec["latents"].to(dev), rec["noise"].to(dev), rec["timesteps"].to(dev))
    assert torch.equal(noisy.cpu(), rec["noisy"])                         # bit-exact add_noise
    prior = rec["prior_loss_weight"]
    crit = DenoiseLoss(dev, prior is not None, prior or 1.0)
    pred = rec["pred"].to(dev).requires_grad_(True)
    loss, elem = crit(pred, target, want_elementwise=True)
    assert torch.equal(elem.cpu(), rec["loss_elem"])                      # bit-exact target switch + (pred - target)^2
    assert abs(loss.item() - rec["loss"].item()) <= 1e-5 * abs(rec["loss"].item())
    # dLoss/dPred against autograd on the oracle expression
    p2 = rec["pred"].clone().requires_grad_(True)
    diffusion_ref.ref_reduce_loss(diffusion_ref.ref_elementwise_loss(p2, target.cpu()), prior is not None, prior or 1.0).backward()
    loss.backward()
    assert torch.allclose(pred.grad.cpu(), p2.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SITE_CASES)
def test_cuda_site_numerics_vs_reference_glue(sdt_lib, name):
    """fp32 kernels through ``get_lora`` against the numbers the reference's own ``get_lora`` module produced (1e-5)."""
    from scal_sdt_b200 import get_lora
    dev = torch.device("cuda:0")
    rec = _glue()[name]
    base = _base_for(name, rec).to(dev).requires_grad_(False)
    lora = get_lora(base, rec["rank"], rec["alpha"])
    with torch.no_grad():
        lora.lora_A.copy_(rec["lora_A"]); lora.lora_B.copy_(rec["lora_B"])
    x = rec["x"].to(dev).requires_grad_(True)
    y = lora(x)
    y.backward(rec["dy"].to(dev))

    def rel(a, b):
        return ((a.detach().cpu().double() - b.double()).norm() / b.double().norm()).item()
    assert rel(y, rec["y"]) <= 1e-5 and rel(x.grad, rec["dx"]) <= 1e-5
    assert rel(lora.lora_A.grad, rec["dA"]) <= 1e-5 and rel(lora.lora_B.grad, rec["dB"]) <= 1e-5
