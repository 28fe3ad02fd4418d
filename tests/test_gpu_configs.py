"""GPU parity of the BASELINE.json configurations that are not the bench line, at reduced width (same code paths):
cfg3 (rank 64 on attention + FF with EMA), cfg4 (aspect-ratio-bucketed mixed resolution + DreamBooth prior preservation),
cfg5 (SD2.x-shaped UNet: linear proj_in/out, 'v' target, native full fine-tune + EMA)."""
import copy
import random

import numpy as np
import pytest
import torch

from oracle.ref_trainer import RefTrainer

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
_LAST_FROZEN = None


def _pair(cfg, targets, dtype, **kw):
    from scal_sdt_b200 import NoiseScheduler
    from scal_sdt_b200.trainer import LatentDiffusionTrainer
    from scal_sdt_b200.unet import UNet2DConditionModel
    torch.manual_seed(11)
    unet_cpu = UNet2DConditionModel(cfg)
    if dtype == torch.bfloat16:
        with torch.no_grad():
            for p in unet_cpu.parameters():
                p.copy_(p.bfloat16().float())
    global _LAST_FROZEN
    _LAST_FROZEN = copy.deepcopy(unet_cpu)
    unet_gpu = copy.deepcopy(unet_cpu).to(DEV).to(dtype)
    if dtype == torch.bfloat16:
        unet_gpu = unet_gpu.to(memory_format=torch.channels_last)
    ptype = kw.pop("prediction_type", "epsilon")
    ema_decay = kw.pop("ema_decay", None)
    prior = kw.pop("prior", None)
    opt = {"lr": 1e-3, "betas": (0.9, 0.999), "weight_decay": 1e-2, "eps": 1e-7}
    ref = RefTrainer(unet_cpu, copy.deepcopy(targets), prediction_type=ptype, optimizer_params=opt, ema_decay=ema_decay,
                     prior_preservation=prior is not None, prior_loss_weight=prior or 1.0)
    ours = LatentDiffusionTrainer(unet_gpu, NoiseScheduler(prediction_type=ptype), copy.deepcopy(targets),
                                  optimizer_params={"lr": 1e-3, "beta1": 0.9, "beta2": 0.999, "weight_decay": 1e-2, "eps": 1e-7},
                                  ema={"enabled": ema_decay is not None, "decay": ema_decay or 0.0},
                                  prior_preservation={"enabled": prior is not None, "prior_loss_weight": prior or 1.0}, seed=0)
    return ref, ours


def _sync_lora(ref, ours, seed=5):
    g = torch.Generator().manual_seed(seed)
    refm = dict(ref.unet.named_modules())
    with torch.no_grad():
        for name, m in ours.arena.sites:
            a = (torch.randn(m.lora_A.shape, generator=g) * 0.05).bfloat16().float()
            b = (torch.randn(m.lora_B.shape, generator=g) * 0.05).bfloat16().float()
            m.lora_A.copy_(a); m.lora_B.copy_(b)
            refm[name].lora_A.copy_(a); refm[name].lora_B.copy_(b)
    ours.arena.pack()
    return refm


def _torch_autocast_grad_rel(unet_frozen, targets, refm, sites, batch, noise, t, **ref_kw):
    """The reference's own mode on this GPU (fp32 modules under autocast(bf16), eager torch, oracle LoRA layers, none of this
    repo's kernels) against the same fp32 CPU oracle gradients: the yardstick for the whole-network bf16 error."""
    from scal_sdt_b200 import fused
    tg = RefTrainer(copy.deepcopy(unet_frozen).to(DEV), copy.deepcopy(targets), **ref_kw)
    tm = dict(tg.unet.named_modules())
    with torch.no_grad():
        for name, _m in sites:
            tm[name].lora_A.copy_(refm[name].lora_A); tm[name].lora_B.copy_(refm[name].lora_B)
    with fused.torch_only():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = tg.training_step({k: v.to(DEV) for k, v in batch.items()}, noise.to(DEV), t.to(DEV))
        loss.backward()
    num = den = 0.0
    for name, _m in sites:
        for pn in ("lora_A", "lora_B"):
            go, gr = getattr(tm[name], pn).grad.float().cpu(), getattr(refm[name], pn).grad
            num += (go - gr).pow(2).sum().item()
            den += gr.pow(2).sum().item()
    return (num / den) ** 0.5


def _whole_network_bound(e_ours, e_torch):
    """north_star's 2e-2 holds per site (test_gpu_lora / test_gpu_fullsize) and at SD1.5 width for the whole network
    (test_gpu_sd15_step); at toy widths the bf16 host model around the sites dominates, for the reference's own bf16 path too,
    so the bound is: not less accurate than that path."""
    return e_ours <= max(2e-2, 1.25 * e_torch)


def _grad_rel(ours, refm):
    num = den = 0.0
    for name, m in ours.arena.sites:
        for pn in ("lora_A", "lora_B"):
            go, gr = getattr(m, pn).grad.float().cpu(), getattr(refm[name], pn).grad
            num += (go - gr).pow(2).sum().item()
            den += gr.pow(2).sum().item()
    return (num / den) ** 0.5


def test_cfg3_rank64_attention_ff_with_ema(sdt_lib):
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.unet import UNetConfig
    targets = lora_unet_targets(rank=64, alpha=64, projections=False)       # attention + FF: 160 sites
    ref, ours = _pair(UNetConfig.tiny(), targets, torch.bfloat16, ema_decay=0.995)
    assert len(ours.arena.sites) == 160 and all(m.r == 64 for _, m in ours.arena.sites)
    refm = _sync_lora(ref, ours)
    g = torch.Generator().manual_seed(1)
    lat, cond = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 7, 64, generator=g).bfloat16().float()
    noise, t = torch.randn(2, 4, 16, 16, generator=g), torch.tensor([3, 977])
    ours.optimizer.zero_grad()
    lo = ours.training_step({"latents": lat.to(DEV), "conds": cond.to(DEV)}, 0, noise.to(DEV), t.to(DEV))
    lo.backward()
    lr_ = ref.training_step({"latents": lat, "conds": cond}, noise, t)
    lr_.backward()
    assert abs(lo.item() - lr_.item()) <= 2e-2 * abs(lr_.item())
    e_torch = _torch_autocast_grad_rel(_LAST_FROZEN, targets, refm, ours.arena.sites, {"latents": lat, "conds": cond}, noise, t)
    assert _whole_network_bound(_grad_rel(ours, refm), e_torch), (_grad_rel(ours, refm), e_torch)
    # optimizer + EMA step: bias-corrected first AdamW step moves every element by ~lr; compare against torch AdamW + RefEMA
    for name, m in ours.arena.sites:           # feed the oracle's exact gradients so that only the update rule is compared
        m.lora_A.grad.copy_(refm[name].lora_A.grad.to(DEV)); m.lora_B.grad.copy_(refm[name].lora_B.grad.to(DEV))
    ours.optimizer_step()
    ref.optimizer.step(); ref.ema.update()
    for name, m in ours.arena.sites[:8]:
        assert torch.allclose(m.lora_A.cpu(), refm[name].lora_A, rtol=1e-5, atol=1e-7)
        assert torch.allclose(ours.unet_ema.shadow_params[name + ".lora_B"].cpu(), ref.ema.shadow_params[name + ".lora_B"],
                              rtol=1e-5, atol=1e-7)
    assert ours.unet_ema.num_updates == ref.ema.num_updates == 1


def test_cfg4_bucketed_mixed_resolution_prior_preservation(sdt_lib):
    from scal_sdt_b200.bucket import DEFAULT_BUCKET_CONFIG, AspectSamplerDB, collate_order
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.unet import UNetConfig
    sizes = [(512, 512), (768, 512), (512, 768), (1024, 768), (768, 1024)]
    rs = np.random.RandomState(0)
    inst = {i: sizes[int(k)] for i, k in enumerate(rs.randint(0, len(sizes), size=64))}
    cls = {i: sizes[int(k)] for i, k in enumerate(rs.randint(0, len(sizes), size=128))}
    cfg = dict(DEFAULT_BUCKET_CONFIG, manual={"max_size": 786432})
    random.seed(114514)
    sampler = AspectSamplerDB(inst, cls, 512, cfg, 2, 114514, world_size=2, global_rank=1)
    pairs = list(sampler)
    batches = [pairs[i:i + 2] for i in range(0, len(pairs), 2)]
    assert len({b[0][0].size for b in batches}) >= 2                       # several resolutions in one epoch
    ref, ours = _pair(UNetConfig.tiny(), lora_unet_targets(rank=16, alpha=16), torch.bfloat16, prior=0.7)
    refm = _sync_lora(ref, ours)
    g = torch.Generator().manual_seed(2)
    seen = set()
    for batch in batches:
        w, h = batch[0][0].size
        if (w, h) in seen or len(seen) >= 3:
            continue
        seen.add((w, h))
        order = collate_order(batch)                                       # instance items, then class items
        assert all(ix.size == (w, h) for ix in order) and len(order) == 4
        hh, ww = h // 8, w // 8                                            # the real latent geometry of the bucket
        lat = torch.randn(4, 4, hh, ww, generator=g)
        cond = torch.randn(4, 7, 64, generator=g).bfloat16().float()
        noise, t = torch.randn(4, 4, hh, ww, generator=g), torch.randint(0, 1000, (4,), generator=g)
        ours.optimizer.zero_grad()
        lo = ours.training_step({"latents": lat.to(DEV), "conds": cond.to(DEV)}, 0, noise.to(DEV), t.to(DEV))
        lo.backward()
        ref.optimizer.zero_grad(set_to_none=True)
        lr_ = ref.training_step({"latents": lat, "conds": cond}, noise, t)
        lr_.backward()
        assert abs(lo.item() - lr_.item()) <= 2e-2 * abs(lr_.item()), (w, h)
        e_torch = _torch_autocast_grad_rel(_LAST_FROZEN, lora_unet_targets(rank=16, alpha=16), refm, ours.arena.sites,
                                           {"latents": lat, "conds": cond}, noise, t, prior_preservation=True, prior_loss_weight=0.7)
        assert _whole_network_bound(_grad_rel(ours, refm), e_torch), (w, h, _grad_rel(ours, refm), e_torch)
    assert len(seen) >= 2


def test_cfg5_sd2x_shape_full_finetune_v_prediction_ema(sdt_lib):
    """Native fine-tune: gradients come from torch autograd; the path pieces exercised are K3 (v target), K4, the flat
    AdamW and the EMA over the whole parameter arena (fp32, TF32 off: tight tolerances)."""
    from scal_sdt_b200.targets import full_unet_targets
    from scal_sdt_b200.unet import UNetConfig
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=96, attention_head_dim=16, num_heads=None,
                     use_linear_projection=True, norm_num_groups=8)
    ref, ours = _pair(cfg, full_unet_targets(lr=1e-3, weight_decay=1e-2), torch.float32, prediction_type="v", ema_decay=0.99)
    n_params = sum(p.numel() for p in ours.unet.parameters())
    assert ours.arena.numel >= n_params and all(p.requires_grad for p in ours.unet.parameters())
    g = torch.Generator().manual_seed(3)
    lat, cond = torch.randn(2, 4, 16, 24, generator=g), torch.randn(2, 9, 96, generator=g)
    noise, t = torch.randn(2, 4, 16, 24, generator=g), torch.tensor([10, 900])
    refp = dict(ref.unet.named_parameters())
    for step in range(2):
        ours.optimizer.zero_grad()
        lo = ours.training_step({"latents": lat.to(DEV), "conds": cond.to(DEV)}, 0, noise.to(DEV), t.to(DEV))
        lo.backward()
        ref.optimizer.zero_grad(set_to_none=True)
        lr_ = ref.training_step({"latents": lat, "conds": cond}, noise, t)
        lr_.backward()
        assert abs(lo.item() - lr_.item()) <= 1e-4 * abs(lr_.item())
        num = sum((p.grad.cpu() - refp[n].grad).pow(2).sum().item() for n, p in ours.unet.named_parameters())
        den = sum(refp[n].grad.pow(2).sum().item() for n, _ in ours.unet.named_parameters())
        assert (num / den) ** 0.5 <= 1e-3
        # Adam's first steps move every element by ~lr * sign(g): feed both optimizers the oracle's exact gradients so that
        # the update rule (AdamW + EMA over the arena) is what is compared, not the sign of noise-level gradients
        for n, p in ours.unet.named_parameters():
            p.grad.copy_(refp[n].grad.to(DEV))
        ours.optimizer_step()
        ref.optimizer.step()
        ref.ema.update()
        for n, p in ours.unet.named_parameters():
            assert torch.allclose(p.detach().cpu(), refp[n], rtol=2e-5, atol=2e-7), (step, n)
            assert torch.allclose(ours.unet_ema.shadow_params[n].cpu(), ref.ema.shadow_params[n], rtol=2e-5, atol=2e-7), (step, n)
    assert ours.unet_ema.num_updates == 2
    assert set(ours.unet_ema.state_dict()["shadow_params"]) == set(refp)
