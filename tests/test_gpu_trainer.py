"""GPU parity of the whole training step (noise/target -> LoRA UNet forward -> MSE -> backward -> AdamW/EMA) against
the CPU oracle trainer on identical synthetic latents, text embeddings, noise and timesteps."""
import pytest

pytestmark = pytest.mark.gpu


def test_smoke_step_matches_oracle():
    import __graft_entry__ as g
    g.smoke()


def test_prior_preservation_step_and_checkpoint_keys(sdt_lib):
    import copy

    import torch

    from oracle.ref_trainer import RefTrainer
    from scal_sdt_b200 import NoiseScheduler
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.trainer import LatentDiffusionTrainer
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    unet_cpu = UNet2DConditionModel(UNetConfig.tiny())
    with torch.no_grad():
        for p in unet_cpu.parameters():
            p.copy_(p.bfloat16().float())
    unet_gpu = copy.deepcopy(unet_cpu).to(dev).to(torch.bfloat16)
    targets = lora_unet_targets(rank=8, alpha=8, feed_forward=False, projections=False)     # attention only (128 sites)
    ref = RefTrainer(unet_cpu, targets, prior_preservation=True, prior_loss_weight=0.5)
    ours = LatentDiffusionTrainer(unet_gpu, NoiseScheduler(), copy.deepcopy(targets),
                                  prior_preservation={"enabled": True, "prior_loss_weight": 0.5}, seed=1)
    assert len(ours.arena.sites) == 128
    g = torch.Generator().manual_seed(2)
    refm = dict(ref.unet.named_modules())
    with torch.no_grad():
        for name, m in ours.arena.sites:
            b = (torch.randn(m.lora_B.shape, generator=g) * 0.05).bfloat16().float()
            a = m.lora_A.detach().cpu().bfloat16().float()
            m.lora_A.copy_(a); m.lora_B.copy_(b)
            refm[name].lora_A.copy_(a); refm[name].lora_B.copy_(b)
    ours.arena.pack()
    lat = torch.randn(4, 4, 16, 16, generator=g)
    cond = torch.randn(4, 7, 64, generator=g).bfloat16().float()
    noise = torch.randn(4, 4, 16, 16, generator=g)
    t = torch.tensor([0, 999, 400, 12])
    ours.optimizer.zero_grad()
    loss = ours.training_step({"latents": lat.to(dev), "conds": cond.to(dev)}, 0, noise.to(dev), t.to(dev))
    loss.backward()
    rl = ref.training_step({"latents": lat, "conds": cond}, noise, t)
    rl.backward()
    assert abs(loss.item() - rl.item()) <= 2e-2 * abs(rl.item())
    parts = ours.criterion.last_parts.cpu()
    assert abs(parts[0] - (parts[1] + 0.5 * parts[2])) < 1e-6
    sd = ours.checkpoint_state_dict()
    assert all(k.startswith("unet.") and k.endswith(("lora_A", "lora_B")) for k in sd)
    assert len(sd) == 256 and "unet.mid_block.attentions.0.transformer_blocks.0.attn1.to_q.lora_A" in sd
    ours.criterion.raise_if_nan()


def test_cuda_graph_step_matches_eager(sdt_lib):
    """The captured step reproduces the eager loss and LoRA gradients on the same inputs (lr = 0 keeps the parameters
    fixed, so the two can be compared directly), and a non-zero lr makes the loss go down over replays."""
    import copy

    import torch

    from scal_sdt_b200 import NoiseScheduler
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.trainer import LatentDiffusionTrainer
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    base = UNet2DConditionModel(UNetConfig.tiny()).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)

    def make(lr):
        tr = LatentDiffusionTrainer(copy.deepcopy(base), NoiseScheduler(), lora_unet_targets(rank=8, alpha=8, lr=lr),
                                    optimizer_params={"lr": lr, "beta1": 0.9, "beta2": 0.999, "weight_decay": 0.0, "eps": 1e-8},
                                    ema={"enabled": True, "decay": 0.99}, seed=3)
        g = torch.Generator(device=dev).manual_seed(1)
        with torch.no_grad():
            for _, m in tr.arena.sites:
                m.lora_B.normal_(0, 0.05, generator=g)
        tr.arena.pack()
        return tr

    g = torch.Generator().manual_seed(2)
    batch = {"latents": torch.randn(2, 4, 16, 16, generator=g).to(dev), "conds": torch.randn(2, 7, 64, generator=g).to(dev)}
    tr = make(0.0)
    tr.enable_cuda_graph(batch)
    loss_g = tr.graphed_step(batch).clone()
    grads_g = tr.arena.grads.clone()
    tr.optimizer.zero_grad()
    loss_e = tr.training_step(batch, 0, tr._g_noise, tr._g_t)
    loss_e.backward()
    # The LoRA kernels are bit-reproducible (test_gpu_lora.py::test_lora_gradients_are_bit_reproducible); the host model around
    # them is not: torch's SDPA backward and the GroupNorm statistics accumulate with float atomics (tools/determinism_probe.py),
    # so two executions of the same step agree to rounding, not to the bit.
    assert abs(loss_g.item() - loss_e.item()) <= 1e-3 * abs(loss_e.item())
    assert (grads_g - tr.arena.grads).norm() <= 2e-2 * tr.arena.grads.norm()
    tr2 = make(2e-3)
    tr2.enable_cuda_graph(batch)
    first = tr2.graphed_step(batch).item()
    for _ in range(20):
        last = tr2.graphed_step(batch).item()
    assert tr2.optimizer.step_count >= 21 and tr2.unet_ema.num_updates >= 21


def _tiny_trainer(lr, seed=3, ema=True):
    import copy

    import torch

    from scal_sdt_b200 import NoiseScheduler
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.trainer import LatentDiffusionTrainer
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    base = UNet2DConditionModel(UNetConfig.tiny()).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    tr = LatentDiffusionTrainer(copy.deepcopy(base), NoiseScheduler(), lora_unet_targets(rank=8, alpha=8, lr=lr),
                                optimizer_params={"lr": lr, "beta1": 0.9, "beta2": 0.999, "weight_decay": 1e-2, "eps": 1e-8},
                                ema={"enabled": ema, "decay": 0.99}, seed=seed)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for _, m in tr.arena.sites:
            m.lora_B.normal_(0, 0.05, generator=g)
    tr.arena.pack()
    return tr


def test_enabling_the_graph_does_not_train(sdt_lib):
    """Warm-up + capture leave parameters, AdamW moments, EMA shadow, counters and the noise generator exactly where they
    were; graphed steps then follow the eager trainer's trajectory (same seeds) instead of being three updates ahead."""
    import torch
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    batch = {"latents": torch.randn(2, 4, 16, 16, generator=g).to(dev), "conds": torch.randn(2, 7, 64, generator=g).to(dev)}
    tr = _tiny_trainer(1e-3)
    before = (tr.arena.params.clone(), tr.optimizer.exp_avg.clone(), tr.optimizer.exp_avg_sq.clone(),
              tr.unet_ema._shadow_flat.clone(), tr.generator.get_state().clone())
    tr.enable_cuda_graph(batch)
    assert torch.equal(tr.arena.params, before[0]) and torch.equal(tr.optimizer.exp_avg, before[1])
    assert torch.equal(tr.optimizer.exp_avg_sq, before[2]) and torch.equal(tr.unet_ema._shadow_flat, before[3])
    assert torch.equal(tr.generator.get_state(), before[4])
    assert tr.global_step == 0 and tr.optimizer.step_count == 0 and tr.unet_ema.num_updates == 0
    eager = _tiny_trainer(1e-3)
    for i in range(3):
        lg = tr.graphed_step(batch).item()
        le = eager.step(batch).item()
        assert abs(lg - le) <= 2e-3 * abs(le), (i, lg, le)
    assert tr.global_step == eager.global_step == 3 and tr.optimizer.step_count == eager.optimizer.step_count == 3
    assert tr.unet_ema.num_updates == eager.unet_ema.num_updates == 3
    # Adam's first steps move every element by ~lr * sign(g): compare the trajectories by the moments, not by sign flips
    assert (tr.optimizer.exp_avg - eager.optimizer.exp_avg).norm() <= 2e-2 * eager.optimizer.exp_avg.norm()
    assert (tr.unet_ema._shadow_flat - eager.unet_ema._shadow_flat).norm() <= 1e-2 * eager.unet_ema._shadow_flat.norm()


def test_bf16_forward_under_ema_weights_uses_the_ema_weights(sdt_lib):
    """``average_parameters()`` / ``apply()`` write the fp32 masters behind the optimizer's back; the bf16 operands the
    kernels read must follow (arena-owned operands and the per-module cache alike), and be restored on exit."""
    import torch
    from torch import nn

    from scal_sdt_b200 import ExponentialMovingAverage, get_lora
    dev = torch.device("cuda:0")
    tr = _tiny_trainer(1e-3)
    g = torch.Generator().manual_seed(2)
    lat = torch.randn(2, 4, 16, 16, generator=g).to(dev).bfloat16()
    cond = torch.randn(2, 7, 64, generator=g).to(dev).bfloat16()
    t = torch.tensor([5, 500], device=dev)
    with torch.no_grad():
        tr.unet_ema._shadow_flat.mul_(0.5)                     # shadow != parameters
        packed_train = tr.arena.packed.clone()
        with tr.unet_ema.average_parameters():
            packed_ema = tr.arena.packed.clone()               # what the kernels read inside the context
            y_ema = tr.unet(lat, t, cond).sample.float()
        packed_back = tr.arena.packed.clone()
        y_back = tr.unet(lat, t, cond).sample.float()
        # the same EMA values loaded the "official" way: copy into the masters, repack
        tr.arena.params.copy_(tr.unet_ema._shadow_flat)
        tr.arena.pack()
        y_loaded = tr.unet(lat, t, cond).sample.float()
    assert torch.equal(packed_back, packed_train)               # restored on exit
    assert torch.equal(packed_ema, tr.arena.packed)             # EMA operands inside == EMA values loaded and packed
    assert not torch.equal(packed_ema, packed_train)
    # forward outputs (the host model's GroupNorm statistics use float atomics: compare to rounding, not to the bit)
    assert (y_ema - y_loaded).norm() <= 2e-2 * y_loaded.norm()
    assert (y_ema - y_back).norm() > 4 * (y_ema - y_loaded).norm()
    # module without an arena: the per-module operand cache is keyed on Parameter versions, which .data writes do not bump
    torch.manual_seed(0)
    lin = get_lora(nn.Linear(64, 32).to(dev).requires_grad_(False), 4, 4)
    with torch.no_grad():
        lin.lora_B.normal_(0, 0.3)
    ema = ExponentialMovingAverage(lin, 0.9)
    x = torch.randn(8, 64, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        y0 = lin(x).clone()
        for s in ema.shadow_params.values():
            s.mul_(0.25)
        with ema.average_parameters():
            y1 = lin(x).clone()
        y2 = lin(x).clone()
    assert torch.equal(y0, y2) and not torch.equal(y0, y1)
