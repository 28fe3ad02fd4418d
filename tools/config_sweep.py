"""Step time of the other BASELINE configurations' shapes on one GPU (graph replay, CUDA events): cfg2 (bench workload),
cfg3 (rank 64 on attention + FF, EMA), cfg4 (largest bucket 768x1024 -> 96x128 latents, DreamBooth halves, rank 16)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import GradExchange, NoiseScheduler  # noqa: E402
from scal_sdt_b200.targets import lora_unet_targets  # noqa: E402
from scal_sdt_b200.trainer import LatentDiffusionTrainer  # noqa: E402
from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig  # noqa: E402

dev = torch.device("cuda:0")
CASES = [
    ("cfg2  r16 all targets, 8 x 64x64", dict(rank=16, projections=True), 8, 64, 64, {}, {}),
    ("cfg3  r64 attn+FF + EMA, 8 x 64x64", dict(rank=64, alpha=64, projections=False), 8, 64, 64, {"enabled": True, "decay": 0.995}, {}),
    ("cfg4  r16 all targets, DreamBooth 4+4 x 96x128 (768x1024)", dict(rank=16, projections=True), 8, 96, 128, {},
     {"enabled": True, "prior_loss_weight": 1.0}),
    ("cfg4b r16 all targets, 8 x 64x96 (512x768 bucket)", dict(rank=16, projections=True), 8, 64, 96, {}, {}),
]
for name, tkw, B, h, w, ema, pp in CASES:
    torch.manual_seed(0)
    with torch.device(dev):
        unet = UNet2DConditionModel(UNetConfig.sd15())
    unet = unet.to(torch.bfloat16).to(memory_format=torch.channels_last)
    tr = LatentDiffusionTrainer(unet, NoiseScheduler(prediction_type="epsilon"), lora_unet_targets(**tkw), batch_size=B,
                                exchange=GradExchange(0, 1), seed=0, ema=ema or None, prior_preservation=pp or None)
    g = torch.Generator().manual_seed(1)
    batch = {"latents": torch.randn(B, 4, h, w, generator=g).to(dev), "conds": torch.randn(B, 77, 768, generator=g).to(dev)}
    for _ in range(3):
        tr.step(batch)
    torch.cuda.synchronize()
    tr.enable_cuda_graph(batch)
    for _ in range(2):
        tr.graphed_step(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        loss = tr.graphed_step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    params = sum(p.numel() for p, _, _ in tr.arena.slots)
    print(f"{name}: {ms:7.2f} ms/step  {B / ms * 1e3:7.1f} latents/s  ({params} LoRA params, loss {loss.item():.4f}, "
          f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB peak)", flush=True)
    tr.release_cuda_graph()
    del tr, unet
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
