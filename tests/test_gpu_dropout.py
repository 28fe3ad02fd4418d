"""LoRA dropout on the rank path (``modules/lora.py:12`` -> loralib 0.1: ``result += (dropout(x) @ A.T @ B.T) * scaling``, training
mode only).  The mask comes from a counter-based generator inside ``sdt_lora_dropout``; the tests check its statistics, and --
with the mask the kernel actually drew, recovered through the same C-ABI call -- forward and gradients against the fp64 oracle
expression."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def test_mask_statistics_and_determinism(sdt_lib):
    from scal_sdt_b200 import _lib
    lib = _lib.load()
    M, K, p = 4096, 640, 0.3
    x = (torch.randn(M, K, device=DEV).abs() + 0.5).bfloat16()          # no zeros: a zero in xd is a dropped element
    seed = torch.tensor([1234567], dtype=torch.int64, device=DEV)

    def run(salt, sd=seed):
        out = torch.empty(M, 2 * K, device=DEV, dtype=torch.bfloat16)
        _lib.check(lib.sdt_lora_dropout(x.data_ptr(), out.data_ptr(), M, K, p, sd.data_ptr(), salt, 0, _lib.SDT_BF16, 0))
        torch.cuda.synchronize()
        return out
    a, b, c = run(7), run(7), run(8)
    assert torch.equal(a, b)                                             # same (seed, salt) -> same mask
    assert torch.equal(a[:, :K], x)                                      # first half is x itself
    keep_a, keep_c = a[:, K:] != 0, c[:, K:] != 0
    n = M * K
    for keep in (keep_a, keep_c):
        frac = keep.float().mean().item()
        assert abs(frac - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5 + 1e-4, frac      # binomial, 5 sigma
    agree = (keep_a == keep_c).float().mean().item()                     # independent masks agree with prob. p^2 + (1-p)^2
    assert abs(agree - (p * p + (1 - p) * (1 - p))) < 5e-3
    assert not torch.equal(run(7, torch.tensor([99], dtype=torch.int64, device=DEV)), a)     # another seed, another mask
    p_real = round(p * 65536) / 65536
    kept = a[:, K:][keep_a].float()
    assert torch.equal(kept, (x[keep_a].float() / (1 - p_real)).bfloat16().float())
    # per-column and per-row keep rates are flat (no structure along either axis)
    assert (keep_a.float().mean(0) - (1 - p)).abs().max() < 0.05 and (keep_a.float().mean(1) - (1 - p)).abs().max() < 0.1


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,N,r,bias", [(1000, 640, 640, 16, True), (300, 320, 1280, 4, False)])
def test_dropout_site_matches_oracle_with_the_drawn_mask(sdt_lib, M, K, N, r, bias, dtype):
    from scal_sdt_b200 import _lib, get_lora
    from scal_sdt_b200.lora import dropout_seed
    lib = _lib.load()
    p = 0.25
    torch.manual_seed(3)
    base = nn.Linear(K, N, bias=bias).to(DEV)
    with torch.no_grad():
        base.weight.copy_(base.weight.bfloat16().half().float())
    base.requires_grad_(False)
    site = get_lora(base, r, r, dropout=p)
    with torch.no_grad():
        site.lora_A.copy_(site.lora_A.bfloat16().half().float())
        site.lora_B.copy_((torch.randn(N, r, device=DEV) * 0.2).bfloat16().half().float())
    site.train()
    x = torch.randn(M, K, device=DEV).bfloat16().half().to(dtype).requires_grad_(True)
    dy = torch.randn(M, N, device=DEV).bfloat16().half().to(dtype)
    # the mask this forward will draw: same seed value, same salt, through the same entry point
    seed = dropout_seed(torch.device(DEV)).clone()
    salt = (site._drop_site << 32) | ((site._drop_calls + 1) & 0xFFFFFFFF)
    xcat = torch.empty(M, 2 * K, device=DEV, dtype=dtype)
    _lib.check(lib.sdt_lora_dropout(x.data_ptr(), xcat.data_ptr(), M, K, p, seed.data_ptr(), salt, 0, _lib.dtype_code(dtype), 0))
    keep = ((xcat[:, K:] != 0) | (x.detach() == 0)).double().cpu()
    y = site(x)
    y.backward(dy)
    # oracle: loralib's expression with that mask, fp64
    p_real = round(p * 65536) / 65536
    xr = x.detach().double().cpu().requires_grad_(True)
    A = site.lora_A.detach().double().cpu().requires_grad_(True)
    B = site.lora_B.detach().double().cpu().requires_grad_(True)
    w = base.weight.double().cpu()
    b = None if base.bias is None else base.bias.double().cpu()
    xd = xr * keep / (1 - p_real)
    yr = F.linear(xr, w, b) + (xd @ A.T @ B.T) * site.scaling
    yr.backward(dy.double().cpu())
    assert rel(y, yr) <= 2e-2 and rel(x.grad, xr.grad) <= 2e-2
    assert rel(site.lora_A.grad, A.grad) <= 2e-2 and rel(site.lora_B.grad, B.grad) <= 2e-2
    # the rank path really saw the dropped-out input: against the no-dropout expression the error is large
    y_nodrop = F.linear(xr, w, b) + (xr @ A.T @ B.T) * site.scaling
    assert rel(y, y_nodrop) > 5e-2
    # eval mode: dropout is inactive (loralib: nn.Dropout under .eval()), the plain fused path runs
    site.eval()
    with torch.no_grad():
        assert rel(site(x.detach()), y_nodrop) <= 2e-2


def test_trainer_with_dropout_targets_draws_new_masks_every_step(sdt_lib):
    import copy

    from scal_sdt_b200 import NoiseScheduler
    from scal_sdt_b200.lora import dropout_seed
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.trainer import LatentDiffusionTrainer
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    dev = torch.device(DEV)
    torch.manual_seed(5)
    unet = UNet2DConditionModel(UNetConfig.tiny()).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    tr = LatentDiffusionTrainer(unet, NoiseScheduler(), lora_unet_targets(rank=8, alpha=8, dropout=0.1), seed=3)
    assert tr._has_dropout
    with torch.no_grad():
        for _, m in tr.arena.sites:
            m.lora_B.normal_(0, 0.05)
    tr.arena.pack()
    g = torch.Generator().manual_seed(2)
    batch = {"latents": torch.randn(2, 4, 16, 16, generator=g).to(dev), "conds": torch.randn(2, 7, 64, generator=g).to(dev)}
    seeds = []
    for _ in range(3):
        loss = tr.step(batch)
        seeds.append(int(dropout_seed(dev).item()))
        assert torch.isfinite(loss)
    assert len(set(seeds)) == 3
    tr.enable_cuda_graph(batch)
    for _ in range(3):
        loss = tr.graphed_step(batch)
        seeds.append(int(dropout_seed(dev).item()))
    assert torch.isfinite(loss) and len(set(seeds)) == 6 and float(tr.arena.grads.abs().sum()) > 0
