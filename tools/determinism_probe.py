"""Which kernels of a training step are bit-reproducible run to run on this GPU?  Per-op: same inputs twice -> equal?"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import fused, get_lora  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def twice(name, fn, n=4):
    outs = [fn() for _ in range(n)]
    same = all(all(torch.equal(a, b) for a, b in zip(outs[0], o)) for o in outs[1:])
    print(f"{name:60s} {'deterministic' if same else 'NOT deterministic'}")


# SDPA forward + backward at the step's shapes (heads 8; head_dim 40 / 80 / 160)
for (b, n, c, nk) in [(8, 4096, 320, 4096), (8, 1024, 640, 1024), (8, 256, 1280, 256), (8, 4096, 320, 77)]:
    q = torch.randn(b, 8, n, c // 8, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(b, 8, nk, c // 8, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(b, 8, nk, c // 8, device=dev, dtype=torch.bfloat16, requires_grad=True)
    do = torch.randn(b, 8, n, c // 8, device=dev, dtype=torch.bfloat16)

    def run():
        q.grad = k.grad = v.grad = None
        o = F.scaled_dot_product_attention(q, k, v)
        o.backward(do)
        return o.detach().clone(), q.grad.clone(), k.grad.clone(), v.grad.clone()
    twice(f"SDPA fwd+bwd  tokens {n} x {nk}  head_dim {c // 8}", run)

# cuDNN convolution forward + input gradient (frozen weights: no weight gradient)
for (cin, cout, hw, ks) in [(320, 320, 64, 3), (640, 1280, 16, 3), (960, 640, 32, 1)]:
    conv = torch.nn.Conv2d(cin, cout, ks, padding=ks // 2).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last).requires_grad_(False)
    x = torch.randn(8, cin, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dy = torch.randn(8, cout, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def run():
        x.grad = None
        y = conv(x)
        y.backward(dy)
        return y.detach().clone(), x.grad.clone()
    twice(f"conv{ks}x{ks} {cin}->{cout} @ {hw}^2 fwd + dgrad", run)

# this repo's GroupNorm(+SiLU) and LayerNorm kernels
gn = torch.nn.GroupNorm(32, 320).to(dev).to(torch.bfloat16).requires_grad_(False)
x = torch.randn(8, 320, 64, 64, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
dy = torch.randn_like(x)


def run_gn():
    x.grad = None
    y = fused.group_norm_act(gn, x, True)
    y.backward(dy)
    return y.detach().clone(), x.grad.clone()


twice("sdt group_norm_nhwc (+SiLU) fwd + bwd", run_gn)
ln = torch.nn.LayerNorm(320).to(dev).to(torch.bfloat16).requires_grad_(False)
xt = torch.randn(8, 4096, 320, device=dev, dtype=torch.bfloat16, requires_grad=True)
rt = torch.randn(8, 4096, 320, device=dev, dtype=torch.bfloat16)
dyt = torch.randn_like(xt)


def run_ln():
    xt.grad = None
    xs, y = fused.add_layer_norm(ln, xt, rt)
    (y * 1.0).backward(dyt)
    return y.detach().clone(), xt.grad.clone()


twice("sdt add + layer_norm fwd + bwd", run_ln)
# this repo's LoRA site: forward, dX, dA, dB
site = get_lora(torch.nn.Linear(320, 320).to(dev).to(torch.bfloat16).requires_grad_(False), 16, 16)
with torch.no_grad():
    site.lora_B.normal_(0, 0.1)
xs_ = torch.randn(32768, 320, device=dev, dtype=torch.bfloat16, requires_grad=True)
dys_ = torch.randn(32768, 320, device=dev, dtype=torch.bfloat16)


def run_site():
    xs_.grad = site.lora_A.grad = site.lora_B.grad = None
    y = site(xs_)
    y.backward(dys_)
    return y.detach().clone(), xs_.grad.clone(), site.lora_A.grad.clone(), site.lora_B.grad.clone()


twice("sdt LoRA site 32768x320x320 r16 fwd + dX + dA + dB", run_site)
