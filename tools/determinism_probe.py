"""Is the tiny UNet's bf16 forward / a whole training step bit-reproducible run to run?  (PDL on/off via SDT_PDL.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_trainer import _tiny_trainer  # noqa: E402

dev = torch.device("cuda:0")
tr = _tiny_trainer(1e-3)
g = torch.Generator().manual_seed(2)
lat = torch.randn(2, 4, 16, 16, generator=g).to(dev).bfloat16()
cond = torch.randn(2, 7, 64, generator=g).to(dev).bfloat16()
t = torch.tensor([5, 500], device=dev)
with torch.no_grad():
    ys = [tr.unet(lat, t, cond).sample.clone() for _ in range(4)]
print("PDL", os.environ.get("SDT_PDL", "1"), "forward equal to first:", [bool(torch.equal(ys[0], y)) for y in ys],
      "max diff", [float((ys[0].float() - y.float()).abs().max()) for y in ys])
# per-module: which layer is the first to differ between two forwards?
outs = [{}, {}]
for k in range(2):
    hooks = []
    for name, m in tr.unet.named_modules():
        if not list(m.children()):
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name, k=k: outs[k].__setitem__(name, o.detach().clone() if torch.is_tensor(o) else None)))
    with torch.no_grad():
        tr.unet(lat, t, cond)
    for h in hooks:
        h.remove()
bad = [n for n in outs[0] if outs[0][n] is not None and not torch.equal(outs[0][n], outs[1][n])]
print("first differing leaf modules:", bad[:5], "of", len(bad))
# whole-step gradients twice from the same state
batch = {"latents": lat.float(), "conds": cond.float()}
noise = torch.randn(2, 4, 16, 16, device=dev)
grads = []
for _ in range(3):
    tr.optimizer.zero_grad()
    loss = tr.training_step(batch, 0, noise, t)
    loss.backward()
    grads.append((loss.item(), tr.arena.grads.clone()))
print("loss", [g[0] for g in grads], "grads equal to first:", [bool(torch.equal(grads[0][1], g[1])) for g in grads],
      "rel diff", [float((grads[0][1] - g[1]).norm() / grads[0][1].norm()) for g in grads])
