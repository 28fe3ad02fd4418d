"""GPU parity of the fused (residual add +) LayerNorm kernels against plain torch fp32 on the same bf16 inputs.

Reference = ``F.layer_norm`` evaluated in fp32 on the bf16-rounded operands, and its autograd for the backward; the residual
stream ``x + res`` must be bit-identical to torch's bf16 add.  Tolerance: the output is rounded once to bf16, so the
relative Frobenius error bound is half a bf16 ulp (2^-9 = 1.95e-3) plus fp32 reduction noise -> 3e-3."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 3e-3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def make_norm(C, seed):
    torch.manual_seed(seed)
    n = nn.LayerNorm(C)
    with torch.no_grad():
        n.weight.normal_(1.0, 0.2)
        n.bias.normal_(0.0, 0.2)
    n = n.to(DEV).to(torch.bfloat16)
    n.requires_grad_(False)
    return n


SHAPES = [(2, 256, 320), (1, 1000, 640), (3, 50, 1280), (5, 7, 64), (2, 33, 2048), (8, 4096, 320)]


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{s[0]}x{s[1]}x{s[2]}" for s in SHAPES])
@pytest.mark.parametrize("with_res", [False, True])
def test_add_layer_norm_forward_backward(sdt_lib, shape, with_res):
    from scal_sdt_b200.fused import add_layer_norm, layer_norm_supported
    C = shape[-1]
    norm = make_norm(C, 3)
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(*shape, generator=g) * 2 + 0.5).bfloat16().to(DEV)
    res = torch.randn(*shape, generator=g).bfloat16().to(DEV) if with_res else None
    dy = torch.randn(*shape, generator=g).bfloat16().to(DEV)
    dxs = torch.randn(*shape, generator=g).bfloat16().to(DEV) if with_res else None
    assert layer_norm_supported(norm, x)
    xo = x.clone().requires_grad_(True)
    ro = res.clone().requires_grad_(True) if with_res else None
    out = add_layer_norm(norm, xo, ro)
    # reference: torch fp32 on the same operands
    w32, b32 = norm.weight.float(), norm.bias.float()
    xr = x.float().requires_grad_(True)
    if with_res:
        xs_o, y_o = out
        xs_ref = (x + res)                                   # torch's bf16 add
        assert torch.equal(xs_o, xs_ref), "residual stream must be torch's bf16 add bit for bit"
        rr = res.float().requires_grad_(True)
        xs32 = (xr + rr)
        # normalise the ROUNDED stream, as the kernel and torch both do; keep the graph through the straight-through sum
        xs_in = xs32 + (xs_ref.float() - xs32).detach()
        y_ref = F.layer_norm(xs_in, (C,), w32, b32, norm.eps)
        torch.autograd.backward([y_ref, xs_in], [dy.float(), dxs.float()])
        torch.autograd.backward([y_o, xs_o], [dy, dxs])
        assert rel(y_o, y_ref) <= TOL, ("y", rel(y_o, y_ref))
        assert rel(xo.grad, xr.grad) <= TOL, ("dx", rel(xo.grad, xr.grad))
        assert torch.equal(xo.grad, ro.grad)
    else:
        y_o = out
        y_ref = F.layer_norm(xr, (C,), w32, b32, norm.eps)
        y_ref.backward(dy.float())
        y_o.backward(dy)
        assert rel(y_o, y_ref) <= TOL, ("y", rel(y_o, y_ref))
        assert rel(xo.grad, xr.grad) <= TOL, ("dx", rel(xo.grad, xr.grad))
    # and against torch's own bf16 LayerNorm kernel: same rounding point, so nearly always the same bf16 value
    y_t = F.layer_norm(xs_o if with_res else x, (C,), norm.weight, norm.bias, norm.eps)
    assert (y_o.float() - y_t.float()).abs().max() <= 2 ** -6 * max(1.0, y_t.float().abs().max().item())


def test_layer_norm_falls_back_for_trainable_affine(sdt_lib):
    """Full fine-tune (cfg5) trains the norm parameters: the frozen-affine kernel must not be used."""
    from scal_sdt_b200.fused import add_layer_norm, layer_norm_supported
    norm = nn.LayerNorm(320).to(DEV).to(torch.bfloat16)
    x = torch.randn(2, 16, 320, device=DEV, dtype=torch.bfloat16)
    assert not layer_norm_supported(norm, x)
    y = add_layer_norm(norm, x)
    y.sum().backward()
    assert norm.weight.grad is not None
