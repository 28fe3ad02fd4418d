"""Host-model fallbacks of ``scal_sdt_b200.fused`` on CPU tensors: when the CUDA kernels do not apply (CPU oracle arm, trainable
norms, fp32) the helpers must compute exactly what the plain torch module sequence computes."""
import torch
import torch.nn.functional as F
from torch import nn

from scal_sdt_b200.fused import add_layer_norm, group_norm_act, layer_norm_supported, residual_bias_add
from scal_sdt_b200.unet import ResnetBlock2D


def test_add_layer_norm_cpu_fallback_is_plain_torch():
    torch.manual_seed(0)
    norm = nn.LayerNorm(64)
    x, r = torch.randn(2, 5, 64), torch.randn(2, 5, 64)
    assert not layer_norm_supported(norm, x)
    assert torch.equal(add_layer_norm(norm, x), norm(x))
    xs, y = add_layer_norm(norm, x, r)
    assert torch.equal(xs, x + r) and torch.equal(y, norm(x + r))


def test_group_norm_act_cpu_fallback_with_channel_bias():
    torch.manual_seed(1)
    norm = nn.GroupNorm(8, 64)
    x, tb = torch.randn(2, 64, 5, 7), torch.randn(2, 64)
    ref = F.silu(norm(x + tb[:, :, None, None]))
    assert torch.equal(group_norm_act(norm, x, True, chan_bias=tb), ref)
    assert torch.equal(group_norm_act(norm, x, False), norm(x))


def test_residual_bias_add_cpu_fallback():
    a, b, bias = torch.randn(2, 16, 3, 3), torch.randn(2, 16, 3, 3), torch.randn(16)
    assert torch.allclose(residual_bias_add(a, b, bias), a + b + bias[None, :, None, None])


def test_resnet_block_cpu_path_is_the_diffusers_sequence():
    """On the CPU (oracle / reference arm) ResnetBlock2D must be the plain diffusers ResnetBlock2D arithmetic."""
    torch.manual_seed(2)
    blk = ResnetBlock2D(32, 64, 128, 8)
    x, temb = torch.randn(2, 32, 6, 6), torch.randn(2, 128)
    h = blk.conv1(F.silu(blk.norm1(x)))
    h = h + blk.time_emb_proj(F.silu(temb))[:, :, None, None]
    h = blk.conv2(F.silu(blk.norm2(h)))
    ref = blk.conv_shortcut(x) + h
    assert torch.allclose(blk(x, temb), ref, atol=1e-6)
