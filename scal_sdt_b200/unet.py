"""SD-shaped conditional UNet skeleton with diffusers' module names (host model for whole-step measurements).

``diffusers`` is not available offline, so the UNet the reference loads (``UNet2DConditionModel``,
``modules/model.py:82-91,304``) is restated here in plain torch with the SAME submodule names, so that
``configs/optim_targets/*.yaml`` resolve unmodified (``down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k`` ...;
names corroborated by ``modules/convert/diffusers_to_sd.py:5-77``; shapes by ``modules/convert/sd_to_diffusers.py:175-209``).
Everything in this file is OUTSIDE the hot path (cuDNN convolutions, SDPA attention, norms) and stays torch; the
hot path enters when ``config_module`` replaces the targeted Linear / 1x1-Conv2d modules with the CUDA-backed LoRA
modules.  Weights are random-init (BASELINE.json configs say so).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .fused import add_layer_norm, fused_enabled, geglu, geglu_supported, group_norm_act, residual_bias_add
from .lora import geglu_projection, geglu_projection_supported, project_group


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: Optional[int] = None   # SD2.x: 64 channels per head
    num_heads: Optional[int] = 8               # SD1.x: 8 heads on every level
    use_linear_projection: bool = False        # SD2.x: proj_in / proj_out are Linear
    norm_num_groups: int = 32
    down_has_attn: tuple = (True, True, True, False)

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sd2x() -> "UNetConfig":
        return UNetConfig(cross_attention_dim=1024, attention_head_dim=64, num_heads=None, use_linear_projection=True)

    @staticmethod
    def tiny(cross_attention_dim=64) -> "UNetConfig":
        """Same topology at toy widths -- for CPU tests of the walker / checkpoint layout."""
        return UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=cross_attention_dim, num_heads=4,
                          norm_num_groups=8)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """Sinusoidal embedding, cos first (``flip_sin_to_cos=True``, ``freq_shift=0`` as SD uses)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    args = timesteps.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def _frozen_biases(self):
        """(conv1.bias as the activation dtype, conv2.bias [+ conv_shortcut.bias] as f32) when every conv bias is frozen: the
        biases are then folded into the passes that follow the convolutions instead of torch's per-conv broadcast add."""
        convs = [self.conv1, self.conv2] + ([self.conv_shortcut] if self.conv_shortcut is not None else [])
        if any(c.bias is None or c.bias.requires_grad for c in convs):
            return None
        key = tuple((c.bias.data_ptr(), c.bias._version) for c in convs)
        cache = getattr(self, "_bias_cache", None)
        if cache is None or cache[0] != key:
            tail = self.conv2.bias.detach().float()
            if self.conv_shortcut is not None:
                tail = tail + self.conv_shortcut.bias.detach().float()
            cache = (key, self.conv1.bias.detach(), tail.contiguous())
            self._bias_cache = cache
        return cache[1], cache[2]

    def forward(self, x, temb):
        folded = self._frozen_biases() if (x.is_cuda and x.dtype == torch.bfloat16 and fused_enabled()) else None
        if folded is None:                          # CPU oracle arm / trainable convolutions: the plain module calls
            h = self.conv1(group_norm_act(self.norm1, x, True))
            # h + temb only feeds norm2: the broadcast add is folded into the norm (torch runs it as a non-vectorised kernel)
            h = self.conv2(group_norm_act(self.norm2, h, True, chan_bias=self.time_emb_proj(F.silu(temb))))
            if self.conv_shortcut is not None:
                x = self.conv_shortcut(x)
            return x + h
        b1, b_tail = folded
        h = F.conv2d(group_norm_act(self.norm1, x, True), self.conv1.weight, None, 1, 1)
        # conv1's bias and the time embedding are both per-(sample, channel) constants in front of norm2
        h = group_norm_act(self.norm2, h, True, chan_bias=self.time_emb_proj(F.silu(temb)) + b1)
        h = F.conv2d(h, self.conv2.weight, None, 1, 1)
        if self.conv_shortcut is not None:
            x = F.conv2d(x, self.conv_shortcut.weight, None)
        return residual_bias_add(x, h, b_tail)


class CrossAttention(nn.Module):
    def __init__(self, dim: int, context_dim: Optional[int], heads: int):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(context_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(context_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self._kv_pre = None          # (k, v) of the text context when UNet.forward projected it ahead of the blocks

    def forward(self, x, context=None):
        b, n, c = x.shape
        # same-shape projections of one input go out as ONE grouped launch when they are LoRA sites (fused
        # q/k/v); plain nn.Linear modules (CPU oracle arm, un-targeted sites) are simply called one by one
        if context is None:
            q, k, v = project_group([self.to_q, self.to_k, self.to_v], x)
            n_kv = n
        else:
            q = self.to_q(x)
            kv = self._kv_pre
            if kv is not None:                      # projected up front together with other blocks' (UNet.forward)
                self._kv_pre = None
                k, v = kv
            else:
                k, v = project_group([self.to_k, self.to_v], context)
            n_kv = context.shape[1]
        q = q.view(b, n, self.heads, c // self.heads).transpose(1, 2)
        k = k.view(b, n_kv, self.heads, c // self.heads).transpose(1, 2)
        v = v.view(b, n_kv, self.heads, c // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, n, c)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        if fused_enabled() and geglu_projection_supported(self.proj, x):
            return geglu_projection(self.proj, x)      # the activation comes out of the projection's epilogue
        y = self.proj(x)
        if geglu_supported(y):                 # one vectorised pass per direction
            return geglu(y)
        h, gate = y.chunk(2, dim=-1)           # host model on the CPU (oracle / reference arm)
        return h * F.gelu(gate)


def _accepts_residual(m) -> bool:
    """LoRA sites add a residual in their own epilogue (``LoRALinear.forward(x, residual=...)``); plain modules do not."""
    import os

    from .lora import _LoRABase
    # Opt-in (SDT_FUSED_RESIDUAL=1): measured neutral on B200 (profiles/README.md: 28.60 vs 28.63 ms per step) -- the 32 torch adds
    # it removes (0.19 ms) come back as residual reads inside store-bound epilogues (+0.15 ms of GEMM time)
    return isinstance(m, _LoRABase) and fused_enabled() and os.environ.get("SDT_FUSED_RESIDUAL", "0") == "1"


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x, residual=None):
        """``residual``: returns ``residual + net(x)`` -- the add goes into the last projection's epilogue when it is a LoRA site"""
        x = self.net[1](self.net[0](x))
        out = self.net[2]
        if residual is None:
            return out(x)
        if _accepts_residual(out) and x.is_cuda:
            return out(x, residual=residual)
        return residual + out(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, context_dim: int, heads: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        # residual add + the LayerNorm that follows it are one pass
        x, h = add_layer_norm(self.norm2, x, self.attn1(add_layer_norm(self.norm1, x)))
        x, h = add_layer_norm(self.norm3, x, self.attn2(h, context))
        return self.ff(h, residual=x)


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, context_dim: int, heads: int, groups: int, linear_proj: bool):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, context_dim, heads)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, context):
        b, c, h, w = x.shape
        res = x
        y = group_norm_act(self.norm, x, False)
        if self.linear_proj:
            y = self.proj_in(y.permute(0, 2, 3, 1).reshape(b, h * w, c))
        else:
            y = self.proj_in(y).permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            y = blk(y, context)
        fuse = _accepts_residual(self.proj_out) and res.is_cuda       # the residual add in the projection's epilogue
        if self.linear_proj:
            if fuse:
                return self.proj_out(y, residual=res.permute(0, 2, 3, 1).reshape(b, h * w, c)).reshape(b, h, w, c).permute(0, 3, 1, 2)
            y = self.proj_out(y).reshape(b, h, w, c).permute(0, 3, 1, 2)
        else:
            if fuse:
                return self.proj_out(y.reshape(b, h, w, c).permute(0, 3, 1, 2), residual=res)
            y = self.proj_out(y.reshape(b, h, w, c).permute(0, 3, 1, 2))
        return y + res


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin, cout, temb_dim, heads, with_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim, cfg.norm_num_groups)
                                      for i in range(cfg.layers_per_block)])
        if with_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, cfg.cross_attention_dim, heads, cfg.norm_num_groups,
                                                                cfg.use_linear_projection)
                                             for _ in range(cfg.layers_per_block)])
        self.has_attn = with_attn
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.has_down = add_down

    def forward(self, x, temb, context):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, context)
            outs.append(x)
        if self.has_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, c, temb_dim, heads):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim, cfg.norm_num_groups) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, cfg.cross_attention_dim, heads, cfg.norm_num_groups,
                                                            cfg.use_linear_projection)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin, cout, prev, temb_dim, heads, with_attn, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = prev if j == 0 else cout
            self.resnets.append(ResnetBlock2D(rin + skip, cout, temb_dim, cfg.norm_num_groups))
        if with_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, cfg.cross_attention_dim, heads, cfg.norm_num_groups,
                                                                cfg.use_linear_projection) for _ in range(n)])
        self.has_attn = with_attn
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.has_up = add_up

    def forward(self, x, skips, temb, context):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, context)
        if self.has_up:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: Optional[UNetConfig] = None):
        super().__init__()
        cfg = cfg or UNetConfig.sd15()
        self.config = cfg
        ch = cfg.block_out_channels
        temb_dim = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb_dim)

        def heads_for(c):
            return cfg.num_heads if cfg.num_heads is not None else c // cfg.attention_head_dim

        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            self.down_blocks.append(DownBlock(cfg, cin, cout, temb_dim, heads_for(cout), cfg.down_has_attn[i], i < len(ch) - 1))
        self.mid_block = MidBlock(cfg, ch[-1], temb_dim, heads_for(ch[-1]))
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        up_attn = list(reversed(cfg.down_has_attn))
        cout = rev[0]
        for i, c in enumerate(rev):
            prev, cout = cout, c
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(UpBlock(cfg, cin, cout, prev, temb_dim, heads_for(cout), up_attn[i], i < len(ch) - 1))
        self.channels_last = True          # activation memory format of the convolutional trunk
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    @property
    def device(self):
        return self.conv_in.weight.device

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def _project_context(self, ctx) -> None:
        """The text context is the input of EVERY cross-attention's to_k / to_v, so those projections do not have to wait
        for their block: sites of equal width are projected up front, ``MAX_GROUP`` per launch (two blocks' k and v), and
        each attention picks its pair up when it runs.  Only LoRA sites on CUDA take part (``project_group`` decides)."""
        from . import _lib
        from .lora import LoRALinear
        attn2s = self.__dict__.get("_attn2_modules")
        if attn2s is None:                        # module tree walk once, not per forward
            attn2s = [m.attn2 for m in self.modules() if isinstance(m, BasicTransformerBlock)]
            self.__dict__["_attn2_modules"] = attn2s
        by_width: dict[int, list] = {}
        for a in attn2s:
            a._kv_pre = None                      # never reuse a pair from an earlier (possibly interrupted) forward
            if ctx.is_cuda and isinstance(a.to_k, LoRALinear) and isinstance(a.to_v, LoRALinear):
                by_width.setdefault(a.to_k.out_features, []).append(a)
        from .lora import multi_projectable, project_multi
        every = [a for attns in by_width.values() for a in attns]
        sites = [p for a in every for p in (a.to_k, a.to_v)]
        ctx2 = ctx.reshape(-1, ctx.shape[-1])
        if sites and multi_projectable(sites, ctx2):
            # ONE launch for to_k / to_v of every cross-attention of the UNet (different widths, same context)
            outs = project_multi(sites, ctx)
            for j, a in enumerate(every):
                a._kv_pre = (outs[2 * j], outs[2 * j + 1])
            return
        per = _lib.MAX_GROUP // 2
        for attns in by_width.values():
            for i in range(0, len(attns), per):
                chunk = attns[i:i + per]
                outs = project_group([p for a in chunk for p in (a.to_k, a.to_v)], ctx)
                for j, a in enumerate(chunk):
                    a._kv_pre = (outs[2 * j], outs[2 * j + 1])

    def forward(self, sample, timesteps, encoder_hidden_states):
        """Returns an object with ``.sample`` like diffusers (``modules/model.py:304``)."""
        dt = self.dtype
        temb = self.time_embedding(timestep_embedding(timesteps, self.config.block_out_channels[0]).to(dt))
        ctx = encoder_hidden_states.to(dt)
        self._project_context(ctx)
        fmt = torch.channels_last if self.channels_last else torch.contiguous_format
        x = self.conv_in(sample.to(dt).contiguous(memory_format=fmt))
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, ctx)
            skips.extend(outs)
        x = self.mid_block(x, temb, ctx)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, ctx)
        x = self.conv_out(group_norm_act(self.conv_norm_out, x, True))
        return SimpleNamespace(sample=x)
