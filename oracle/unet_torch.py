"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

Pure-torch restatement of the diffusers==0.11.0 U-Net module tree that the reference's
hot path differentiates.  The reference (`/root/reference/requirements.txt:4`) pins
diffusers 0.11.0, which is neither vendored in the reference nor installable here (no
network), so its *published* architecture is restated with the same attribute names the
reference touches:

  * `src/utils/utils.py:438-527`  (`get_h`, SD):   time_proj, time_embedding, conv_in,
    down_blocks[i](hidden_states, temb[, encoder_hidden_states]) -> (x, res_tuple),
    .has_cross_attention, mid_block(x, emb, encoder_hidden_states=), up_blocks[i].resnets,
    up_blocks[i](hidden_states, temb, res_hidden_states_tuple, ...), .dtype
  * `src/utils/utils.py:114-163`  (`get_h_uncond`): same for UNet2DModel.

State-dict key names follow diffusers so real checkpoints would load.  Cross-checked
against the only U-Net whose source *is* in the reference repo
(`src/models/ddpm/diffusion.py:22-126`, `:816-966`) by `tests/test_oracle.py` (weight mapping
UNet2DModel <-> in-repo DDPM, run in the authoring container where /root/reference exists).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False,
                           downscale_freq_shift=1.0, scale=1.0, max_period=10000):
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels,
                                      flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------------------
# resnet / resampling
# --------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5,
                 output_scale_factor=1.0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.output_scale_factor = output_scale_factor
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = None
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0)

    def forward(self, input_tensor, temb):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        t = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = h + t
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels, out_channels=None, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels, out_channels=None):
        super().__init__()
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)

    def forward(self, x, output_size=None):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x)


# --------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------
class CrossAttention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False):
        super().__init__()
        inner = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def _split(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def _merge(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def forward(self, hidden_states, context=None):
        q = self.to_q(hidden_states)
        context = context if context is not None else hidden_states
        k, v = self.to_k(context), self.to_v(context)
        q, k, v = self._split(q), self._split(k), self._split(v)
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=self.scale)
        probs = scores.softmax(dim=-1)
        out = self._merge(torch.bmm(probs, v))
        return self.to_out[1](self.to_out[0](out))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, cross_attention_dim, heads, dim_head)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, x, encoder_hidden_states=None, timestep=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=encoder_hidden_states) + x
        x = self.ff(self.norm3(x)) + x
        return x


@dataclass
class _Sample:
    sample: torch.Tensor


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, groups=32,
                 use_linear_projection=False, num_layers=1):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)
             for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None):
        b, c, hh, ww = hidden_states.shape
        residual = hidden_states
        x = self.norm(hidden_states)
        if not self.use_linear_projection:
            x = self.proj_in(x)
            inner = x.shape[1]
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, inner)
        else:
            inner = x.shape[1]
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, inner)
            x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states, timestep=timestep)
        if not self.use_linear_projection:
            x = x.reshape(b, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
            x = self.proj_out(x)
        else:
            x = self.proj_out(x)
            x = x.reshape(b, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
        return _Sample(sample=x + residual)


class AttentionBlock(nn.Module):
    """diffusers 0.11.0 `AttentionBlock` (UNet2DModel self-attention; in-repo analogue
    `src/models/ddpm/diffusion.py:914-966`)."""

    def __init__(self, channels, num_head_channels=None, groups=32, eps=1e-5,
                 rescale_output_factor=1.0):
        super().__init__()
        self.channels = channels
        self.num_heads = channels // num_head_channels if num_head_channels is not None else 1
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.rescale_output_factor = rescale_output_factor
        self.proj_attn = nn.Linear(channels, channels, 1)

    def _split(self, t):
        b, n, c = t.shape
        h = self.num_heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def _merge(self, t):
        bh, n, d = t.shape
        h = self.num_heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def forward(self, hidden_states):
        residual = hidden_states
        b, c, hh, ww = hidden_states.shape
        x = self.group_norm(hidden_states)
        x = x.view(b, c, hh * ww).transpose(1, 2)
        q, k, v = self.query(x), self.key(x), self.value(x)
        scale = 1 / math.sqrt(self.channels / self.num_heads)
        q, k, v = self._split(q), self._split(k), self._split(v)
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=scale)
        probs = torch.softmax(scores.float(), dim=-1).type(scores.dtype)
        x = self._merge(torch.bmm(probs, v))
        x = self.proj_attn(x)
        # .contiguous(): numerics-free; torch 2.11's CPU group_norm forward-AD / backward reject
        # the channels-last-strided view that 0.11.0's `transpose(-1,-2).reshape(...)` yields.
        x = x.transpose(-1, -2).reshape(b, c, hh, ww).contiguous()
        return (x + residual) / self.rescale_output_factor


# --------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------
class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, eps, groups,
                 add_downsample, downsample_padding):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                          groups=groups, eps=eps) for i in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList(
                [Downsample2D(out_channels, out_channels, padding=downsample_padding)])

    def forward(self, hidden_states, temb=None):
        out = ()
        for r in self.resnets:
            hidden_states = r(hidden_states, temb)
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class AttnDownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, eps, groups,
                 attn_num_head_channels, add_downsample, downsample_padding):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                          groups=groups, eps=eps) for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            AttentionBlock(out_channels, attn_num_head_channels, groups=groups, eps=eps)
            for _ in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList(
                [Downsample2D(out_channels, out_channels, padding=downsample_padding)])

    def forward(self, hidden_states, temb=None):
        out = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = a(r(hidden_states, temb))
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, eps, groups,
                 heads, cross_attention_dim, add_downsample, downsample_padding,
                 use_linear_projection):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                          groups=groups, eps=eps) for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim,
                               groups=groups, use_linear_projection=use_linear_projection)
            for _ in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList(
                [Downsample2D(out_channels, out_channels, padding=downsample_padding)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None):
        out = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = r(hidden_states, temb)
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, eps, groups, heads, cross_attention_dim,
                 use_linear_projection):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels, in_channels, temb_channels, groups=groups, eps=eps),
            ResnetBlock2D(in_channels, in_channels, temb_channels, groups=groups, eps=eps)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, in_channels // heads, in_channels, cross_attention_dim,
                               groups=groups, use_linear_projection=use_linear_projection)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for a, r in zip(self.attentions, self.resnets[1:]):
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
            hidden_states = r(hidden_states, temb)
        return hidden_states


class UNetMidBlock2D(nn.Module):
    def __init__(self, in_channels, temb_channels, eps, groups, attn_num_head_channels):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels, in_channels, temb_channels, groups=groups, eps=eps),
            ResnetBlock2D(in_channels, in_channels, temb_channels, groups=groups, eps=eps)])
        self.attentions = nn.ModuleList([
            AttentionBlock(in_channels, attn_num_head_channels, groups=groups, eps=eps)])

    def forward(self, hidden_states, temb=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for a, r in zip(self.attentions, self.resnets[1:]):
            hidden_states = r(a(hidden_states), temb)
        return hidden_states


class UpBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers,
                 eps, groups, add_upsample):
        super().__init__()
        rs = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            rs.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, groups=groups, eps=eps))
        self.resnets = nn.ModuleList(rs)
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(out_channels, out_channels)])

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None):
        for r in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = r(torch.cat([hidden_states, res], dim=1), temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class AttnUpBlock2D(nn.Module):
    """diffusers 0.11.0 `AttnUpBlock2D` (UNet2DModel): [concat skip -> ResnetBlock2D -> AttentionBlock] x num_layers, then
    Upsample2D.  In-repo analogue: the `up` levels with `attn` of `src/models/ddpm/diffusion.py:96-118`."""
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, eps, groups,
                 attn_num_head_channels, add_upsample):
        super().__init__()
        rs = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            rs.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, groups=groups, eps=eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([
            AttentionBlock(out_channels, attn_num_head_channels, groups=groups, eps=eps) for _ in range(num_layers)])
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(out_channels, out_channels)])

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None):
        for r, a in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = a(r(torch.cat([hidden_states, res], dim=1), temb))
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers,
                 eps, groups, heads, cross_attention_dim, add_upsample, use_linear_projection):
        super().__init__()
        rs, at = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            rs.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, groups=groups, eps=eps))
            at.append(Transformer2DModel(heads, out_channels // heads, out_channels,
                                         cross_attention_dim, groups=groups,
                                         use_linear_projection=use_linear_projection))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(at)
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(out_channels, out_channels)])

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None,
                encoder_hidden_states=None, upsample_size=None):
        for r, a in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = r(torch.cat([hidden_states, res], dim=1), temb)
            hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


# --------------------------------------------------------------------------------------
# configs + top-level models
# --------------------------------------------------------------------------------------
@dataclass
class CondConfig:
    """UNet2DConditionModel config subset."""
    in_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)
    up_block_types: Tuple[str, ...] = ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3
    layers_per_block: int = 2
    attention_head_dim: object = 8          # = number of heads (int or per-block list)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    downsample_padding: int = 1
    use_linear_projection: bool = False
    sample_size: int = 64
    ctx_len: int = 77


@dataclass
class UncondConfig:
    """UNet2DModel config subset (google/ddpm-ema-celebahq-256)."""
    in_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 128, 256, 256, 512, 512)
    down_block_types: Tuple[str, ...] = ("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D")
    layers_per_block: int = 2
    attention_head_dim: Optional[int] = None
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    flip_sin_to_cos: bool = False
    freq_shift: int = 1
    downsample_padding: int = 0
    sample_size: int = 256


CONFIGS = {
    "sd15": CondConfig(),
    "sd21_768": CondConfig(attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                           use_linear_projection=True, sample_size=96),
    "sd21_base": CondConfig(attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                            use_linear_projection=True, sample_size=64),
    "celebahq": UncondConfig(),
    # reduced configs: same topology, small widths, CPU-oracle-in-seconds
    "sd_tiny": CondConfig(block_out_channels=(32, 64, 64, 64), attention_head_dim=2,
                          cross_attention_dim=32, sample_size=16, ctx_len=7),
    "sd_tiny_lin": CondConfig(block_out_channels=(32, 64, 64, 64), attention_head_dim=(1, 2, 2, 2),
                              cross_attention_dim=32, use_linear_projection=True, sample_size=16,
                              ctx_len=7),
    "sd_small": CondConfig(block_out_channels=(64, 128, 256, 256), attention_head_dim=4,
                           cross_attention_dim=64, sample_size=32, ctx_len=13),
    "uncond_tiny": UncondConfig(block_out_channels=(32, 32, 64, 64), sample_size=32,
                                down_block_types=("DownBlock2D", "DownBlock2D",
                                                  "AttnDownBlock2D", "DownBlock2D")),
}


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: CondConfig, build_up: bool = True):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        ted = boc[0] * 4
        heads = cfg.attention_head_dim
        if isinstance(heads, int):
            heads = (heads,) * len(boc)
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], ted)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            final = i == len(boc) - 1
            if t == "CrossAttnDownBlock2D":
                blk = CrossAttnDownBlock2D(in_ch, out_ch, ted, cfg.layers_per_block, cfg.norm_eps,
                                           cfg.norm_num_groups, heads[i], cfg.cross_attention_dim,
                                           not final, cfg.downsample_padding,
                                           cfg.use_linear_projection)
            elif t == "DownBlock2D":
                blk = DownBlock2D(in_ch, out_ch, ted, cfg.layers_per_block, cfg.norm_eps,
                                  cfg.norm_num_groups, not final, cfg.downsample_padding)
            else:
                raise ValueError(t)
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], ted, cfg.norm_eps, cfg.norm_num_groups,
                                                 heads[-1], cfg.cross_attention_dim,
                                                 cfg.use_linear_projection)
        self.up_blocks = nn.ModuleList()
        if build_up:
            rboc = list(reversed(boc))
            rheads = list(reversed(heads))
            out_ch = rboc[0]
            for i, t in enumerate(cfg.up_block_types):
                final = i == len(boc) - 1
                prev, out_ch = out_ch, rboc[i]
                in_ch = rboc[min(i + 1, len(boc) - 1)]
                if t == "UpBlock2D":
                    blk = UpBlock2D(in_ch, prev, out_ch, ted, cfg.layers_per_block + 1, cfg.norm_eps,
                                    cfg.norm_num_groups, not final)
                elif t == "CrossAttnUpBlock2D":
                    blk = CrossAttnUpBlock2D(in_ch, prev, out_ch, ted, cfg.layers_per_block + 1,
                                             cfg.norm_eps, cfg.norm_num_groups, rheads[i],
                                             cfg.cross_attention_dim, not final,
                                             cfg.use_linear_projection)
                else:
                    raise ValueError(t)
                self.up_blocks.append(blk)
            # diffusers 0.11.0 UNet2DConditionModel tail: GroupNorm -> SiLU -> Conv3x3(C0 -> out_channels = in_channels)
            self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
            self.conv_act = nn.SiLU()                                   # diffusers' attribute name (read by get_h_to_e, utils.py:631)
            self.conv_out = nn.Conv2d(boc[0], cfg.in_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states=None):
        """The full noise prediction eps(x_t, t, prompt) (diffusers 0.11.0 `UNet2DConditionModel.forward`, what the
        reference calls as `self.unet(latents, t, encoder_hidden_states=...).sample`, `edit.py:164-168`, `:458-462`)."""
        from oracle.pullback_oracle import get_h
        x = get_h(self, sample, timestep, encoder_hidden_states, "up", len(self.up_blocks) - 1)
        return self.conv_out(F.silu(self.conv_norm_out(x)))

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device


class UNet2DModel(nn.Module):
    def __init__(self, cfg: UncondConfig, build_up: bool = True):
        super().__init__()
        self.cfg = cfg
        boc = cfg.block_out_channels
        ted = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], ted)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            final = i == len(boc) - 1
            if t == "DownBlock2D":
                blk = DownBlock2D(in_ch, out_ch, ted, cfg.layers_per_block, cfg.norm_eps,
                                  cfg.norm_num_groups, not final, cfg.downsample_padding)
            elif t == "AttnDownBlock2D":
                blk = AttnDownBlock2D(in_ch, out_ch, ted, cfg.layers_per_block, cfg.norm_eps,
                                      cfg.norm_num_groups, cfg.attention_head_dim, not final,
                                      cfg.downsample_padding)
            else:
                raise ValueError(t)
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlock2D(boc[-1], ted, cfg.norm_eps, cfg.norm_num_groups,
                                        cfg.attention_head_dim)
        if build_up:
            # up path + eps head of diffusers 0.11.0 `UNet2DModel` (mirror of the down path: an Attn block where the down path has one)
            self.up_blocks = nn.ModuleList()
            rboc = list(reversed(boc))
            up_types = ["AttnUpBlock2D" if t == "AttnDownBlock2D" else "UpBlock2D" for t in reversed(cfg.down_block_types)]
            out_ch = rboc[0]
            for i, t in enumerate(up_types):
                final = i == len(boc) - 1
                prev, out_ch = out_ch, rboc[i]
                in_ch = rboc[min(i + 1, len(boc) - 1)]
                if t == "UpBlock2D":
                    blk = UpBlock2D(in_ch, prev, out_ch, ted, cfg.layers_per_block + 1, cfg.norm_eps, cfg.norm_num_groups, not final)
                else:
                    blk = AttnUpBlock2D(in_ch, prev, out_ch, ted, cfg.layers_per_block + 1, cfg.norm_eps, cfg.norm_num_groups,
                                        cfg.attention_head_dim, not final)
                self.up_blocks.append(blk)
            self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
            self.conv_out = nn.Conv2d(boc[0], cfg.in_channels, 3, padding=1)

    def forward(self, sample, timestep):
        """The full noise prediction of the unconditional model (diffusers 0.11.0 `UNet2DModel.forward`; the reference calls it
        as `self.unet(x, t).sample` in its uncond DDIM loops, `edit.py:1601-1714`)."""
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        emb = self.time_embedding(self.time_proj(t).to(dtype=self.dtype))
        x = self.conv_in(sample)
        skips = (x,)
        for blk in self.down_blocks:
            x, res = blk(hidden_states=x, temb=emb)
            skips += res
        x = self.mid_block(x, emb)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res, skips = skips[-n:], skips[:-n]
            x = blk(x, res, emb)
        return self.conv_out(F.silu(self.conv_norm_out(x)))

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (no checkpoints offline)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def seeded_init_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Order-independent deterministic init: every parameter is drawn from its own generator
    seeded by crc32(name) ^ seed.  Conv/Linear follow PyTorch's default scheme
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias); norm affine parameters are
    perturbed away from (1, 0) so parity tests exercise them."""
    for name, p in model.named_parameters():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        is_norm = ".norm" in name or name.startswith("norm") or "group_norm" in name or "norm_out" in name
        if is_norm:
            if name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            continue
        if name.endswith("weight"):
            fan_in = p[0].numel()
        else:
            w = dict(model.named_parameters())[name[:-4] + "weight"]
            fan_in = w[0].numel()
        bound = 1.0 / math.sqrt(fan_in)
        p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    return model


def build_unet(name: str, seed: int = 0, build_up: bool = True) -> nn.Module:
    cfg = CONFIGS[name]
    if isinstance(cfg, CondConfig):
        m = UNet2DConditionModel(cfg, build_up=build_up)
    else:
        m = UNet2DModel(cfg, build_up=build_up)
    seeded_init_(m, seed)
    return m.eval().requires_grad_(False)


def synthetic_inputs(name: str, seed: int = 1234):
    """x_t, t, ctx exactly as BASELINE.md section 4 / SURVEY.md section 8(d) specify."""
    cfg = CONFIGS[name]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g)
    t = torch.tensor(999.0 * 69.0 / 99.0)      # edit_t=0.7 -> idx 30 of linspace(0,1,100)*999
    ctx = None
    if isinstance(cfg, CondConfig):
        g2 = torch.Generator().manual_seed(4321)
        ctx = torch.randn(1, cfg.ctx_len, cfg.cross_attention_dim, generator=g2)
    return x, t, ctx
