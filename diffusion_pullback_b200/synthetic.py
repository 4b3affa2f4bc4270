"""Synthetic workloads for benchmarks and smoke runs: the reference's model configurations as plain dicts
(no checkpoints or network offline: BASELINE.json asks for random-init weights + synthetic latents), a
deterministic state_dict generator keyed by parameter NAME, and the synthetic (x_t, t, ctx) of SURVEY.md
s.8(d).  The draw per parameter is a pure function of (name, shape, seed), so the CPU oracle's module
tree (oracle/unet_torch.py, same rule) and this product-side generator hold identical weights without
sharing code; tests/test_synthetic.py pins the two against each other."""
from __future__ import annotations

import math
import zlib

import torch

_SD_DOWN = ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]
_SD_UP = ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3


def _sd(boc=(320, 640, 1280, 1280), heads=8, ctx=768, size=64, ctx_len=77):
    return dict(in_channels=4, block_out_channels=list(boc), down_block_types=_SD_DOWN, up_block_types=_SD_UP,
                layers_per_block=2, attention_head_dim=heads, cross_attention_dim=ctx, norm_num_groups=32, norm_eps=1e-5,
                flip_sin_to_cos=True, freq_shift=0, downsample_padding=1, sample_size=size, ctx_len=ctx_len)


def _uncond(boc=(128, 128, 256, 256, 512, 512), size=256, down=None):
    return dict(in_channels=3, block_out_channels=list(boc),
                down_block_types=down or ["DownBlock2D"] * 4 + ["AttnDownBlock2D", "DownBlock2D"], up_block_types=[],
                layers_per_block=2, attention_head_dim=None, norm_num_groups=32, norm_eps=1e-6, flip_sin_to_cos=False,
                freq_shift=1, downsample_padding=0, sample_size=size, ctx_len=0)


CONFIGS = {
    "sd15": _sd(),                                                        # runwayml/stable-diffusion-v1-5
    "sd21_768": _sd(heads=[5, 10, 20, 20], ctx=1024, size=96),             # stabilityai/stable-diffusion-2-1
    "sd21_base": _sd(heads=[5, 10, 20, 20], ctx=1024, size=64),
    "celebahq": _uncond(),                                                 # google/ddpm-ema-celebahq-256
    "sd_tiny": _sd((32, 64, 64, 64), 2, 32, 16, 7),
    "sd_tiny_lin": _sd((32, 64, 64, 64), [1, 2, 2, 2], 32, 16, 7),
    "sd_small": _sd((64, 128, 256, 256), 4, 64, 32, 13),
    "uncond_tiny": _uncond((32, 32, 64, 64), 32, ["DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D"]),
}


class SyntheticUNet:
    """Stands where the diffusers `unet` object stands in the reference: `.config`, `.state_dict()`, and
    (for conditional models) an `.up_blocks` attribute so `patch_unet` picks the SD methods."""

    def __init__(self, name_or_cfg, seed: int = 0, device="cpu", upto=None):
        """`upto=(op, block_idx)` materialises only the weights the truncated forward up to that point reads."""
        self.config = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
        self.seed, self.device, self._sd, self.upto = seed, torch.device(device), None, upto
        if self.config["up_block_types"]:
            self.up_blocks = tuple(self.config["up_block_types"])

    def state_dict(self):
        if self._sd is None:
            from .engine import PullbackEngine, unet_config
            cfg = unet_config(self)
            op, bi = ("full", 0)
            if self.upto is not None:
                op, bi = self.upto
            s = self.config["sample_size"]
            specs = _weight_specs(cfg, s, op, bi, max(1, self.config["ctx_len"]))
            self._sd = synthetic_state_dict(specs, self.seed, self.device)
        return self._sd


def _weight_specs(cfg, size, op, bi, ctx_len):
    from . import _native as N
    import ctypes as C
    from .engine import PB_OP
    L = N.lib()
    c = N.PbUnetCfg()
    c.kind, c.in_channels, c.n_levels = cfg["kind"], cfg["in_channels"], len(cfg["block_out_channels"])
    for i, v in enumerate(cfg["block_out_channels"]):
        c.block_out_channels[i], c.down_has_attn[i] = v, cfg["down_has_attn"][i]
        c.up_has_attn[i], c.heads[i] = cfg["up_has_attn"][i], cfg["heads"][i]
    c.layers_per_block, c.cross_attention_dim = cfg["layers_per_block"], cfg["cross_attention_dim"]
    c.norm_num_groups, c.norm_eps = cfg["norm_num_groups"], cfg["norm_eps"]
    c.flip_sin_to_cos, c.freq_shift, c.downsample_padding = cfg["flip_sin_to_cos"], cfg["freq_shift"], cfg["downsample_padding"]
    h = C.c_void_p()
    if L.pb_create(C.byref(c), C.byref(h)) != 0:
        raise ValueError("invalid U-Net configuration")
    try:
        if L.pb_plan(h, size, size, PB_OP[op], bi, 1, ctx_len, None) != 0:
            raise ValueError(L.pb_last_error(h).decode())
        out = []
        for i in range(L.pb_weight_count(h)):
            name, nd, shp = C.c_char_p(), C.c_int32(), (C.c_int64 * 4)()
            L.pb_weight_info(h, i, C.byref(name), C.byref(nd), shp)
            out.append((name.value.decode(), tuple(int(shp[j]) for j in range(nd.value))))
        return out
    finally:
        L.pb_destroy(h)


def synthetic_state_dict(specs, seed: int = 0, device="cpu"):
    """One generator per parameter, seeded by crc32(name) ^ seed: Conv/Linear follow PyTorch's default
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias; norm affine parameters are 1 + 0.1 N(0,1) / 0.1 N(0,1)."""
    shapes = dict(specs)
    sd = {}
    for name, shape in specs:
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        if ".norm" in name or name.startswith("norm") or "group_norm" in name or "norm_out" in name:
            p = 1.0 + 0.1 * torch.randn(shape, generator=g) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = shape if name.endswith("weight") else shapes[name[:-4] + "weight"]
            fan_in = math.prod(wshape[1:])
            p = (torch.rand(shape, generator=g) * 2 - 1) * (1.0 / math.sqrt(fan_in))
        sd[name] = p.to(device)
    return sd


def synthetic_inputs(name_or_cfg, seed: int = 1234, device="cpu"):
    """x_t ~ N(0,1) seed 1234; t = 999*69/99 (edit_t = 0.7 -> index 30 of the reference's 100-step float schedule,
    utils.py:283-285); ctx ~ N(0,1) seed 4321 standing in for CLIP embeddings."""
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, cfg["in_channels"], cfg["sample_size"], cfg["sample_size"], generator=g)
    t = torch.tensor(999.0 * 69.0 / 99.0)
    ctx = None
    if cfg["up_block_types"]:
        g2 = torch.Generator().manual_seed(4321)
        ctx = torch.randn(1, cfg["ctx_len"], cfg["cross_attention_dim"], generator=g2)
    return x.to(device), t, (ctx.to(device) if ctx is not None else None)
