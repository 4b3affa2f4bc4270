"""The product-side synthetic workload (diffusion_pullback_b200/synthetic.py) must be the SAME workload the
oracle runs: identical weights (per-name seeded draws) and identical (x_t, t, ctx)."""
import pytest
import torch

from diffusion_pullback_b200 import synthetic as SY
from oracle import unet_torch as UT


@pytest.mark.parametrize("name", ["sd_tiny", "sd_tiny_lin", "sd_small", "uncond_tiny"])
def test_state_dict_matches_oracle_module_tree(name):
    ref = UT.build_unet(name).state_dict()
    sd = SY.SyntheticUNet(name).state_dict()
    assert set(sd) == set(ref), (set(sd) ^ set(ref))
    for k, v in sd.items():
        assert v.numel() == ref[k].numel() and torch.equal(v.reshape(ref[k].shape), ref[k]), k
    mid = SY.SyntheticUNet(name, upto=("mid", 0)).state_dict()
    assert set(mid) < set(sd) or name == "uncond_tiny"
    assert not any(k.startswith("up_blocks") for k in mid)


@pytest.mark.parametrize("name", ["sd15", "sd21_768", "celebahq", "sd_tiny"])
def test_inputs_and_config_match_oracle(name):
    x, t, c = SY.synthetic_inputs(name)
    xr, tr, cr = UT.synthetic_inputs(name)
    assert torch.equal(x, xr) and float(t) == float(tr)
    assert (c is None and cr is None) or torch.equal(c, cr)
    from diffusion_pullback_b200.engine import unet_config
    a = unet_config(SY.SyntheticUNet(name))
    m = UT.UNet2DConditionModel.__new__(UT.UNet2DConditionModel) if False else None
    cfg = UT.CONFIGS[name]
    assert a["block_out_channels"] == list(cfg.block_out_channels) and a["norm_eps"] == cfg.norm_eps
    assert a["in_channels"] == cfg.in_channels and a["downsample_padding"] == cfg.downsample_padding
