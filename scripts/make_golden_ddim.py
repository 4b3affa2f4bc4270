"""Mint the golden DDIM trajectories: the reference's OWN scheduler functions
(`/root/reference/src/utils/utils.py:273-315` `set_timesteps` / `step`, `:1302-1317` `extract`), run verbatim, driving the
loops of `src/modules/edit.py:112-183` (inversion) and `:385-482` (sampling with classifier-free guidance) over the restated
diffusers-0.11.0 U-Net (oracle/unet_torch.py, full forward), torch-CPU fp32.  Authoring container only (needs /root/reference).

    python scripts/make_golden_ddim.py            -> tests/golden/ddim_sd_tiny.pt, tests/golden/ddim_sd_small.pt
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ddim_oracle as DO          # noqa: E402
from oracle import reference_shim as RS       # noqa: E402
from oracle import unet_torch as UT           # noqa: E402

U = RS.load()
for name, inv_steps, for_steps, t_end in (("sd_tiny", 8, 6, 4), ("sd_small", 6, 5, 3)):
    m = UT.build_unet(name)
    z0, _, ctx = UT.synthetic_inputs(name)
    neg = torch.randn(ctx.shape, generator=torch.Generator().manual_seed(9))
    sch = types.SimpleNamespace(t_max=999.0, alphas_cumprod=DO.sd_alphas_cumprod())
    sch.set_timesteps = types.MethodType(U.set_timesteps, sch)
    sch.step = types.MethodType(U.step, sch)
    scale = 2.5
    with torch.no_grad():
        # edit.py:131-178 (no guidance during inversion: guidance=None)
        sch.set_timesteps(inv_steps, is_inversion=True)
        lat = z0
        for i, t in enumerate(sch.timesteps):
            if i == len(sch.timesteps) - 1:
                break
            lat = sch.step(m(lat, t, encoder_hidden_states=ctx), t, lat, eta=0).prev_sample
        zT = lat
        # edit.py:404-470 with classifier-free guidance, stopped at t_end_idx
        sch.set_timesteps(for_steps)
        lat, t_edit, idx_edit = zT, None, None
        for t_idx, t in enumerate(sch.timesteps):
            if t_idx == t_end:
                t_edit, idx_edit = t, t_idx
                break
            e_un = m(lat, t, encoder_hidden_states=neg)
            e_c = m(lat, t, encoder_hidden_states=ctx)
            lat = sch.step(e_un + scale * (e_c - e_un), t, lat, eta=0).prev_sample
    out = dict(config=name, inv_steps=inv_steps, for_steps=for_steps, t_end_idx=t_end, guidance_scale=scale, neg=neg,
               zT=zT, z_edit=lat, t_edit=float(t_edit), idx_edit=idx_edit, eps0=m(z0, torch.tensor(999.0 * 69.0 / 99.0), encoder_hidden_states=ctx))
    path = os.path.join(ROOT, "tests", "golden", f"ddim_{name}.pt")
    torch.save(out, path)
    print(path, float(zT.std()), float(lat.std()), t_edit)
