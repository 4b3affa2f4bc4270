"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

CPU restatement of the reference's custom DDIM scheduler and sampling / inversion loops (the row SURVEY.md s.8f ranks
first among the "next" rows).  Each function cites the reference lines it follows; `tests/test_oracle.py` pins
`set_timesteps`, `step` and `extract` against the reference's own functions run verbatim (oracle/reference_shim.py).
Only tests/, smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import torch


def sd_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
    """`scaled_linear` betas of the Stable-Diffusion DDIMScheduler the reference reads `alphas_cumprod` from
    (`utils.py:319-345`): linspace over sqrt(beta), squared."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def extract(a, t, x_shape):
    """`utils.py:1302-1317`: gather a[t.long()] per batch element, broadcastable to x_shape."""
    if isinstance(t, int):
        t = torch.tensor([t]).repeat(x_shape[0])
    elif isinstance(t, torch.Tensor):
        t = t.repeat(x_shape[0])
    else:
        raise ValueError(f"t must be int or torch.Tensor, got {type(t)}")
    out = torch.gather(a, 0, t.long())
    return out.reshape((x_shape[0],) + (1,) * (len(x_shape) - 1))


class Scheduler:
    """State the reference's scheduler methods read: t_max, alphas_cumprod, timesteps, timesteps_next."""

    def __init__(self, alphas_cumprod, t_max: float = 999.0):
        self.alphas_cumprod, self.t_max = alphas_cumprod, t_max
        self.timesteps = self.timesteps_next = None

    def set_timesteps(self, num_inferences, device=None, is_inversion=False):
        """`utils.py:273-286`: float timesteps linspace(0, 1, n) * t_max; inversion walks them upwards (+1e-6)."""
        device = "cpu" if device is None else device
        seq = torch.linspace(0, 1, num_inferences, device=device) * self.t_max
        if is_inversion:
            seq = seq + 1e-6
            seq_prev = torch.cat([torch.tensor([-1], device=device), seq[:-1]], dim=0)
            self.timesteps, self.timesteps_next = seq_prev[1:], seq[1:]
        else:
            seq_prev = torch.cat([torch.tensor([-1], device=device), seq[:-1]], dim=0)
            self.timesteps, self.timesteps_next = reversed(seq[1:]), reversed(seq_prev[1:])

    def step(self, et, t, xt, eta=0.0):
        """`utils.py:288-315` (= `YHCustomScheduler.step`, `:1202-1237`, without learned variances): returns (x_next, pred_x0).
        eta = 0: the deterministic DDIM update; eta != 0: the stochastic branch (one `torch.randn_like(xt)` draw)."""
        t_idx = self.timesteps.tolist().index(t)
        t_next = self.timesteps_next[t_idx]
        at = extract(self.alphas_cumprod, t, xt.shape)
        at_next = extract(self.alphas_cumprod, t_next, xt.shape)
        p_xt = (xt - et * (1 - at).sqrt()) / at.sqrt()
        if eta == 0:
            return at_next.sqrt() * p_xt + (1 - at_next).sqrt() * et, p_xt
        sigma_t = ((1 - at / (at_next)) * (1 - at_next) / (1 - at)).sqrt()
        d_xt = (1 - at_next - eta * sigma_t ** 2).sqrt() * et
        return at_next.sqrt() * p_xt + d_xt + eta * sigma_t * torch.randn_like(xt), p_xt


def yh_schedule(noise_schedule="linear", t_max=999, dtype=torch.float32):
    """`YHCustomScheduler.get_alphas_cumprod` (`utils.py:1247-1286`): (betas, alphas_cumprod) of the unconditional family --
    'linear': linspace(1e-4, 0.02, 1000) in float64; 'cosine': improved-DDPM over t_max + 1 steps; cumulative product in
    float64, cast to `dtype` afterwards."""
    import math
    if noise_schedule == "linear":
        betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64)
    else:
        timesteps, sc = t_max + 1, 0.008
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
        ac = torch.cos(((x / timesteps) + sc) / (1 + sc) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    return betas.to(dtype), torch.cumprod(1.0 - betas, dim=0).to(dtype)


def _eps(unet, latents, t, ctx, guidance_scale, neg_ctx):
    """noise prediction with optional classifier-free guidance (`edit.py:150-175`, `:447-468`)."""
    if ctx is None:
        return unet(latents, t)
    if guidance_scale > 1.0 and neg_ctx is not None:
        e_un = unet(latents, t, encoder_hidden_states=neg_ctx)
        e_c = unet(latents, t, encoder_hidden_states=ctx)
        return e_un + guidance_scale * (e_c - e_un)
    return unet(latents, t, encoder_hidden_states=ctx)


@torch.no_grad()
def ddim_inversion(unet, sched: Scheduler, z0, ctx, num_steps, guidance_scale=1.0, neg_ctx=None):
    """`edit.py:112-183` without the VAE: z0 -> z_T along the float inversion schedule; the last timestep is skipped."""
    sched.set_timesteps(num_steps, is_inversion=True)
    latents = z0
    for i, t in enumerate(sched.timesteps):
        if i == len(sched.timesteps) - 1:
            break
        latents = sched.step(_eps(unet, latents, t, ctx, guidance_scale, neg_ctx), t, latents)[0]
    return latents


@torch.no_grad()
def ddim_forward_steps(unet, sched: Scheduler, zt, ctx, num_steps, t_start_idx=0, t_end_idx=-1, guidance_scale=1.0, neg_ctx=None):
    """`edit.py:385-482` without the VAE / buffering: denoise from schedule index t_start_idx; returns
    (latents, t, t_idx) when t_end_idx is reached, else the final latents."""
    sched.set_timesteps(num_steps)
    latents = zt
    for t_idx, t in enumerate(sched.timesteps):
        if t_idx < t_start_idx:
            continue
        if t_idx == t_end_idx and t_idx != t_start_idx:
            return latents, t, t_idx
        latents = sched.step(_eps(unet, latents, t, ctx, guidance_scale, neg_ctx), t, latents)[0]
    return latents


@torch.no_grad()
def ddim_forward_steps_uncond(unet, sched: Scheduler, xt, num_steps, t_start_idx=0, t_end_idx=-1, performance_boosting=False,
                              performance_boosting_t_idx=None):
    """`EditUncondDiffusion.DDIMforwardsteps` (`edit.py:1601-1714`) without buffering / image saving: the end test precedes the
    skip test (`:1638-1645`); eta = 1 from `performance_boosting_t_idx` on under `performance_boosting` (`:1650-1653`)."""
    sched.set_timesteps(num_steps)
    timesteps = sched.timesteps
    for i, t in enumerate(timesteps):
        if t_end_idx == i:
            return xt, t, i
        elif i < t_start_idx:
            continue
        if performance_boosting and (performance_boosting_t_idx <= i) and (performance_boosting_t_idx != len(timesteps) - 1):
            eta = 1
        else:
            eta = 0
        xt = sched.step(unet(xt, t), t, xt, eta=eta)[0]
    return xt


@torch.no_grad()
def x_space_guidance(unet, sched: Scheduler, zt, t_idx, vk, single_edit_step, edit_prompt_emb, scale):
    """`edit.py:484-502`: one batch-2 U-Net call on [zt, zt + step * vk] with the edit prompt, then
    zt + scale * (eps_edit - eps_null)."""
    t = sched.timesteps[t_idx]
    zt_edit = zt + single_edit_step * vk
    if edit_prompt_emb is None:                       # `EditUncondDiffusion.x_space_guidance`, edit.py:1716-1734
        et = unet(torch.cat([zt, zt_edit], dim=0), t)
    else:
        et = unet(torch.cat([zt, zt_edit], dim=0), t, encoder_hidden_states=edit_prompt_emb.repeat(2, 1, 1))
    et_null, et_edit = et.chunk(2)
    return zt + scale * (et_edit - et_null)
