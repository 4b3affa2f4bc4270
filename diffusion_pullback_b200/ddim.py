"""The reference's custom DDIM scheduler and its sampling / inversion loops on the B200 engine (first
"next" row): same names and call shapes as `/root/reference/src/utils/utils.py:273-315` (`set_timesteps`, `step`,
bound onto the scheduler with `types.MethodType` at `utils.py:340-342`) and `/root/reference/src/modules/edit.py`
`run_DDIMinversion` (`:112-183`) / `DDIMforwardsteps` (`:385-482`), minus the VAE and the CPU latent buffering.

    unet   = pb200.patch_unet(model.unet)           # also binds unet.eps(sample, timestep, encoder_hidden_states)
    sched  = pb200.DDIMSchedule(alphas_cumprod)     # stands where `self.scheduler` stands
    z_T    = pb200.ddim_inversion(unet, sched, z_0, inv_prompt_emb, inv_steps)
    z_t, t, t_idx = pb200.ddim_forward_steps(unet, sched, z_T, for_prompt_emb, for_steps, 0, t_edit_idx)

The U-Net runs as the engine's primal pass over the FULL plan (pb_plan op = PB_OP_FULL: every block + conv_norm_out +
SiLU + conv_out, hand-written sm_100a kernels) and the update is `pb_ddim_step`; no torch op touches the latents.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N


class SchedulerOutput:
    """`utils.py:1166-1169`: `.prev_sample` (x_next) and `.x0` (the predicted x_0)."""

    def __init__(self, xt_next, P_xt):
        self.prev_sample, self.x0 = xt_next, P_xt


class DDIMSchedule:
    """`alphas_cumprod` [T] (the diffusers scheduler's table) + the reference's float-timestep logic."""

    def __init__(self, alphas_cumprod, t_max: float = 999.0, _lib=None):
        self.alphas_cumprod = torch.as_tensor(alphas_cumprod, dtype=torch.float32).cpu()
        self.t_max = float(t_max)
        self.timesteps = self.timesteps_next = None
        self._L = _lib

    def set_timesteps(self, num_inferences, device=None, is_inversion=False):
        """`utils.py:273-286`: float timesteps linspace(0, 1, n) * t_max; inversion adds 1e-6 and walks upwards.  The body follows
        the reference's arithmetic step for step on purpose: the float schedule must be bit-identical (the golden test asserts
        `torch.equal` on it), so these ~8 lines are the one place the product restates reference code nearly line for line."""
        device = "cpu" if device is None else device
        seq = torch.linspace(0, 1, num_inferences, device=device) * self.t_max
        if is_inversion:
            seq = seq + 1e-6
            seq_prev = torch.cat([torch.tensor([-1], device=device), seq[:-1]], dim=0)
            self.timesteps, self.timesteps_next = seq_prev[1:], seq[1:]
        else:
            seq_prev = torch.cat([torch.tensor([-1], device=device), seq[:-1]], dim=0)
            self.timesteps, self.timesteps_next = reversed(seq[1:]), reversed(seq_prev[1:])

    def scale_model_input(self, sample, t):
        return sample

    def _alpha(self, t):
        # `extract` (utils.py:1302-1317): torch.gather(a, 0, t.long()) -- the float timestep is truncated
        return float(self.alphas_cumprod[int(torch.as_tensor(t).long())])

    def step(self, et, t, xt, eta=0.0, **kwargs):
        """`utils.py:288-315` / `:1202-1245`.  eta = 0 (the SD loops): one fused elementwise kernel (`pb_ddim_step`).
        eta != 0 (the stochastic branch the unconditional loop switches to under `performance_boosting`, `edit.py:1650-1653`):
        sigma_t = sqrt((1 - a_t / a_next)(1 - a_next) / (1 - a_t)),
        x_next = sqrt(a_next) P_xt + sqrt(1 - a_next - eta sigma_t^2) e_t + eta sigma_t z with z = torch.randn_like(x_t) -- the same
        draw from the same generator as the reference -- combined by `pb_lincomb3`."""
        if getattr(self, "learn_sigma", False):
            raise NotImplementedError("learn_sigma is the constant False in the reference (utils.py:1177): the learned-variance "
                                      "branch of YHCustomScheduler.step (utils.py:1239-1244) is unreachable there and not built")
        assert et.shape == xt.shape, "et, xt shape should be same"                     # utils.py:1210
        t_idx = self.timesteps.tolist().index(float(t))
        t_next = self.timesteps_next[t_idx]
        L = self._L if self._L is not None else N.lib()
        if xt.device.type != "cuda" and self._L is None:
            raise RuntimeError("diffusion_pullback_b200 runs on a CUDA (sm_100a) device only (no CPU fallback exists)")
        xt = xt.contiguous().float()
        et = et.contiguous().float()
        at, at_next = self._alpha(t), self._alpha(t_next)
        if eta == 0:
            x_next, p_xt = torch.empty_like(xt), torch.empty_like(xt)
            st = C.c_void_p(torch.cuda.current_stream(xt.device).cuda_stream) if xt.device.type == "cuda" else C.c_void_p(0)
            rc = L.pb_ddim_step(C.c_void_p(xt.data_ptr()), C.c_void_p(et.data_ptr()), at, at_next,
                                C.c_void_p(x_next.data_ptr()), C.c_void_p(p_xt.data_ptr()), xt.numel(), st)
            if rc != 0:
                raise ValueError("pb_ddim_step: invalid arguments")
            return SchedulerOutput(x_next, p_xt)
        import math
        sigma = math.sqrt((1.0 - at / at_next) * (1.0 - at_next) / (1.0 - at))
        d_coef = math.sqrt(1.0 - at_next - eta * sigma ** 2)
        z = torch.randn_like(xt)                                                        # utils.py:311 / :1237
        inv = 1.0 / math.sqrt(at)
        p_xt = _lincomb3(inv, xt, -math.sqrt(1.0 - at) * inv, et, 0.0, None, self._L)   # (x_t - sqrt(1 - a_t) e_t) / sqrt(a_t)
        x_next = _lincomb3(math.sqrt(at_next), p_xt, d_coef, et, eta * sigma, z, self._L)
        return SchedulerOutput(x_next, p_xt)


class YHCustomScheduler(DDIMSchedule):
    """`utils.py:1171-1286`: the scheduler of the unconditional (CelebA-HQ / DDPM) family -- same float-timestep
    `set_timesteps` / `step` as above over its own SNR schedule: 'linear' (betas = linspace(1e-4, 0.02, 1000) in float64,
    `:1248-1254`, `:1267-1268`) or 'cosine' (improved-DDPM, `:1275-1286`, t_max + 1 steps).  `args` is the reference's
    argument object (`noise_schedule`, `device`, `dtype` are read); keyword arguments do the same without one."""

    def __init__(self, args=None, noise_schedule=None, _lib=None):
        import math
        self.t_max = 999
        ns = getattr(args, "noise_schedule", None) if args is not None else noise_schedule
        self.noise_schedule = "linear" if ns is None else ns
        self.timesteps = self.timesteps_next = None
        self.learn_sigma = False
        self._L = _lib
        if self.noise_schedule == "linear":
            betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64)
        elif self.noise_schedule == "cosine":
            timesteps, sc = self.t_max + 1, 0.008
            x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
            ac = torch.cos(((x / timesteps) + sc) / (1 + sc) * math.pi * 0.5) ** 2
            ac = ac / ac[0]
            betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
        else:
            raise ValueError(f"unknown noise schedule {self.noise_schedule!r}")
        dtype = getattr(args, "dtype", torch.float32) if args is not None else torch.float32
        self.betas = betas.to(dtype=dtype).cpu()
        # float64 cumulative product, cast afterwards (utils.py:1262-1265); kept on the host: `step` reads two scalars per call
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).to(dtype=dtype).float().cpu()


def _eps(unet, latents, t, ctx, guidance_scale, neg_ctx):
    """Noise prediction with optional classifier-free guidance (`edit.py:150-175`, `:447-468`); the engine evaluates one
    latent at a time, so the guided pair is two primal passes."""
    if ctx is None:                                   # unconditional UNet2DModel: `self.unet(x, t).sample` (edit.py:1601-1714)
        return torch.cat([unet.eps(latents[i:i + 1], t) for i in range(latents.shape[0])], 0)
    outs = []
    for i in range(latents.shape[0]):
        x = latents[i:i + 1]
        c = ctx[i:i + 1] if ctx.shape[0] == latents.shape[0] else ctx[:1]
        if guidance_scale > 1.0 and neg_ctx is not None:
            n = neg_ctx[i:i + 1] if neg_ctx.shape[0] == latents.shape[0] else neg_ctx[:1]
            e_un, e_c = unet.eps(x, t, n), unet.eps(x, t, c)
            outs.append(e_un + guidance_scale * (e_c - e_un))
        else:
            outs.append(unet.eps(x, t, c))
    return torch.cat(outs, 0)


@torch.no_grad()
def ddim_inversion(unet, scheduler, z0, prompt_emb, num_inference_steps, guidance_scale=1.0, null_prompt_emb=None):
    """`run_DDIMinversion` (`edit.py:112-183`) from the latent on: z_0 -> z_T; the last timestep is skipped (`:151-152`)."""
    scheduler.set_timesteps(num_inference_steps, device="cpu", is_inversion=True)
    latents = z0
    for i, t in enumerate(scheduler.timesteps):
        if i == len(scheduler.timesteps) - 1:
            break
        noise_pred = _eps(unet, latents, t, prompt_emb, guidance_scale, null_prompt_emb)
        latents = scheduler.step(noise_pred, t, latents, eta=0).prev_sample
    return latents


@torch.no_grad()
def ddim_forward_steps(unet, scheduler, zt, prompt_emb, num_inference_steps, t_start_idx=0, t_end_idx=-1, guidance_scale=1.0,
                       neg_prompt_emb=None):
    """`DDIMforwardsteps` (`edit.py:385-482`) up to the VAE decode: returns `(latents, t, t_idx)` when `t_end_idx` is
    reached (`:432-434`), else the final latents."""
    scheduler.set_timesteps(num_inference_steps, device="cpu")
    latents = zt
    for t_idx, t in enumerate(scheduler.timesteps):
        if t_idx < t_start_idx:
            continue
        if t_idx == t_end_idx and t_idx != t_start_idx:
            return latents, t, t_idx
        noise_pred = _eps(unet, latents, t, prompt_emb, guidance_scale, neg_prompt_emb)
        latents = scheduler.step(noise_pred, t, latents, eta=0).prev_sample
    return latents


@torch.no_grad()
def ddim_forward_steps_uncond(unet, scheduler, xt, num_inference_steps, t_start_idx=0, t_end_idx=-1, performance_boosting=False,
                              performance_boosting_t_idx=None):
    """`EditUncondDiffusion.DDIMforwardsteps` (`edit.py:1601-1714`) up to the image save.  Differences from the SD loop that
    `ddim_forward_steps` follows: the end test comes BEFORE the skip test (`:1638-1645`: `t_end_idx == t_start_idx` returns at
    once), there is no prompt, and with `performance_boosting` the steps from `performance_boosting_t_idx` on (unless it is the
    last index) run the stochastic branch, eta = 1 (`:1650-1653`).  A batch is evaluated one image at a time."""
    scheduler.set_timesteps(num_inference_steps, device="cpu")
    timesteps = scheduler.timesteps
    assert (t_start_idx < num_inference_steps) and (t_end_idx <= num_inference_steps)            # edit.py:1617
    for i, t in enumerate(timesteps):
        if t_end_idx == i:
            return xt, t, i
        elif i < t_start_idx:
            continue
        boost = bool(performance_boosting) and performance_boosting_t_idx is not None and \
            (performance_boosting_t_idx <= i) and (performance_boosting_t_idx != len(timesteps) - 1)
        eta = 1 if boost else 0
        et = unet.eps(xt, t)
        xt = scheduler.step(et, t, xt, eta=eta).prev_sample
    return xt


def _lincomb3(a, x, b, y, c, z, _lib=None):
    """a x + b y + c z through `pb_lincomb3` (one fused elementwise kernel on the tensors' device)."""
    L = _lib if _lib is not None else N.lib()
    if x.device.type != "cuda" and _lib is None:
        raise RuntimeError("diffusion_pullback_b200 runs on a CUDA (sm_100a) device only (no CPU fallback exists)")
    x = x.contiguous().float()
    y = y.contiguous().float() if y is not None else None
    z = z.contiguous().float() if z is not None else None
    out = torch.empty_like(x)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream) if x.device.type == "cuda" else C.c_void_p(0)
    if L.pb_lincomb3(p(out), float(a), p(x), float(b), p(y), float(c), p(z), x.numel(), st) != 0:
        raise ValueError("pb_lincomb3: invalid arguments")
    return out


@torch.no_grad()
def x_space_guidance(unet, scheduler, zt, t_idx, vk, single_edit_step, edit_prompt_emb, x_space_guidance_scale, _lib=None):
    """`EditStableDiffusion.x_space_guidance` (`edit.py:484-502`): perturb z_t along the pullback direction v_k, predict the
    noise at both points with the edit prompt, and move z_t by the scaled difference (DDS regularisation).
    `edit_prompt_emb=None`: the unconditional variant, `EditUncondDiffusion.x_space_guidance` (`edit.py:1716-1734`)."""
    t = scheduler.timesteps[t_idx]
    zt_edit = _lincomb3(1.0, zt, single_edit_step, vk.reshape(zt.shape), 0.0, None, _lib)          # zt + step * vk
    both = torch.cat([zt, zt_edit], dim=0)                     # ONE batch-2 U-Net call like the reference (edit.py:492-497)
    if edit_prompt_emb is None:
        et = unet.eps(both, t)
    else:
        et = unet.eps(both, t, edit_prompt_emb.repeat(both.shape[0] // edit_prompt_emb.shape[0], 1, 1))
    et_null, et_edit = et.chunk(2)
    return _lincomb3(1.0, zt, x_space_guidance_scale, et_edit, -x_space_guidance_scale, et_null, _lib)


@torch.no_grad()
def x_space_guidance_edit(unet, scheduler, zt, t_idx, vk, num_step, single_edit_step, edit_prompt_emb, x_space_guidance_scale,
                          _lib=None):
    """The edit loop of `edit.py:290-301`: `num_step` guidance steps from z_t; returns the list [z_t, z_t^1, ...]."""
    zt_list = [zt.clone()]
    for _ in range(num_step):
        zt_list.append(x_space_guidance(unet, scheduler, zt_list[-1], t_idx, vk, single_edit_step, edit_prompt_emb,
                                        x_space_guidance_scale, _lib))
    return zt_list
