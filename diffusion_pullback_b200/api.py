"""Drop-in replacements for the reference's monkey-patched U-Net methods
(`/root/reference/src/utils/utils.py`): same names, same positional/keyword signatures, same return
contract, bindable with `types.MethodType` exactly like `utils.py:103-104`, `:326`, `:333` do:

    unet.get_h                     <- get_h / get_h_uncond          (utils.py:438-527 / :114-163)
    unet.local_encoder_pullback_zt <- local_encoder_pullback_zt     (utils.py:722-816)
    unet.local_encoder_pullback_xt <- local_encoder_pullback_xt     (utils.py:165-249)

`patch_unet(unet)` performs the three bindings.  All arithmetic runs in libpullback_b200.so (sm_100a
CUDA); there is no PyTorch / CPU fallback -- a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import time
import types

import torch

from .engine import PullbackEngine, unet_config


def _engine_for(unet, sample, op, block_idx, k, ctx_len) -> PullbackEngine:
    if sample.device.type != "cuda":
        raise RuntimeError("diffusion_pullback_b200 runs on a CUDA (sm_100a) device only; got a "
                           f"{sample.device.type} tensor (no CPU fallback exists)")
    cache = unet.__dict__.setdefault("_pb200_engines", {})
    key = (sample.device.index, sample.shape[2], sample.shape[3], op, block_idx, ctx_len)
    eng = cache.get(key)
    if eng is None or eng.k_max < k:
        cfg = unet_config(unet)
        if sample.shape[1] != cfg["in_channels"]:
            raise ValueError("sample has the wrong number of channels")
        eng = PullbackEngine(cfg, sample.shape[2], sample.shape[3], op, block_idx, max(k, 1), ctx_len, sample.device)
        eng.bind(unet.state_dict())
        cache[key] = eng
    return eng


def refresh_weights(unet):
    """Re-pack weights after the module's parameters changed (engines cache a packed copy)."""
    for eng in unet.__dict__.get("_pb200_engines", {}).values():
        eng.bind(unet.state_dict())


def _timestep(t, i=0):
    """Sample i's timestep: a scalar / 0-dim / 1-element value is shared by the batch (the reference expands it,
    utils.py:461-468); a tensor with one entry per sample is indexed."""
    if torch.is_tensor(t):
        t = t.reshape(-1)
        return float(t[i] if t.numel() > 1 else t[0])
    return float(t)


def get_h(self, sample=None, timestep=None, encoder_hidden_states=None, op=None, block_idx=None, verbose=False):
    """`utils.py:438-527`: the truncated U-Net forward.  A batch (the reference expands the timestep over it,
    utils.py:461-468) is evaluated one latent at a time on the same planned engine."""
    if sample.shape[0] != 1:
        ehs = encoder_hidden_states
        return torch.cat([get_h(self, sample[i:i + 1], _timestep(timestep, i), None if ehs is None else ehs[i:i + 1], op, block_idx)
                          for i in range(sample.shape[0])], 0)
    ctx_len = encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0
    eng = _engine_for(self, sample, op, block_idx, 1, ctx_len)
    return eng.set_point(sample, _timestep(timestep), encoder_hidden_states, want_h=True)


def eps(self, sample, timestep, encoder_hidden_states=None):
    """The whole U-Net, x_t -> noise prediction: what the reference calls as `self.unet(latents, t,
    encoder_hidden_states=prompt_emb).sample` in its DDIM loops (`edit.py:164-168`, `:458-462`); one latent per call."""
    if sample.shape[0] != 1:
        ehs = encoder_hidden_states
        return torch.cat([eps(self, sample[i:i + 1], _timestep(timestep, i), None if ehs is None else ehs[i:i + 1])
                          for i in range(sample.shape[0])], 0)
    ctx_len = encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0
    eng = _engine_for(self, sample, "full", 0, 1, ctx_len)
    return eng.set_point(sample, _timestep(timestep), encoder_hidden_states, want_h=True)


def eps_uncond(self, x, t):
    """The whole unconditional U-Net (`UNet2DModel`), x_t -> noise prediction: `self.unet(x, t).sample` of the reference's
    uncond DDIM loops (`edit.py:1601-1714`); one image per call."""
    if x.shape[0] != 1:
        return torch.cat([eps_uncond(self, x[i:i + 1], _timestep(t, i)) for i in range(x.shape[0])], 0)
    eng = _engine_for(self, x, "full", 0, 1, 0)
    return eng.set_point(x, _timestep(t), None, want_h=True)


def get_h_uncond(self, x=None, t=None, op=None, block_idx=None, verbose=False):
    """`utils.py:114-163`; only ('mid', 0) is valid, anything else raises ValueError like the reference."""
    if x.shape[0] != 1:
        return torch.cat([get_h_uncond(self, x[i:i + 1], _timestep(t, i), op, block_idx) for i in range(x.shape[0])], 0)
    eng = _engine_for(self, x, op, block_idx, 1, 0)
    return eng.set_point(x, _timestep(t), None, want_h=True)


def _pullback(self, sample, timestep, ctx, op, block_idx, pca_rank, min_iter, max_iter, convergence_threshold, v0, return_info,
              tangent_group=None):
    time_s = time.time()
    k = int(pca_rank)
    ctx_len = ctx.shape[1] if ctx is not None else 0
    eng = _engine_for(self, sample, op, block_idx, k, ctx_len)
    n_in = sample[0].numel()
    if v0 is None:
        # identical RNG consumption and V0 to the reference (utils.py:750-752 / :193-195)
        vT = torch.randn(n_in, k, device=sample.device, dtype=torch.float)
        vT, _ = torch.linalg.qr(vT)
        v0 = vT.T.contiguous()
    eng.set_point(sample, _timestep(timestep), ctx)
    if tangent_group is not None:
        # one problem, k columns split over the ranks of the group (sharding.pullback_tangent_sharded); V0 of the
        # group's first rank is used by everyone
        import torch.distributed as dist
        from .sharding import pullback_tangent_sharded
        v0 = v0.to(sample.device, torch.float32).contiguous()
        dist.broadcast(v0, src=dist.get_global_rank(tangent_group, 0), group=tangent_group)
        u, s, vT, info = pullback_tangent_sharded(eng, v0, min_iter, max_iter, convergence_threshold, tangent_group)
    else:
        u, s, vT, info = eng.pullback(v0, min_iter, max_iter, convergence_threshold)
    print(f"power method : {info.iters_done - 1}-th step convergence : ", info.last_dist)
    if info.converged:
        print("reach convergence threshold : ", info.last_dist)
    print("power method runtime ==", time.time() - time_s)
    u = u.T                                    # [n_out, k] view, like the reference's `.view(k, n_out).T`
    if return_info:
        return u, s, vT, dict(iters_done=info.iters_done, converged=bool(info.converged), last_dist=info.last_dist)
    return u, s, vT


@torch.no_grad()
def local_encoder_pullback_zt(self, sample, timestep, encoder_hidden_states=None, op=None, block_idx=None,
                              pca_rank=50, chunk_size=25, min_iter=10, max_iter=100, convergence_threshold=1e-3,
                              v0=None, return_info=False, tangent_group=None):
    """`utils.py:722-816`.  `chunk_size` bounds memory in the reference and does not change results; the
    k tangents always ride the batch axis of one pass here.  Returns (u [n_out,k], s [k], vT [k,n_in]).
    Extensions (defaults keep the reference behaviour): `v0` start subspace [k, n_in]; `return_info`;
    `tangent_group`: a torch.distributed group whose ranks all make this call on the same problem and split its
    k columns (one all-gather of W per iteration)."""
    return _pullback(self, sample, timestep, encoder_hidden_states, op, block_idx, pca_rank, min_iter, max_iter,
                     convergence_threshold, v0, return_info, tangent_group)


@torch.no_grad()
def local_encoder_pullback_xt(self, x, t, op=None, block_idx=None, pca_rank=50, chunk_size=25, min_iter=10, max_iter=100,
                              convergence_threshold=1e-3, v0=None, return_info=False, tangent_group=None):
    """`utils.py:165-249` (unconditional `UNet2DModel`)."""
    return _pullback(self, x, t, None, op, block_idx, pca_rank, min_iter, max_iter, convergence_threshold, v0, return_info,
                     tangent_group)


@torch.no_grad()
def local_encoder_pullback_many(self, samples, timesteps, encoder_hidden_states=None, op=None, block_idx=None, pca_rank=50,
                                min_iter=10, max_iter=100, convergence_threshold=1e-3, v0=None):
    """Extension (throughput mode, no counterpart in the reference, which solves its problems one process after another --
    `src/scripts/*.sh`, `main.py:71-76`): P independent problems of the same geometry in one call.  `samples` [P, C, H, W],
    `timesteps` P values, `encoder_hidden_states` [P, L, D] (or None for the unconditional family), `v0` optional
    [P, k, n_in].  The problems share the packed weights and run their k tangent columns each as one batch of P * k images
    (pb_set_slots): weight GEMMs and data movement in one launch over all problems, primal-dependent ops once per problem.
    Every problem runs the same number of iterations (the early exit needs all of them converged at the same check).
    Returns a list of P tuples (u [n_out, k], s [k], vT [k, n_in])."""
    P, k = samples.shape[0], int(pca_rank)
    ctx = encoder_hidden_states
    ctx_len = ctx.shape[1] if ctx is not None else 0
    if samples.device.type != "cuda":
        raise RuntimeError("diffusion_pullback_b200 runs on a CUDA (sm_100a) device only (no CPU fallback exists)")
    if P * k > 64:
        raise ValueError("P * pca_rank must be <= 64 (k_max of one handle)")
    engines = self.__dict__.setdefault("_pb200_slot_engines", {})
    key = (samples.device.index, samples.shape[2], samples.shape[3], op, block_idx, ctx_len, P, k)
    eng = engines.get(key)
    if eng is None:
        eng = PullbackEngine(unet_config(self), samples.shape[2], samples.shape[3], op, block_idx, P * k, ctx_len, samples.device)
        eng.bind(self.state_dict())
        eng.set_slots(P)
        engines[key] = eng
    return _pullback_many_on(eng, samples, timesteps, ctx, k, min_iter, max_iter, convergence_threshold, v0)


def _pullback_many_on(eng, samples, timesteps, ctx, k, min_iter, max_iter, convergence_threshold, v0=None):
    """The body of `local_encoder_pullback_many` on an engine whose slots are set (one per sample)."""
    P, n_in = samples.shape[0], samples[0].numel()
    if v0 is None:
        v0 = []
        for _ in range(P):                       # per problem the reference's own start (utils.py:750-752)
            q, _ = torch.linalg.qr(torch.randn(n_in, k, device=samples.device, dtype=torch.float))
            v0.append(q.T)
        v0 = torch.stack(v0, 0)
    for p in range(P):
        eng.set_point(samples[p:p + 1], _timestep(timesteps[p]), None if ctx is None else ctx[p:p + 1], slot=p)
    u, s, vT, info = eng.pullback(v0.reshape(P * k, n_in), min_iter, max_iter, convergence_threshold)
    return [(u[p * k:(p + 1) * k].T, s[p * k:(p + 1) * k], vT[p * k:(p + 1) * k]) for p in range(P)]


# ---- decoder side: get_h_to_e, local_decoder_pullback_zt, inv_jac_zt ----
def get_h_to_e(self, sample=None, timestep=None, encoder_hidden_states=None, input_h=None, op=None, block_idx=None, verbose=False):
    """`utils.py:529-635`: the decoder half of the U-Net from a substituted mid-block feature -- every row of `input_h`
    ([pca_rank, C, H, W] or flattened) is pushed through the up path, `conv_norm_out`, SiLU and `conv_out` on the skip connections
    of `sample` (one latent); returns the noise predictions [pca_rank, c, H, W].  Like the reference, only ('mid', 0) is a valid
    place to substitute h (`assert op in ['mid', 'down']`, and 'down' never substitutes)."""
    if not (op == "mid" and block_idx == 0):
        raise AssertionError("up block is not implemented yet" if op == "up" else f"(op, block_idx) = ({op, block_idx}) is not valid")
    if sample.shape[0] != 1:
        raise ValueError("get_h_to_e takes one latent (the reference repeats its skip connections pca_rank times)")
    ctx_len = encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0
    eng = _engine_for(self, sample, "dec", 0, 1, ctx_len)
    input_h = input_h.reshape(-1, *eng.in_shape)
    eng.set_point(sample, _timestep(timestep), encoder_hidden_states)
    return torch.cat([eng.decode_from(input_h[i:i + 1]) for i in range(input_h.shape[0])], 0)


@torch.no_grad()
def local_decoder_pullback_zt(self, sample, timestep, encoder_hidden_states=None, op=None, block_idx=None,
                              pca_rank=50, chunk_size=25, min_iter=10, max_iter=100, convergence_threshold=None,
                              v0=None, return_info=False):
    """`utils.py:818-898`: the subspace iteration on the DECODER Jacobian d eps / d h at h = get_h(sample) (tangents enter at the
    mid-block output, the skip connections are constants).  Returns the reference's triple: `u` [numel(h), k] (directions in
    h-space), `s` [k] (square roots of the singular values of U^T J, as the reference returns them), `vT` [k, numel(x_t)] (the
    images of the directions in eps-space) -- "decoder jac do not return vT" (`:895-896`).  `convergence_threshold=None` is the
    reference's default, with which its `allclose` raises once `i > min_iter`; here None runs `max_iter` iterations."""
    if not (op == "mid" and block_idx == 0):
        raise AssertionError("up block is not implemented yet" if op == "up" else f"(op, block_idx) = ({op, block_idx}) is not valid")
    time_s = time.time()
    k = int(pca_rank)
    ctx_len = encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0
    eng = _engine_for(self, sample, "dec", 0, k, ctx_len)
    if v0 is None:
        vT = torch.randn(eng.n_in, k, device=sample.device, dtype=torch.float)           # utils.py:850-852
        vT, _ = torch.linalg.qr(vT)
        v0 = vT.T.contiguous()
    eng.set_point(sample, _timestep(timestep), encoder_hidden_states)
    tol = 0.0 if convergence_threshold is None else float(convergence_threshold)
    lo = max_iter if convergence_threshold is None else min_iter
    e_dirs, s, h_dirs, info = eng.pullback(v0, lo, max_iter, tol)
    print(f"power method : {info.iters_done - 1}-th step convergence : ", info.last_dist)
    print("power method runtime ==", time.time() - time_s)
    out = (h_dirs.T, s, e_dirs)
    if return_info:
        return out + (dict(iters_done=info.iters_done, converged=bool(info.converged), last_dist=info.last_dist),)
    return out


@torch.no_grad()
def inv_jac_zt(self, sample=None, timestep=None, encoder_hidden_states=None, op=None, block_idx=None, u=None, perturb_h=1e-1):
    """`utils.py:1117-1160`: the x-space direction that moves h along `u`: the normalised gradient of
    ||h + perturb_h u - get_h(x)|| at x = sample, i.e. -J^T u / ||J^T u|| -- one transpose pass of the encoder Jacobian (the
    perturbation size drops out: the residual at x is exactly perturb_h u).  `u` is one direction in h-space (the reference raises
    NotImplementedError for several)."""
    if u.dim() > 1 and u.shape[-1] != 1 and u.numel() != u.shape[0]:
        raise NotImplementedError("")
    assert sample.size(0) == 1, "sample size should be 1"
    ctx_len = encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0
    eng = _engine_for(self, sample, op, block_idx, 1, ctx_len)
    eng.set_point(sample, _timestep(timestep), encoder_hidden_states)
    w = eng.vjp(u.reshape(1, -1).to(sample.device, torch.float32))
    vT = -w
    return vT / vT.norm(dim=1, keepdim=True)


def patch_unet(unet):
    """The monkey-patch of `utils.py:103-104` (uncond) / `:326`, `:333` (Stable Diffusion)."""
    if hasattr(unet, "up_blocks") and unet_config(unet)["kind"] == 0:
        unet.get_h = types.MethodType(get_h, unet)
        unet.local_encoder_pullback_zt = types.MethodType(local_encoder_pullback_zt, unet)
        unet.local_encoder_pullback_many = types.MethodType(local_encoder_pullback_many, unet)
        unet.eps = types.MethodType(eps, unet)
        unet.get_h_to_e = types.MethodType(get_h_to_e, unet)                               # utils.py:327
        unet.local_decoder_pullback_zt = types.MethodType(local_decoder_pullback_zt, unet)  # utils.py:334 (commented out there)
        unet.inv_jac_zt = types.MethodType(inv_jac_zt, unet)                               # utils.py:329
    else:
        unet.get_h = types.MethodType(get_h_uncond, unet)
        unet.local_encoder_pullback_xt = types.MethodType(local_encoder_pullback_xt, unet)
        unet.local_encoder_pullback_many = types.MethodType(local_encoder_pullback_many, unet)
        unet.eps = types.MethodType(eps_uncond, unet)
    return unet
