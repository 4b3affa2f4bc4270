"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

CPU (torch autograd, fp32) restatement of the reference's hot path.  Each function cites
the reference lines it follows.  The restatement is *pinned* against the reference's own
functions executed verbatim (see `oracle/reference_shim.py`, `scripts/make_golden.py` and
`tests/test_oracle.py`): the reference repository has no tests / golden vectors of its own
(SURVEY.md section 4), so golden vectors were minted by running the unmodified reference
functions in the authoring container and are committed under `tests/golden/`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product path never does.
"""
from __future__ import annotations

import torch


# ------------------------------------------------------------------------------------
# truncated U-Net forward
# ------------------------------------------------------------------------------------
def get_h(unet, sample, timestep, encoder_hidden_states=None, op=None, block_idx=None):
    """Follows `src/utils/utils.py:438-527` (SD `get_h`): time embedding, conv_in, down
    blocks (collecting skips), mid block (return for ('mid',0)), up blocks (return after
    up_blocks[block_idx] incl. its upsampler).  `op='down'` raises in the reference
    (TypeError from a wrong kwarg) and is rejected here too."""
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.float64 if isinstance(t, float) else torch.int64,
                         device=sample.device)
    elif t.dim() == 0:
        t = t[None].to(sample.device)
    t = t.expand(sample.shape[0])
    emb = unet.time_embedding(unet.time_proj(t).to(dtype=unet.dtype))
    x = unet.conv_in(sample)
    skips = (x,)
    if op == "down":
        raise ValueError(f"(op, block_idx) = ({op, block_idx}) is not valid")
    for blk in unet.down_blocks:
        if getattr(blk, "has_cross_attention", False):
            x, res = blk(hidden_states=x, temb=emb, encoder_hidden_states=encoder_hidden_states)
        else:
            x, res = blk(hidden_states=x, temb=emb)
        skips += res
    x = unet.mid_block(x, emb, encoder_hidden_states=encoder_hidden_states)
    if op == "mid" and block_idx == 0:
        return x
    if op == "up":
        for i, blk in enumerate(unet.up_blocks):
            n = len(blk.resnets)
            res, skips = skips[-n:], skips[:-n]
            if getattr(blk, "has_cross_attention", False):
                x = blk(hidden_states=x, temb=emb, res_hidden_states_tuple=res,
                        encoder_hidden_states=encoder_hidden_states)
            else:
                x = blk(hidden_states=x, temb=emb, res_hidden_states_tuple=res, upsample_size=None)
            if block_idx == i:
                return x
    raise ValueError(f"(op, block_idx) = ({op, block_idx}) is not valid")


def get_h_uncond(unet, x, t, op=None, block_idx=None):
    """Follows `src/utils/utils.py:114-163` (`get_h_uncond`); only ('mid',0) is valid."""
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.long, device=x.device)
    elif t.dim() == 0:
        t = t[None].to(x.device)
    t = t * torch.ones(x.shape[0], dtype=t.dtype, device=t.device)
    emb = unet.time_embedding(unet.time_proj(t).to(dtype=unet.dtype))
    h = unet.conv_in(x)
    for blk in unet.down_blocks:
        h, _ = blk(hidden_states=h, temb=emb)
    h = unet.mid_block(h, emb)
    if op == "mid" and block_idx == 0:
        return h
    raise ValueError(f"(op, block_idx) = ({op, block_idx}) is not valid")


def make_h_fn(unet, t, ctx, op, block_idx):
    """x[B,C,H,W] -> h[B,Co,Ho,Wo] closure, SD or uncond by model type."""
    if type(unet).__name__ == "UNet2DConditionModel":      # UNet2DModel has up_blocks too (full forward)
        def f(x):
            c = None if ctx is None else ctx.expand(x.shape[0], -1, -1)
            return get_h(unet, x, t, c, op, block_idx)
    else:
        def f(x):
            return get_h_uncond(unet, x, t, op, block_idx)
    return f


# ------------------------------------------------------------------------------------
# J and J^T applied to k directions (the two halves of one iteration)
# ------------------------------------------------------------------------------------
def jvp_columns(f, x, V):
    """U = J V, V:[k,C,H,W] -> U:[k,Co,Ho,Wo].  Reference: `utils.py:766-775` does this with
    jacfwd of a |-> get_h(x + a*vi) at a=0 (one dual-number forward over the tangent batch);
    identical to a forward-mode JVP at x with tangent vi for each batch element."""
    k = V.shape[0]
    _, u = torch.func.jvp(f, (x.expand(k, *x.shape[1:]).contiguous(),), (V.contiguous(),))
    return u


def vjp_rows(f, x, U):
    """W = U^T J, U:[k,Co,Ho,Wo] -> W:[k,n_in].  Reference: `utils.py:790-797`
    (`autograd.functional.jacobian` of x |-> einsum(u, get_h(x)))."""
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        h = f(xr)
        rows = []
        for i in range(U.shape[0]):
            g, = torch.autograd.grad((h * U[i:i + 1]).sum(), xr, retain_graph=i + 1 < U.shape[0])
            rows.append(g.reshape(1, -1))
    return torch.cat(rows, 0)


# ------------------------------------------------------------------------------------
# subspace iteration
# ------------------------------------------------------------------------------------
def initial_subspace(n_in, k, device="cpu", generator=None):
    """`utils.py:750-752`: V0 = qr(randn(n_in, k, dtype=float))[0].T"""
    v = torch.randn(n_in, k, device=device, dtype=torch.float, generator=generator)
    q, _ = torch.linalg.qr(v)
    return q.T.contiguous()


@torch.no_grad()
def local_encoder_pullback(unet, x, t, ctx=None, op="mid", block_idx=0, pca_rank=5,
                           min_iter=10, max_iter=50, convergence_threshold=1e-4, v0=None,
                           trace=None):
    """Follows `src/utils/utils.py:722-816` (zt / SD) and `:165-249` (xt / uncond): rank-k
    subspace iteration  U = J V ;  W = U^T J ;  (_, s, V) = svd(W) ; returns
    (u[n_out,k] = last U (un-normalised), sqrt(s)[k], vT[k,n_in] = last V).
    Chunking of the tangent batch (`:761-764`, `:178`) does not change results and is omitted.
    The early-exit test is the reference's sign-sensitive allclose with `i > min_iter`."""
    f = make_h_fn(unet, t, ctx, op, block_idx)
    h_shape = f(x).shape
    n_in = x[0].numel()
    k = pca_rank
    V = (initial_subspace(n_in, k, x.device) if v0 is None else v0.to(x.device, torch.float))
    V = V.reshape(k, *x.shape[1:]).to(x.dtype)
    U = s = None
    for i in range(max_iter):
        v_prev = V.detach().clone()
        U = jvp_columns(f, x, V)
        W = vjp_rows(f, x, U)
        _, s, Vh = torch.linalg.svd(W, full_matrices=False)
        V = Vh.reshape(k, *x.shape[1:])
        if trace is not None:
            trace.append(s.sqrt().clone())
        if torch.allclose(v_prev, V, atol=convergence_threshold) and i > min_iter:
            break
    n_out = h_shape[1] * h_shape[2] * h_shape[3]
    return U.reshape(k, n_out).T, s.sqrt(), V.reshape(k, n_in)


# ------------------------------------------------------------------------------------
# parity metrics
# ------------------------------------------------------------------------------------
def parity_report(s, vT, s_ref, vT_ref, u=None, u_ref=None):
    s, s_ref = s.double().cpu(), s_ref.double().cpu()
    V, Vr = vT.double().cpu(), vT_ref.double().cpu()
    rep = {"s_rel_max": float(((s - s_ref).abs() / s_ref).max())}
    cos = (V * Vr).sum(1).abs() / (V.norm(dim=1) * Vr.norm(dim=1))
    k = len(s_ref)
    gaps = []
    for i in range(k):
        nb = [abs(float(s_ref[i] - s_ref[j])) for j in (i - 1, i + 1) if 0 <= j < k]
        gaps.append(min(nb) / float(s_ref[i]) if nb else 1.0)
    rep["cos"] = [float(c) for c in cos]
    rep["gap"] = gaps
    rep["cos_min_gapped"] = min([float(c) for c, g in zip(cos, gaps) if g > 1e-2] or [1.0])
    rep["subspace"] = float((Vr @ V.T).pow(2).sum() / k)
    if u is not None:
        A, B = u.double().cpu(), u_ref.double().cpu()
        A, B = A / A.norm(dim=0, keepdim=True), B / B.norm(dim=0, keepdim=True)
        cu = (A * B).sum(0).abs()
        rep["u_cos"] = [float(c) for c in cu]
        rep["u_subspace"] = float((B.T @ A).pow(2).sum() / k)
    return rep


# ------------------------------------------------------------------------------------
# decoder side
# ------------------------------------------------------------------------------------
def get_h_to_e(unet, sample, timestep, encoder_hidden_states=None, input_h=None, op=None, block_idx=None):
    """Follows `src/utils/utils.py:529-635`: the forward up to the mid block on `sample` (one latent), then the mid-block
    output is REPLACED by the rows of `input_h` ([pca_rank, C, H, W]) with the skip connections and the prompt repeated
    pca_rank times, then up blocks, conv_norm_out, SiLU, conv_out.  Only ('mid', 0) substitutes (the reference asserts
    `op in ['mid', 'down']` and has no substitution for 'down')."""
    assert op in ("mid", "down"), "up block is not implemented yet"
    k = input_h.size(0)
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.float64 if isinstance(t, float) else torch.int64, device=sample.device)
    elif t.dim() == 0:
        t = t[None].to(sample.device)
    t = t.expand(sample.shape[0])
    emb = unet.time_embedding(unet.time_proj(t).to(dtype=unet.dtype))
    x = unet.conv_in(sample)
    skips = (x,)
    for blk in unet.down_blocks:
        if getattr(blk, "has_cross_attention", False):
            x, res = blk(hidden_states=x, temb=emb, encoder_hidden_states=encoder_hidden_states)
        else:
            x, res = blk(hidden_states=x, temb=emb)
        skips += res
    x = unet.mid_block(x, emb, encoder_hidden_states=encoder_hidden_states)
    if op == "mid" and block_idx == 0:
        x = input_h.view(k, *x.shape[1:])
        skips = tuple(s.repeat(k, 1, 1, 1) for s in skips)
        encoder_hidden_states = encoder_hidden_states.repeat(k, 1, 1)
    for blk in unet.up_blocks:
        n = len(blk.resnets)
        res, skips = skips[-n:], skips[:-n]
        if getattr(blk, "has_cross_attention", False):
            x = blk(hidden_states=x, temb=emb, res_hidden_states_tuple=res, encoder_hidden_states=encoder_hidden_states)
        else:
            x = blk(hidden_states=x, temb=emb, res_hidden_states_tuple=res, upsample_size=None)
    if unet.conv_norm_out:
        x = unet.conv_act(unet.conv_norm_out(x))
    return unet.conv_out(x)


def local_decoder_pullback_zt(unet, sample, timestep, encoder_hidden_states=None, op=None, block_idx=None, pca_rank=50,
                              min_iter=10, max_iter=100, convergence_threshold=None, v0=None):
    """Follows `src/utils/utils.py:818-898`: the same subspace iteration on g(h) = get_h_to_e(h) at h = get_h(sample):
    U = J_dec V by forward-mode JVPs, W = U^T J_dec by one reverse pass per row, SVD of W.  Returns the reference's triple
    (v.T [numel(h), k], sqrt(s), u [k, numel(x)]) -- it returns the h-space directions first (`:895-896`).  With
    `convergence_threshold=None` (the reference's default, whose `allclose(atol=None)` raises) all `max_iter` iterations run."""
    h = get_h(unet, sample, timestep, encoder_hidden_states, op=op, block_idx=block_idx)
    g = lambda hh: get_h_to_e(unet, sample, timestep, encoder_hidden_states, input_h=hh, op=op, block_idx=block_idx)
    n_h = h[0].numel()
    if v0 is None:
        vT = torch.randn(n_h, pca_rank, device=sample.device, dtype=torch.float)
        vT, _ = torch.linalg.qr(vT)
        v = vT.T
    else:
        v = v0
    v = v.reshape(-1, *h.shape[1:])
    s = u = None
    for i in range(max_iter):
        v_prev = v.detach().clone()
        u = torch.cat([torch.func.jvp(g, (h,), (vi[None],))[1].detach() for vi in v], 0)                 # [k, c, H, W]
        w = torch.autograd.functional.jacobian(lambda hh: (u * g(hh)).flatten(1).sum(1), h).reshape(-1, n_h)
        _, s, vv = torch.linalg.svd(w, full_matrices=False)
        v = vv.view(-1, *h.shape[1:])
        if convergence_threshold is not None and torch.allclose(v_prev, v, atol=convergence_threshold) and i > min_iter:
            break
    return v.reshape(-1, n_h).T.detach(), s.sqrt().detach(), u.reshape(u.shape[0], -1).detach()


def inv_jac_zt(unet, sample, timestep, encoder_hidden_states=None, op=None, block_idx=None, u=None, perturb_h=1e-1):
    """Follows `src/utils/utils.py:1117-1160`: the normalised gradient of ||h + perturb_h u - get_h(x)|| at x = sample."""
    h = get_h(unet, sample, timestep, encoder_hidden_states, op=op, block_idx=block_idx)
    perturbed_h = (h + perturb_h * u.view(-1, *h.shape[1:])).detach()
    jacx = lambda x: (perturbed_h - get_h(unet, x, timestep, encoder_hidden_states, op=op, block_idx=block_idx)).view(1, -1).norm(dim=-1)
    jac = torch.autograd.functional.jacobian(jacx, sample)
    vT = jac.view(1, -1)
    return vT / vT.norm(dim=1, keepdim=True)
