"""CPU tests of the oracle itself: (1) the restated algorithm vs the reference's functions run
verbatim (only where /root/reference exists), (2) the restated diffusers U-Net vs the in-repo
DDPM U-Net (`src/models/ddpm/diffusion.py`) through a weight mapping, (3) the oracle vs the
committed golden vectors (runs everywhere), (4) an absolute anchor: dense-Jacobian svdvals."""
import contextlib
import io
import os
import types

import pytest
import torch

from oracle import pullback_oracle as PO
from oracle import reference_shim as RS
from oracle import unet_torch as UT

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not RS.available(), reason="/root/reference not on this machine")


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


@needs_ref
@pytest.mark.parametrize("name,op,bi", [("sd_tiny", "mid", 0), ("sd_tiny", "up", 1),
                                        ("sd_tiny_lin", "mid", 0), ("uncond_tiny", "mid", 0)])
def test_restatement_matches_verbatim_reference(name, op, bi):
    m = RS.bind(UT.build_unet(name))
    x, t, ctx = UT.synthetic_inputs(name)
    k, iters = 3, 4
    torch.manual_seed(0)
    with torch.no_grad():
        if ctx is not None:
            u, s, vT = _quiet(m.local_encoder_pullback_zt, x, t, ctx, op=op, block_idx=bi, pca_rank=k,
                              chunk_size=5, min_iter=iters, max_iter=iters, convergence_threshold=0.)
        else:
            u, s, vT = _quiet(m.local_encoder_pullback_xt, x, t, op=op, block_idx=bi, pca_rank=k,
                              chunk_size=2, min_iter=iters, max_iter=iters, convergence_threshold=0.)
    torch.manual_seed(0)
    u2, s2, vT2 = PO.local_encoder_pullback(m, x, t, ctx, op, bi, k, iters, iters, 0.)
    assert torch.allclose(s, s2, rtol=1e-5)
    rep = PO.parity_report(s2, vT2, s, vT, u2, u)
    assert rep["subspace"] > 0.9999 and min(rep["cos"]) > 0.9999 and min(rep["u_cos"]) > 0.9999


@needs_ref
def test_restated_unet2dmodel_matches_inrepo_ddpm():
    """Maps the weights of the reference's in-repo DDPM (the only U-Net source in the reference)
    onto the restated diffusers UNet2DModel and compares get_h at ('mid',0)."""
    PullBackDDPM, _ = RS.load_ddpm()
    ns = types.SimpleNamespace
    cfg = ns(model=ns(ch=32, out_ch=3, ch_mult=[1, 1, 2, 2], num_res_blocks=2, attn_resolutions=[8],
                      dropout=0.0, in_channels=3, resamp_with_conv=True),
             data=ns(image_size=32))
    torch.manual_seed(3)
    ddpm = PullBackDDPM(ns(config=cfg, device="cpu", dtype=torch.float32)).eval()
    m = UT.build_unet("uncond_tiny")
    sd = {}
    src = ddpm.state_dict()

    def cp(dst, s_, lin=False):
        for suf in ("weight", "bias"):
            w = src[f"{s_}.{suf}"]
            if lin and suf == "weight":
                w = w.reshape(w.shape[0], w.shape[1])
            sd[f"{dst}.{suf}"] = w

    cp("time_embedding.linear_1", "temb.dense.0"); cp("time_embedding.linear_2", "temb.dense.1")
    cp("conv_in", "conv_in")

    def res(dst, s_):
        for a, b in [("norm1", "norm1"), ("conv1", "conv1"), ("time_emb_proj", "temb_proj"),
                     ("norm2", "norm2"), ("conv2", "conv2")]:
            cp(f"{dst}.{a}", f"{s_}.{b}")
        if f"{s_}.nin_shortcut.weight" in src:
            cp(f"{dst}.conv_shortcut", f"{s_}.nin_shortcut")

    def att(dst, s_):
        for a, b in [("group_norm", "norm"), ("query", "q"), ("key", "k"), ("value", "v"),
                     ("proj_attn", "proj_out")]:
            cp(f"{dst}.{a}", f"{s_}.{b}", lin=a != "group_norm")

    for i in range(4):
        for j in range(2):
            res(f"down_blocks.{i}.resnets.{j}", f"down.{i}.block.{j}")
            if i == 2:
                att(f"down_blocks.{i}.attentions.{j}", f"down.{i}.attn.{j}")
        if i != 3:
            cp(f"down_blocks.{i}.downsamplers.0.conv", f"down.{i}.downsample.conv")
    res("mid_block.resnets.0", "mid.block_1"); att("mid_block.attentions.0", "mid.attn_1")
    res("mid_block.resnets.1", "mid.block_2")
    # up path + eps head (diffusion.py:96-126): diffusers up_blocks[i] is the DDPM level 3 - i
    for i in range(4):
        lvl = 3 - i
        for j in range(3):
            res(f"up_blocks.{i}.resnets.{j}", f"up.{lvl}.block.{j}")
            if lvl == 2:
                att(f"up_blocks.{i}.attentions.{j}", f"up.{lvl}.attn.{j}")
        if lvl != 0:
            cp(f"up_blocks.{i}.upsamplers.0.conv", f"up.{lvl}.upsample.conv")
    cp("conv_norm_out", "norm_out"); cp("conv_out", "conv_out")
    missing = m.load_state_dict(sd, strict=True)
    x, t, _ = UT.synthetic_inputs("uncond_tiny")
    with torch.no_grad():
        h_ref = ddpm.get_h(x, t, op="mid", block_idx=0)
        h = PO.get_h_uncond(m, x, t, "mid", 0)
    assert h.shape == h_ref.shape
    assert torch.allclose(h, h_ref, rtol=1e-4, atol=1e-5), float((h - h_ref).abs().max())
    # the full noise prediction: the reference's own forward (diffusion.py:145-203) against the restated UNet2DModel.forward
    with torch.no_grad():
        e_ref = ddpm(x, t)
        e = m(x, t)
    assert e.shape == e_ref.shape == x.shape
    assert torch.allclose(e, e_ref, rtol=1e-4, atol=1e-5), float((e - e_ref).abs().max())


@pytest.mark.parametrize("name,op,bi", [("sd_tiny", "mid", 0)])
def test_oracle_matches_dense_jacobian(name, op, bi):
    """Absolute anchor: top singular value of the dense Jacobian."""
    m = UT.build_unet(name)
    x, t, ctx = UT.synthetic_inputs(name)
    f = PO.make_h_fn(m, t, ctx, op, bi)
    J = torch.autograd.functional.jacobian(lambda z: f(z).reshape(-1), x).reshape(-1, x.numel())
    sv = torch.linalg.svdvals(J.double())
    torch.manual_seed(0)
    u, s, vT = PO.local_encoder_pullback(m, x, t, ctx, op, bi, 3, 40, 40, 0.)
    assert abs(float(s[0]) - float(sv[0])) / float(sv[0]) < 1e-3
    assert torch.allclose(u.norm(dim=0), s, rtol=2e-2)      # ||u_i|| ~= s_i (one half-step apart)
    assert torch.allclose(vT @ vT.T, torch.eye(3), atol=1e-5)


def _golden_files():
    if not os.path.isdir(GOLDEN):
        return []
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".pt") and not f.startswith("ddim_"))


@pytest.mark.parametrize("fname", _golden_files())
def test_oracle_reproduces_golden(fname):
    """Golden vectors were produced by the reference's verbatim functions
    (scripts/make_golden.py); the oracle must reproduce them on this machine."""
    g = torch.load(os.path.join(GOLDEN, fname))
    if g["n_params"] > 60e6:
        pytest.skip("full-size golden: covered by the gpu parity test")
    m = UT.build_unet(g["config"], build_up=g["op"] == "up")
    x, t, ctx = UT.synthetic_inputs(g["config"])
    u, s, vT = PO.local_encoder_pullback(m, x, t, ctx, g["op"], g["block_idx"], g["k"],
                                         g["iters"], g["iters"], 0., v0=g["v0"])
    rep = PO.parity_report(s, vT, g["s"], g["vT"])
    assert rep["s_rel_max"] < 1e-4, rep
    assert rep["subspace"] > 0.9995 and rep["cos_min_gapped"] > 0.995, rep


# ---- DDIM scheduler: the restatement against the reference's own functions, run verbatim ----
@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
@pytest.mark.parametrize("inversion", [False, True])
def test_ddim_restatement_matches_verbatim_reference(inversion):
    import types
    from oracle import ddim_oracle as DO
    U = RS.load()
    ac = DO.sd_alphas_cumprod()
    ref = types.SimpleNamespace(t_max=999.0, alphas_cumprod=ac)
    ref.set_timesteps = types.MethodType(U.set_timesteps, ref)        # utils.py:273-286, bound like utils.py:340-342
    ref.step = types.MethodType(U.step, ref)                           # utils.py:288-315
    ours = DO.Scheduler(ac)
    ref.set_timesteps(12, is_inversion=inversion)
    ours.set_timesteps(12, is_inversion=inversion)
    assert torch.equal(torch.as_tensor(list(ref.timesteps)), torch.as_tensor(list(ours.timesteps)))
    assert torch.equal(torch.as_tensor(list(ref.timesteps_next)), torch.as_tensor(list(ours.timesteps_next)))
    g = torch.Generator().manual_seed(3)
    xt, et = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    for t in list(ref.timesteps)[:-1] if inversion else list(ref.timesteps):
        out = ref.step(et, t, xt, eta=0)
        x2, p2 = ours.step(et, t, xt)
        assert torch.equal(out.prev_sample, x2) and torch.equal(out.x0, p2)              # SchedulerOutput: utils.py:1166-1169
        assert torch.equal(U.extract(ac, t, xt.shape), DO.extract(ac, t, xt.shape))


def test_full_forward_is_the_up_path_plus_the_eps_head():
    m = UT.build_unet("sd_tiny")
    x, t, ctx = UT.synthetic_inputs("sd_tiny")
    e = m(x, t, encoder_hidden_states=ctx)
    assert e.shape == x.shape
    hl = PO.get_h(m, x, t, ctx, "up", 3)
    assert torch.equal(e, m.conv_out(torch.nn.functional.silu(m.conv_norm_out(hl))))


@pytest.mark.parametrize("name", ["sd_tiny", "sd_small"])
def test_ddim_oracle_reproduces_golden(name):
    """oracle/ddim_oracle.py against the trajectories minted from the reference's verbatim scheduler functions."""
    from oracle import ddim_oracle as DO
    g = torch.load(os.path.join(GOLDEN, f"ddim_{name}.pt"))
    m = UT.build_unet(name)
    z0, t, ctx = UT.synthetic_inputs(name)
    assert torch.allclose(m(z0, t, encoder_hidden_states=ctx), g["eps0"], rtol=1e-5, atol=1e-6)
    s = DO.Scheduler(DO.sd_alphas_cumprod())
    zT = DO.ddim_inversion(m, s, z0, ctx, g["inv_steps"])
    assert torch.allclose(zT, g["zT"], rtol=1e-5, atol=1e-6)
    z, te, ie = DO.ddim_forward_steps(m, s, zT, ctx, g["for_steps"], 0, g["t_end_idx"], g["guidance_scale"], g["neg"])
    assert ie == g["idx_edit"] and float(te) == g["t_edit"] and torch.allclose(z, g["z_edit"], rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
def test_x_space_guidance_restatement_matches_verbatim_reference():
    """oracle x_space_guidance against `EditStableDiffusion.x_space_guidance` (edit.py:484-502) run verbatim on a stand-in self."""
    from oracle import ddim_oracle as DO
    RS.load()
    import modules.edit as E                                   # the reference module, imported in place
    m = UT.build_unet("sd_tiny")
    zt, _, ctx = UT.synthetic_inputs("sd_tiny")
    s = DO.Scheduler(DO.sd_alphas_cumprod())
    s.set_timesteps(10)
    vk = torch.randn(zt.shape, generator=torch.Generator().manual_seed(2))
    vk = vk / vk.norm()
    me = types.SimpleNamespace(scheduler=s, edit_prompt_emb=ctx, x_space_guidance_scale=0.7,
                               unet=lambda x, t, encoder_hidden_states: types.SimpleNamespace(sample=m(x, t, encoder_hidden_states=encoder_hidden_states)))
    ref = E.EditStableDiffusion.x_space_guidance(me, zt, 3, vk, 1.5)
    ours = DO.x_space_guidance(m, s, zt, 3, vk, 1.5, ctx, 0.7)
    assert torch.equal(ref, ours)


@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
def test_uncond_x_space_guidance_restatement_matches_verbatim_reference():
    """oracle x_space_guidance (no prompt) against `EditUncondDiffusion.x_space_guidance` (edit.py:1716-1734) run verbatim."""
    from oracle import ddim_oracle as DO
    RS.load()
    import modules.edit as E
    m = UT.build_unet("uncond_tiny")
    xt, _, _ = UT.synthetic_inputs("uncond_tiny")
    s = DO.Scheduler(torch.cumprod(1.0 - torch.linspace(1e-4, 2e-2, 1000), dim=0))
    s.set_timesteps(10)
    vk = torch.randn(xt.shape, generator=torch.Generator().manual_seed(2))
    vk = vk / vk.norm()
    me = types.SimpleNamespace(scheduler=s, x_space_guidance_scale=0.7, unet=lambda x, t: m(x, t))
    ref = E.EditUncondDiffusion.x_space_guidance(me, xt, 3, vk, 1.5)
    ours = DO.x_space_guidance(m, s, xt, 3, vk, 1.5, None, 0.7)
    assert torch.equal(ref, ours)


# ---- stochastic step, YHCustomScheduler and the unconditional loop: the restatements against the reference run verbatim ----
@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
def test_stochastic_step_restatement_matches_verbatim_reference():
    """eta != 0 (`utils.py:306-311`): same generator state -> the same `torch.randn_like` draw -> bit-identical x_next."""
    from oracle import ddim_oracle as DO
    U = RS.load()
    ac = DO.sd_alphas_cumprod()
    ref = types.SimpleNamespace(t_max=999.0, alphas_cumprod=ac)
    ref.set_timesteps = types.MethodType(U.set_timesteps, ref)
    ref.step = types.MethodType(U.step, ref)
    ours = DO.Scheduler(ac)
    ref.set_timesteps(9); ours.set_timesteps(9)
    g = torch.Generator().manual_seed(5)
    xt, et = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 3, 8, 8, generator=g)
    for eta in (1, 0.5):
        for t in list(ref.timesteps)[1:-1]:
            torch.manual_seed(11); out = ref.step(et, t, xt, eta=eta)
            torch.manual_seed(11); x2, p2 = ours.step(et, t, xt, eta=eta)
            assert torch.equal(out.prev_sample, x2) and torch.equal(out.x0, p2)


@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
@pytest.mark.parametrize("noise_schedule", [None, "linear", "cosine"])
def test_yh_scheduler_restatement_matches_verbatim_reference(noise_schedule):
    """`YHCustomScheduler` (`utils.py:1171-1286`): SNR tables, timesteps and both step branches against the class itself."""
    from oracle import ddim_oracle as DO
    U = RS.load()
    args = types.SimpleNamespace(noise_schedule=noise_schedule, device="cpu", dtype=torch.float32)
    ref = U.YHCustomScheduler(args)
    betas, ac = DO.yh_schedule("linear" if noise_schedule is None else noise_schedule)
    assert torch.equal(torch.as_tensor(ref.betas), betas) and torch.equal(torch.as_tensor(ref.alphas_cumprod), ac)
    ours = DO.Scheduler(ac, t_max=999)
    ref.set_timesteps(7); ours.set_timesteps(7)
    assert torch.equal(torch.as_tensor(list(ref.timesteps)), torch.as_tensor(list(ours.timesteps)))
    g = torch.Generator().manual_seed(1)
    xt, et = torch.randn(1, 3, 8, 8, generator=g), torch.randn(1, 3, 8, 8, generator=g)
    for eta in (0, 1):
        t = list(ref.timesteps)[2]
        torch.manual_seed(4); out = ref.step(et, t, xt, eta=eta)
        torch.manual_seed(4); x2, p2 = ours.step(et, t, xt, eta=eta)
        assert torch.equal(out.prev_sample, x2) and torch.equal(out.x0, p2)


@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
@pytest.mark.parametrize("kw", [dict(t_start_idx=0, t_end_idx=-1), dict(t_start_idx=1, t_end_idx=3), dict(t_start_idx=2, t_end_idx=2),
                                dict(t_start_idx=0, t_end_idx=-1, performance_boosting=True)])
def test_uncond_forward_loop_restatement_matches_verbatim_reference(kw, tmp_path):
    """oracle `ddim_forward_steps_uncond` against `EditUncondDiffusion.DDIMforwardsteps` (edit.py:1601-1714) run verbatim on a
    stand-in self: end-before-skip ordering, eta = 1 under performance boosting."""
    from oracle import ddim_oracle as DO
    RS.load()
    import modules.edit as E
    m = UT.build_unet("uncond_tiny")
    xt, _, _ = UT.synthetic_inputs("uncond_tiny")
    _, ac = DO.yh_schedule("linear")
    s_ref, s_our = DO.Scheduler(ac, t_max=999), DO.Scheduler(ac, t_max=999)
    s_ref_step = s_ref.step
    s_ref.step = lambda et, t, x, eta=0.0, **k: types.SimpleNamespace(prev_sample=s_ref_step(et, t, x, eta=eta)[0])
    me = types.SimpleNamespace(for_steps=5, use_yh_custom_scheduler=True, scheduler=s_ref, device="cpu", dtype=torch.float32,
                               buffer_device="cpu", memory_bound=5, performance_boosting_t_idx=2, unet=lambda x, t: m(x, t),
                               result_folder=str(tmp_path), obs_folder=str(tmp_path), EXP_NAME="t")
    pb = kw.get("performance_boosting", False)
    torch.manual_seed(3)
    ref = E.EditUncondDiffusion.DDIMforwardsteps(me, xt.clone(), kw["t_start_idx"], kw["t_end_idx"], save_image=False,
                                                 performance_boosting=pb)
    torch.manual_seed(3)
    ours = DO.ddim_forward_steps_uncond(m, s_our, xt.clone(), 5, kw["t_start_idx"], kw["t_end_idx"], performance_boosting=pb,
                                        performance_boosting_t_idx=2)
    if isinstance(ref, tuple):
        assert isinstance(ours, tuple) and ref[2] == ours[2] and float(ref[1]) == float(ours[1]) and torch.equal(ref[0], ours[0])
    else:
        assert torch.equal(ref, ours)


# ---- decoder side: the restatements against the reference's own functions run verbatim ----
@pytest.mark.skipif(not RS.available(), reason="reference sources not on this machine")
def test_decoder_side_restatements_match_verbatim_reference():
    """`get_h_to_e` (utils.py:529-635), `local_decoder_pullback_zt` (:818-898) and `inv_jac_zt` (:1117-1160) bound onto the oracle
    U-Net exactly like utils.py:327-334 and run verbatim; the reference iteration is given an explicit threshold (its default
    `None` makes `torch.allclose` raise once i > min_iter) and a chunk size equal to the rank (its `v.chunk(k // chunk_size)`)."""
    U = RS.load()
    m = UT.build_unet("sd_tiny")
    x, t, ctx = UT.synthetic_inputs("sd_tiny")
    m.after_res = m.after_sa = False                                   # read by get_h_to_e (utils.py:540)
    m.get_h = types.MethodType(U.get_h, m)
    m.get_h_to_e = types.MethodType(U.get_h_to_e, m)
    h = PO.get_h(m, x, t, ctx, op="mid", block_idx=0)
    hs = torch.cat([h, 0.5 * h + 0.1, -h], 0)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = m.get_h_to_e(sample=x, timestep=t, encoder_hidden_states=ctx, input_h=hs, op="mid", block_idx=0)
    ours = PO.get_h_to_e(m, x, t, ctx, input_h=hs, op="mid", block_idx=0)
    assert torch.equal(ref, ours)
    # substituting the true h reproduces the full forward
    assert torch.allclose(ours[:1], m(x, t, encoder_hidden_states=ctx), atol=1e-5, rtol=1e-4)
    k = 2
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ur, sr, vr = U.local_decoder_pullback_zt(m, x, t, ctx, op="mid", block_idx=0, pca_rank=k, chunk_size=k, min_iter=1, max_iter=3,
                                                 convergence_threshold=0.0)
    torch.manual_seed(0)
    uo, so, vo = PO.local_decoder_pullback_zt(m, x, t, ctx, op="mid", block_idx=0, pca_rank=k, min_iter=1, max_iter=3, convergence_threshold=0.0)
    assert ur.shape == uo.shape == (h[0].numel(), k) and vr.shape == vo.shape == (k, x[0].numel())
    assert torch.allclose(sr, so, rtol=1e-4) and torch.allclose(ur.abs(), uo.abs(), atol=1e-4) and torch.allclose(vr.abs(), vo.abs(), atol=1e-4, rtol=1e-3)
    # inv_jac_zt
    m.inv_jac_zt = types.MethodType(U.inv_jac_zt, m)
    ud = torch.randn(h[0].numel(), generator=torch.Generator().manual_seed(3))
    with contextlib.redirect_stdout(io.StringIO()):
        vref = m.inv_jac_zt(sample=x, timestep=t, encoder_hidden_states=ctx, op="mid", block_idx=0, u=ud)
    vour = PO.inv_jac_zt(m, x, t, ctx, op="mid", block_idx=0, u=ud)
    assert torch.allclose(vref, vour, atol=1e-6)
    # ... which is -J^T u normalised
    w = torch.autograd.functional.vjp(lambda z: PO.get_h(m, z, t, ctx, op="mid", block_idx=0), x, ud.view_as(h))[1].reshape(1, -1)
    assert torch.allclose(vour, -w / w.norm(), atol=1e-5)
