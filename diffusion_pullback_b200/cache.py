"""The reference's on-disk contract for a computed local basis (`src/modules/edit.py:218-268`):
three `torch.save`d tensors `u-<name>.pt`, `s-<name>.pt`, `vT-<name>.pt` under
`./inputs/local_encoder_pullback_stable_diffusion-dataset_<ds>-num_steps_<n>-pca_rank_<k>/`, re-used when present, and the
normalisation the callers apply right after loading (`u / u.norm(dim=0)`, `vT / vT.norm(dim=1)`, `:267-268`), plus the two pictures
the reference draws after a fresh computation (`:249-263`).  Files written here are readable by an unmodified reference run
and vice versa."""
from __future__ import annotations

import os

import torch


def local_basis_name(dataset_name, idx, edit_t, edit_prompt, op, block_idx, seed):
    """`edit.py:218`."""
    return f'local_basis-{dataset_name}_{idx}-{edit_t}T-"{edit_prompt}"-{op}-block_{block_idx}-seed_{seed}'


def local_basis_dir(dataset_name, for_steps, pca_rank, root="."):
    """`edit.py:220`."""
    return os.path.join(root, "inputs", f"local_encoder_pullback_stable_diffusion-dataset_{dataset_name}-num_steps_{for_steps}-pca_rank_{pca_rank}")


def local_basis_paths(save_dir, name):
    """`edit.py:223-225`: (u_path, s_path, vT_path)."""
    return tuple(os.path.join(save_dir, f"{p}-{name}.pt") for p in ("u", "s", "vT"))


def normalize_basis(u, vT):
    """`edit.py:267-268`: unit columns of u [n_out, k], unit rows of vT [k, n_in]."""
    return u / u.norm(dim=0, keepdim=True), vT / vT.norm(dim=1, keepdim=True)


def save_eigenvalue_spectrum(s, path):
    """`edit.py:249-252`: scatter plot of the singular values over their index (`plt.scatter(range(k), s, s=1)`).  matplotlib when
    it is installed (the reference's call, verbatim); otherwise the same picture drawn with PIL: 640 x 480 canvas, the
    matplotlib default axes box, one dot per value, the value range on the y axis."""
    vals = [float(v) for v in torch.as_tensor(s).detach().cpu().reshape(-1).tolist()]
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.scatter(range(len(vals)), vals, s=1)
        plt.savefig(path)
        plt.close()
        return path
    except ImportError:
        pass
    from PIL import Image, ImageDraw
    W, H, l, r, t, b = 640, 480, 80, 576, 58, 428                        # matplotlib's default figure / axes geometry
    img = Image.new("RGB", (W, H), "white")
    d = ImageDraw.Draw(img)
    d.rectangle([l, t, r, b], outline="black")
    if vals:
        lo, hi = min(vals), max(vals)
        pad = 0.05 * ((hi - lo) or max(abs(hi), 1.0))
        lo, hi = lo - pad, hi + pad
        n = max(len(vals) - 1, 1)
        for i, v in enumerate(vals):
            x = l + 0.05 * (r - l) + 0.9 * (r - l) * (i / n)
            y = b - (b - t) * (v - lo) / (hi - lo)
            d.ellipse([x - 1.5, y - 1.5, x + 1.5, y + 1.5], fill=(31, 119, 180))
        for frac in (0.0, 0.5, 1.0):
            y = b - (b - t) * frac
            d.line([l - 4, y, l, y], fill="black")
            d.text((8, y - 6), f"{lo + frac * (hi - lo):.4g}", fill="black")
        d.text((l, b + 8), "0", fill="black")
        d.text((r - 24, b + 8), str(len(vals) - 1), fill="black")
    img.save(path)
    return path


def visualize_vT(vT, latent_shape, path=None):
    """`edit.py:254-263`: project the k right singular vectors, viewed as [k, C, H, W] latents, onto the 3 principal axes of their
    channel vectors (`torch.pca_lowrank(q=3, center=True, niter=2)` over the k H W channel rows), min-max normalise and save as
    an image grid (`torchvision.utils.save_image`).  Returns the [k, 3, H, W] tensor in [0, 1]."""
    from einops import einsum
    lat = vT.reshape(-1, *latent_shape)
    pca_vT = lat.permute(0, 2, 3, 1).reshape(-1, latent_shape[0])
    _, _, pca_basis = torch.pca_lowrank(pca_vT, q=3, center=True, niter=2)
    vis = einsum(lat, pca_basis, "b c w h, c p -> b p w h")
    vis = vis - vis.min()
    vis = vis / vis.max()
    if path is not None:
        import torchvision.utils as tvu
        tvu.save_image(vis, path)
    return vis


def load_or_compute_local_basis(unet, zt, t, prompt_emb, save_dir, name, op, block_idx, pca_rank, device=None, dtype=torch.float32,
                                obs_folder=None, **pullback_kwargs):
    """`edit.py:227-268`: load `u` / `vT` when both files exist, else run `unet.local_encoder_pullback_zt` with the reference's
    arguments (`chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-4` unless overridden), save the three
    tensors -- and, like the reference right after a fresh computation, the eigenvalue-spectrum plot next to them
    (`eigenvalue_spectrum-<name>.png`, `:249-252`) and, when `obs_folder` is given, the PCA picture of vT (`vT-<name>.png`,
    `:254-263`) -- and return the normalised `(u, s, vT)` (`s` is None on a cache hit, as in the reference)."""
    os.makedirs(save_dir, exist_ok=True)
    u_path, s_path, vT_path = local_basis_paths(save_dir, name)
    device = device if device is not None else zt.device
    s = None
    if os.path.exists(u_path) and os.path.exists(vT_path):
        u = torch.load(u_path, map_location=device).type(dtype)
        vT = torch.load(vT_path, map_location=device).type(dtype)
    else:
        kw = dict(chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-4)
        kw.update(pullback_kwargs)
        u, s, vT = unet.local_encoder_pullback_zt(sample=zt, timestep=t, encoder_hidden_states=prompt_emb, op=op, block_idx=block_idx,
                                                  pca_rank=pca_rank, **kw)
        vT = vT.to(device=device, dtype=dtype)
        torch.save(u, u_path)
        torch.save(s, s_path)
        torch.save(vT, vT_path)
        save_eigenvalue_spectrum(s, os.path.join(save_dir, f"eigenvalue_spectrum-{name}.png"))
        if obs_folder is not None:
            os.makedirs(obs_folder, exist_ok=True)
            visualize_vT(vT, tuple(zt.shape[1:]), os.path.join(obs_folder, f"vT-{name}.png"))
    u, vT = normalize_basis(u, vT)
    return u, s, vT
