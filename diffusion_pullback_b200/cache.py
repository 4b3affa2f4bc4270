"""The reference's on-disk contract for a computed local basis (`src/modules/edit.py:218-268`):
three `torch.save`d tensors `u-<name>.pt`, `s-<name>.pt`, `vT-<name>.pt` under
`./inputs/local_encoder_pullback_stable_diffusion-dataset_<ds>-num_steps_<n>-pca_rank_<k>/`, re-used when present, and the
normalisation the callers apply right after loading (`u / u.norm(dim=0)`, `vT / vT.norm(dim=1)`, `:267-268`).  Files written
here are readable by an unmodified reference run and vice versa."""
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


def load_or_compute_local_basis(unet, zt, t, prompt_emb, save_dir, name, op, block_idx, pca_rank, device=None, dtype=torch.float32,
                                **pullback_kwargs):
    """`edit.py:227-268`: load `u` / `vT` when both files exist, else run `unet.local_encoder_pullback_zt` with the reference's
    arguments (`chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-4` unless overridden), save the three
    tensors, and return the normalised `(u, s, vT)` (`s` is None on a cache hit, as in the reference)."""
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
    u, vT = normalize_basis(u, vT)
    return u, s, vT
