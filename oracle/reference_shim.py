"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

Imports the reference's hot-path functions *verbatim* from `/root/reference/src` (read-only,
present only in the authoring container) so that golden vectors can be minted from the
unmodified reference code and the restatement in `oracle/pullback_oracle.py` can be pinned
against it.  Nothing is copied: the reference modules are imported in place, with the
third-party packages that are absent offline (`diffusers`, `matplotlib`, `skimage`) stubbed in
`sys.modules`.  On the GPU box `/root/reference` does not exist and
`available()` returns False; nothing there may depend on this module.
"""
from __future__ import annotations

import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REF_SRC)


_cache = {}


def load():
    """Returns the reference's `utils.utils` module (get_h, get_h_uncond,
    local_encoder_pullback_zt, local_encoder_pullback_xt -- `src/utils/utils.py:114-249`,
    `:438-527`, `:722-816`)."""
    if "U" in _cache:
        return _cache["U"]
    if not available():
        raise RuntimeError("reference sources not present on this machine")
    for n in ["diffusers", "matplotlib", "matplotlib.pyplot", "skimage"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    for a in ["DDIMScheduler", "DDIMPipeline", "StableDiffusionPipeline"]:
        if not hasattr(sys.modules["diffusers"], a):
            setattr(sys.modules["diffusers"], a, object)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import utils.utils as U  # noqa: E402  (the reference module, unmodified)
    _cache["U"] = U
    return U


def load_ddpm():
    """The in-repo DDPM U-Net (`src/models/ddpm/diffusion.py:22-126`, `PullBackDDPM` `:131`)."""
    load()
    from models.ddpm.diffusion import PullBackDDPM, DDPM  # noqa: E402
    return PullBackDDPM, DDPM


def bind(unet):
    """Monkey-patches the reference methods onto `unet` exactly like `utils.py:103-104`,
    `:326`, `:333` do."""
    U = load()
    if type(unet).__name__ == "UNet2DConditionModel":      # UNet2DModel has up_blocks too (full forward)
        unet.get_h = types.MethodType(U.get_h, unet)
        unet.local_encoder_pullback_zt = types.MethodType(U.local_encoder_pullback_zt, unet)
    else:
        unet.get_h = types.MethodType(U.get_h_uncond, unet)
        unet.local_encoder_pullback_xt = types.MethodType(U.local_encoder_pullback_xt, unet)
    return unet
