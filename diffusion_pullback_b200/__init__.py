"""pullback_b200 -- B200-native (sm_100a) implementation of Diffusion-Pullback's
`local_encoder_pullback_zt/xt` hot path.  See DESIGN.md."""
from .api import (eps, eps_uncond, get_h, get_h_to_e, get_h_uncond, inv_jac_zt, local_decoder_pullback_zt,  # noqa: F401
                  local_encoder_pullback_many, local_encoder_pullback_xt, local_encoder_pullback_zt,
                  patch_unet, refresh_weights)
from .cache import (load_or_compute_local_basis, local_basis_dir, local_basis_name, local_basis_paths,  # noqa: F401
                    normalize_basis, save_eigenvalue_spectrum, visualize_vT)
from .ddim import (DDIMSchedule, YHCustomScheduler, ddim_forward_steps, ddim_forward_steps_uncond, ddim_inversion,  # noqa: F401
                   x_space_guidance, x_space_guidance_edit)
from .engine import PullbackEngine, unet_config  # noqa: F401
