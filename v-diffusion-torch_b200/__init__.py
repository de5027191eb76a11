"""B200-native sampling hot path of tqch/v-diffusion-torch behind the reference's own Python surface.

    from v_diffusion_b200 import UNet, GaussianDiffusion, get_logsnr_schedule, fill_with_defaults, DATA_INFO

mirrors ``from v_diffusion import ...`` for the names generate.py uses on this path.
"""
from .config import DATA_INFO, fill_with_defaults, load_config, build_from_config  # noqa: F401
from .diffusion import GaussianDiffusion, get_logsnr_schedule  # noqa: F401
from .unet import UNet  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["UNet", "GaussianDiffusion", "get_logsnr_schedule", "fill_with_defaults", "DATA_INFO",
           "load_config", "build_from_config"]
