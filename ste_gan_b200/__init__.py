"""ste_gan_b200 - B200-native (sm_100a) implementation of the STE-GAN data-parallel hot path.

Drop-in mirrors of the reference's `ste_gan.models`, `ste_gan.layers` and `ste_gan.losses`
interfaces over a C-ABI library of hand-written CUDA kernels (include/stegan_b200.h).
"""
from .constants import *  # noqa: F401,F403
from .precision import get_precision, precision, set_precision  # noqa: F401

__version__ = "0.1.0"
