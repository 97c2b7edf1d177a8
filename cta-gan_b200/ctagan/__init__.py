"""ctagan: B200-native (sm_100a) implementation of the CTA-GAN training/inference hot path.

Python is host-side plumbing only (module surface, schedules, streams); all arithmetic on the path runs in the
hand-written CUDA kernels of libctagan.so (C ABI: include/ctagan.h).  There is no CPU or PyTorch fallback.
"""
from .engine import get_precision, invalidate_weight_cache, set_conv_engine, set_precision  # noqa: F401
from .nn import (Discriminator, Discriminator_m, GANLoss, Generator, L1Loss, MSELoss, NLayerDiscriminator, Reg,  # noqa: F401
                 ResidualBlock, ResUnet, Transformer_2D, l1_loss, masked_l1_loss, mse_const, plane_mean, smooothing_loss)
