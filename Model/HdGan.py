"""Drop-in for the reference's Model/HdGan.py: Generator/Discriminator (identical to CycleGan's), the feature-tap
discriminators and the LSGAN GANLoss.  The dead DataPrefetcher of the reference (never instantiated) is not reproduced."""
import _ctagan_path  # noqa: F401
from ctagan.nn import (Discriminator, Discriminator_m, GANLoss, Generator, NLayerDiscriminator,  # noqa: F401
                       ResidualBlock)

__all__ = ["ResidualBlock", "Generator", "Discriminator", "NLayerDiscriminator", "Discriminator_m", "GANLoss"]
