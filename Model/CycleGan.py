"""Drop-in for the reference's Model/CycleGan.py (same class names, constructor signatures and state_dict keys);
the implementation is the fused sm_100a schedule in cta-gan_b200/ctagan."""
import _ctagan_path  # noqa: F401
from ctagan.nn import Discriminator, Generator, ResidualBlock  # noqa: F401

__all__ = ["ResidualBlock", "Generator", "Discriminator"]
