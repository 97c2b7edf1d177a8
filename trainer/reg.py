"""Drop-in for trainer/reg.py: Reg(height, width, in_channels_a, in_channels_b).forward(img_a, img_b) -> flow."""
import _ctagan_path  # noqa: F401
from ctagan.nn import Reg, ResUnet  # noqa: F401
