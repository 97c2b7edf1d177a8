"""Drop-in for the reference's trainer/RegTrainer.py entry point(s): Reg_Trainer (see cta-gan_b200/ctagan/trainers.py)."""
import _ctagan_path  # noqa: F401
from ctagan.trainers import Reg_Trainer  # noqa: F401
