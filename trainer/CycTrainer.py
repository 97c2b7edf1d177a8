"""Drop-in for the reference's trainer/CycTrainer.py entry point(s): Cyc_Trainer (see cta-gan_b200/ctagan/trainers.py)."""
import _ctagan_path  # noqa: F401
from ctagan.trainers import Cyc_Trainer  # noqa: F401
