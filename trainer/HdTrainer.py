"""Drop-in for the reference's trainer/HdTrainer.py entry point(s): Hd_Trainer_x, Hd_Trainer_x1, Hd_Trainer_x2 (see cta-gan_b200/ctagan/trainers.py)."""
import _ctagan_path  # noqa: F401
from ctagan.trainers import Hd_Trainer_x, Hd_Trainer_x1, Hd_Trainer_x2  # noqa: F401
