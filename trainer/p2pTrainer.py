"""Drop-in for the reference's trainer/p2pTrainer.py entry point(s): P2p_Trainer (see cta-gan_b200/ctagan/trainers.py)."""
import _ctagan_path  # noqa: F401
from ctagan.trainers import P2p_Trainer  # noqa: F401
