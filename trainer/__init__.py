"""Drop-in for the reference's `trainer` package.  Fixes its packaging defects: `Reg_Trainer` is exported and
`Hd_Trainer_x` aliases the stage-1 trainer (train.py:42 asks users to rename one by hand)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import _ctagan_path  # noqa: F401,E402
from .CycTrainer import Cyc_Trainer  # noqa: F401,E402
from .HdTrainer import Hd_Trainer_x, Hd_Trainer_x1, Hd_Trainer_x2  # noqa: F401,E402
from .p2pTrainer import P2p_Trainer  # noqa: F401,E402
from .RegTrainer import Reg_Trainer  # noqa: F401,E402
