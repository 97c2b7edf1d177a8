"""Drop-in for the reference's `trainer` package.  Fixes its packaging defects: `Reg_Trainer` is exported and
`Hd_Trainer_x` aliases the stage-1 trainer (train.py:42 asks users to rename one by hand)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import _ctagan_path  # noqa: F401,E402
