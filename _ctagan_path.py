"""Puts the `cta-gan_b200/` package directory on sys.path (its name is not a valid Python identifier)."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cta-gan_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
