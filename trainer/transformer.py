"""Drop-in for trainer/transformer.py: Transformer_2D().forward(src, flow) (fused warp kernel, fwd + bwd)."""
import _ctagan_path  # noqa: F401
from ctagan.nn import Transformer_2D  # noqa: F401
