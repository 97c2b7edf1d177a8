"""The live building blocks of the reference's trainer/layers.py (Conv, DownBlock, ResnetTransformer, ResnetBlock) exist
here only as parameter holders inside ctagan.nn.ResUnet: their arithmetic is fused into the Reg kernel schedule
(ctagan.engine.reg_forward / reg_backward).  Dead helpers of the reference (UpBlock, AttentionGate) are not reproduced."""
import _ctagan_path  # noqa: F401
from ctagan.nn import _DownBlockParams as DownBlock  # noqa: F401
from ctagan.nn import _ParamConv as Conv  # noqa: F401
from ctagan.nn import _ResnetBlockParams as ResnetBlock  # noqa: F401
from ctagan.nn import _ResnetTransformer as ResnetTransformer  # noqa: F401
