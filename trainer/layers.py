"""Drop-in for the live building blocks of the reference's trainer/layers.py -- Conv (:71-104), DownBlock (:156-183),
ResnetTransformer (:216-240), ResnetBlock (:243-300) and get_init_function (:23-53) -- with the reference's constructor signatures,
sub-module names and initialisation draws, and forwards that run on the sm_100a kernels.  Inside `Reg` their arithmetic is fused into
the network's kernel schedule (ctagan.engine.reg_forward / reg_backward).  Dead helpers of the reference (UpBlock, AttentionGate) are
not reproduced."""
import _ctagan_path  # noqa: F401
from ctagan.nn import Conv, DownBlock, ResnetBlock, ResnetTransformer, get_init_function  # noqa: F401
