"""Developer probe: forward convolution of the 1-2 output channel layers -- tensor-core two-step kernel (default) against the generic
CUDA-core engine (parity) and timing; CTAGAN_THIN_TC=0 times the CUDA-core kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _ctagan_path  # noqa
import torch
from ctagan import engine as E, lib as L, ops
from dev_thin_time import graph_time, rel  # noqa: E402

CASES = [
    # name, N, Ci, Co, H, W (conv input incl. physical padding), K, stride, pad, act
    ("G tail 7x7 64->1 tanh b1", 1, 64, 1, 262, 262, 7, 1, 0, L.ACT_TANH),
    ("G tail 7x7 64->1 tanh b8", 8, 64, 1, 262, 262, 7, 1, 0, L.ACT_TANH),
    ("G tail 7x7 64->1 512 b4", 4, 64, 1, 518, 518, 7, 1, 0, L.ACT_TANH),
    ("Reg flow 3x3 32->2 b8", 8, 32, 2, 256, 256, 3, 1, 1, L.ACT_NONE),
    ("D patch 4x4 64->1 p1 b2", 2, 64, 1, 64, 64, 4, 1, 1, L.ACT_NONE),
]
for name, N, Ci, Co, H, W, K, s, p, act in CASES:
    x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
    w = torch.randn(Co, Ci, K, K, device="cuda") / (Ci * K * K) ** 0.5
    b = torch.randn(Co, device="cuda")
    prim = E.ConvPrim(w, b, s, p)
    E.set_conv_engine("generic"); ref = prim.fprop(x, act=act, use_bias=True)
    E.set_conv_engine("auto"); y = prim.fprop(x, act=act, use_bias=True)
    y2 = prim.fprop(x, act=act, use_bias=True)
    us = graph_time(lambda: prim.fprop(x, act=act, use_bias=True))
    by = x.numel() * 2 + y.numel() * 2
    print(f"{name:28s} {us:7.1f} us {by / us * 1e-3:7.1f} GB/s | err {rel(y, ref):.1e} | reproducible {bool(torch.equal(y, y2))}", flush=True)
